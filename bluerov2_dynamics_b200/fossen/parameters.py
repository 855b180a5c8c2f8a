"""Constants of the reduced 9-state model, exported under the reference's names (fossen/parameters.py:3-33).

Kept as one table so that the CUDA side (csrc/brov_api.cu: red9_consts) and this module can be compared line by
line; the module-level names are generated from it."""

_TABLE = {
    # rigid body / hydrostatics
    "m": 11.4, "g": 9.82,
    # added mass
    "X_ud": -2.6, "Y_vd": -18.5, "Z_wd": -13.3, "K_pd": -0.054, "M_qd": -0.0173, "N_rd": -0.28,
    # inertia
    "I_xx": 0.21, "I_yy": 0.245, "I_zz": 0.245,
    # linear damping
    "X_u": -0.09, "Y_v": -0.26, "Z_w": -0.19, "K_p": -0.895, "M_q": -0.287, "N_r": -4.64,
    # quadratic damping
    "X_uc": -34.96, "Y_vc": -103.25, "Z_wc": -74.23, "K_pc": -0.084, "M_qc": -0.028, "N_rc": -0.43,
    # centre of buoyancy offset
    "z_b": -0.1,
}
_TABLE["F_bouy"] = 1026 * 0.0115 * _TABLE["g"]

globals().update(_TABLE)
__all__ = sorted(_TABLE)
