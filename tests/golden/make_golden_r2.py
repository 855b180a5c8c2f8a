#!/usr/bin/env python3
"""Round-2 golden vectors, produced by EXECUTING THE UNMODIFIED REFERENCE (build container only):

    python tests/golden/make_golden_r2.py      ->  tests/golden/reference_vectors_r2.npz

  clamp_*    the cos(theta) clamp of euler_kinematics_matrix (fossen/BlueROV2.py:52-56: |cos| < 1e-7 -> 1e-7 sign(cos)):
             right-hand sides of the 8-thruster and the wrench 12-state model, and one explicit-Euler step, at pitch
             angles on and next to theta = +-pi/2 (both signs of the tiny cosine) plus one control row outside the clamp
  lagtail_*  the per-thruster hidden states `ThrusterLag._x` (fossen/BlueROV2.py:503-510) after RK4 and Euler rollouts
             of a few lengths (shorter and longer than the filter's memory), from non-zero initial lag states —
             what the engine's lag epilogue has to reproduce
Nothing from this repository's package is imported.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(OUT, "..", ".."))]
for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
           "matplotlib.lines", "matplotlib.cm", "matplotlib.colors"):
    sys.modules.setdefault(_m, MagicMock())

from fossen.BlueROV2 import BlueROV2 as RefThruster, euler_kinematics_matrix  # noqa: E402
from fossen.BlueROV2_thrust import BlueROV2 as RefWrench12  # noqa: E402

assert sys.modules["fossen"].__path__[0].startswith(REF)
G = {}
DT = 0.02
rng = np.random.default_rng(20261018)

# ---------------------------------------------------------------------------------------------- cos(theta) clamp
thetas = np.array([np.pi / 2, -np.pi / 2, np.pi / 2 + 5e-8, np.pi / 2 - 5e-8, -np.pi / 2 + 3e-8, -np.pi / 2 - 3e-8,
                   3 * np.pi / 2, np.pi / 2 - 1e-5])
K = len(thetas)
x = np.zeros((K, 12))
x[:, 0:3] = rng.uniform(-1, 1, (K, 3))
x[:, 3] = rng.uniform(-0.5, 0.5, K)
x[:, 4] = thetas
x[:, 5] = rng.uniform(-3, 3, K)
x[:, 6:9] = rng.uniform(-0.4, 0.4, (K, 3))
x[:, 9:12] = rng.uniform(-0.2, 0.2, (K, 3))
u8 = rng.uniform(-0.4, 0.4, (K, 8))
tau = rng.uniform(-1, 1, (K, 6)) * np.array([40, 40, 40, 5, 5, 5.0])
xd_thr, xd_w, x_euler_thr, x_euler_w, J2, cth = [], [], [], [], [], []
for i in range(K):
    xd = RefThruster().dynamics(x[i], u8[i], DT)
    xd_thr.append(xd)
    x_euler_thr.append(x[i] + DT * xd)                 # train_tank_brov2_full_comparison.py:462-465
    xdw = RefWrench12().dynamics(x[i], tau[i], DT)
    xd_w.append(xdw)
    x_euler_w.append(x[i] + DT * xdw)
    J2.append(euler_kinematics_matrix(x[i, 3], x[i, 4]))
    cth.append(np.cos(x[i, 4]))
G["clamp_x"], G["clamp_u8"], G["clamp_tau"] = x, u8, tau
G["clamp_xdot_thr"], G["clamp_xdot_wrench"] = np.array(xd_thr), np.array(xd_w)
G["clamp_euler_thr"], G["clamp_euler_wrench"] = np.array(x_euler_thr), np.array(x_euler_w)
G["clamp_J2"], G["clamp_cos_theta"] = np.array(J2), np.array(cth)
assert np.sum(np.abs(G["clamp_cos_theta"]) < 1e-7) == K - 1          # every row but the control is clamped


# ---------------------------------------------------------------------------------------------- per-thruster lag after a rollout
def smooth_inputs(T, nu, sigma=0.05):
    U = np.zeros((T, nu))
    u = np.zeros(nu)
    for k in range(T):
        u = np.clip(0.98 * u + sigma * rng.standard_normal(nu), -1.0, 1.0)
        U[k] = u
    return U


def lag_states(rov):
    return np.stack([np.array(l._x, float).reshape(3) for l in rov.thruster_lags])


def rollout(x0, U, lag0, rk4):
    rov = RefThruster()
    for i, l in enumerate(rov.thruster_lags):
        l._x = np.array(lag0[i], float).reshape(l._x.shape)
    x = x0.copy()
    for k in range(len(U)):
        u = U[k]
        if rk4:       # training/train_tank_brov2_rk4.py:388-393
            k1 = rov.dynamics(x, u, DT)
            k2 = rov.dynamics(x + 0.5 * DT * k1, u, DT)
            k3 = rov.dynamics(x + 0.5 * DT * k2, u, DT)
            k4 = rov.dynamics(x + DT * k3, u, DT)
            x = x + (DT / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        else:
            x = x + DT * rov.dynamics(x, u, DT)
    return x, lag_states(rov)


NV, TMAX = 5, 260
x0 = np.zeros((NV, 12))
x0[:, 0:3] = rng.uniform(-1, 1, (NV, 3))
x0[:, 3:5] = rng.uniform(-0.2, 0.2, (NV, 2))
x0[:, 5] = rng.uniform(-3, 3, NV)
x0[:, 6:12] = rng.uniform(-0.2, 0.2, (NV, 6))
U = np.stack([smooth_inputs(TMAX, 8) for _ in range(NV)], axis=1)            # [T, NV, 8]
lag0 = rng.uniform(-0.05, 0.05, (NV, 8, 3))
G["lagtail_x0"], G["lagtail_U"], G["lagtail_lag0"] = x0, U, lag0
for integ in ("rk4", "euler"):
    for T in (1, 7, 60, TMAX):
        xs, ls = zip(*[rollout(x0[i], U[:T, i], lag0[i], integ == "rk4") for i in range(NV)])
        G[f"lagtail_{integ}_T{T}_x"], G[f"lagtail_{integ}_T{T}_lag"] = np.array(xs), np.array(ls)

np.savez_compressed(os.path.join(OUT, "reference_vectors_r2.npz"), **G)
with open(os.path.join(OUT, "reference_vectors_r2.meta.txt"), "w") as f:
    import numpy, scipy
    f.write(f"generated by tests/golden/make_golden_r2.py from the unmodified reference at {REF}\n")
    f.write(f"numpy {numpy.__version__}, scipy {scipy.__version__}\n")
    for k in sorted(G):
        f.write(f"{k:28s} {str(G[k].shape):16s} {G[k].dtype}\n")
print("wrote", len(G), "arrays")
