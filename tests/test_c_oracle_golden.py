"""Pins the plain-C oracle (oracle/brov_oracle.c via oracle/c_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/reference_vectors.npz) and against the numpy oracle on a seeded ensemble."""
import os
import subprocess

import numpy as np
import pytest

from oracle import fossen_np as O
from conftest import normwise, ROOT

DT = 0.02
TOL = 1e-12


@pytest.fixture(scope="module")
def CO():
    from oracle import c_oracle
    if not c_oracle.available():
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    return c_oracle


def test_cfg1_thruster_rk4_1000_steps(golden, CO):
    U = np.tile(golden["cfg1_u_const"], (1000, 1))
    snaps, xT, lagT = CO.rollout("thruster8", "rk4", DT, golden["cfg1_x0"][None], U, stride=10)
    assert normwise(snaps[:, 0], golden["cfg1_const_traj_s10"][1:]) < TOL
    assert normwise(lagT[0], golden["cfg1_const_lagT"]) < TOL
    snaps, xT, lagT = CO.rollout("thruster8", "rk4", DT, golden["cfg1_x0"][None], golden["cfg1_U_var"], stride=10)
    assert normwise(snaps[:, 0], golden["cfg1_var_traj_s10"][1:]) < TOL
    assert normwise(lagT[0], golden["cfg1_var_lagT"]) < TOL


def test_cfg1_euler_dt001(golden, CO):
    U = np.tile(golden["cfg1_u_const"], (500, 1))
    snaps, xT, lagT = CO.rollout("thruster8", "euler", 0.01, golden["cfg1_x0"][None], U, stride=10)
    assert normwise(snaps[:, 0], golden["cfg1_euler_dt001_traj_s10"][1:]) < TOL
    assert normwise(lagT[0], golden["cfg1_euler_dt001_lagT"]) < TOL


@pytest.mark.parametrize("kind,integ,x0k,uk,outk", [
    ("thruster8", "rk4", "ens_x0", "ens_U8", "ens_thr_rk4_s20"),
    ("thruster8", "euler", "ens_x0", "ens_U8", "ens_thr_euler_s20"),
    ("wrench12", "rk4", "ens_x0", "ens_W6", "ens_w12_rk4_s20"),
    ("wrench12", "euler", "ens_x0", "ens_W6", "ens_w12_euler_s20"),
    ("quat13", "rk4", "ens_x0_q13", "ens_W6", "ens_q13_rk4_s20"),
    ("quat13", "euler", "ens_x0_q13", "ens_W6", "ens_q13_euler_s20"),
])
def test_ensembles(golden, CO, kind, integ, x0k, uk, outk):
    U = np.transpose(golden[uk], (1, 0, 2))
    snaps, xT, lagT = CO.rollout(kind, integ, DT, golden[x0k], U, stride=20)
    assert normwise(snaps, np.transpose(golden[outk], (1, 0, 2))[1:]) < TOL
    if kind == "thruster8":
        assert normwise(lagT, golden[f"ens_thr_{integ}_lagT"]) < TOL


def test_wrench_kat(golden, CO):
    _, xT, _ = CO.rollout("wrench12", "rk4", DT, golden["cfg1_x0"][None], np.tile(golden["kat_w12_rk4_tau"], (1000, 1)))
    assert normwise(xT[0], golden["kat_w12_rk4_xend"]) < TOL


def test_monte_carlo_params(golden, CO):
    from test_oracle_golden import _mc_params
    p = _mc_params(golden["mc_scales"])
    U = np.transpose(golden["mc_W6"], (1, 0, 2))
    _, xT, _ = CO.rollout("wrench12", "rk4", DT, golden["mc_x0"], U, params=p)
    assert normwise(xT, golden["mc_w12_rk4_xT"]) < TOL
    _, xT, _ = CO.rollout("quat13", "rk4", DT, golden["mc_x0_q13"], U, params=p)
    assert normwise(xT, golden["mc_q13_rk4_xT"]) < TOL


def test_multistep_rmse_reset(golden, CO):
    HS = [int(h) for h in golden["rmse_H"]]
    for integ in ("rk4", "euler"):
        got = [CO.multistep_se("thruster8", integ, DT, golden["rmse_X12"], golden["rmse_U8"], h)[2] for h in HS]
        assert np.allclose(got, golden[f"rmse_thr_{integ}_reset"], rtol=1e-11)
    got = [CO.multistep_se("wrench12", "euler", DT, golden["rmse_X12"], golden["rmse_W6"], h)[2] for h in HS]
    assert np.allclose(got, golden["rmse_w12_euler"], rtol=1e-11)
    got = [CO.multistep_se("quat13", "euler", DT, golden["rmse_X13"], golden["rmse_W6"], h)[2] for h in HS]
    assert np.allclose(got, golden["rmse_q13_euler"], rtol=1e-11)
    assert np.isnan(CO.multistep_se("wrench12", "euler", DT, golden["rmse_X12"][:5], golden["rmse_W6"][:5], 10)[2])


def test_sim_data_generator(golden, CO):
    snaps, _, _ = CO.rollout("thruster8", "euler", 0.05, np.zeros((1, 12)), golden["simgen_inputs"], stride=1)
    assert normwise(snaps[::10, 0], golden["simgen_states_true_s10"]) < TOL


def test_c_vs_numpy_oracle_ensemble(CO):
    """Two independent restatements agree on a seeded 512-vehicle ensemble with current and a carried lag state."""
    rng = np.random.default_rng(11)
    n, T = 512, 120
    x0 = np.zeros((n, 12))
    x0[:, :3] = rng.uniform(-2, 2, (n, 3))
    x0[:, 3:5] = rng.uniform(-0.2, 0.2, (n, 2))
    x0[:, 5] = rng.uniform(-3, 3, n)
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.05)
    lag0 = rng.normal(0, 0.1, (n, 8, 3))
    p = O.default_params(current=(0.1, -0.05, 0.02))
    s_np, x_np, l_np = O.rollout(O.Model("thruster8", DT, p), "rk4", x0, U, lag0=lag0, stride=40)
    s_c, x_c, l_c = CO.rollout("thruster8", "rk4", DT, x0, U, params=p, lag0=lag0, stride=40)
    assert normwise(s_c, s_np) < TOL and normwise(x_c, x_np) < TOL and normwise(l_c, l_np) < TOL
