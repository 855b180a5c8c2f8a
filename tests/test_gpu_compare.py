"""GPU parity tests of the comparison models against the frozen outputs of the unmodified
reference (tests/golden/reference_vectors_cmp.npz) and the numpy oracle (oracle/compare_np.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import normwise, ROOT
from oracle import compare_np as CN

pytestmark = pytest.mark.gpu
TOL64 = 1e-10
TOL32 = 1e-4

DI_CASES = [("rk4_u8", "rk4", "rmse_X12", "rmse_U8"), ("euler_u8", "euler", "rmse_X12", "rmse_U8"),
            ("euler_u6", "euler", "rmse_X12", "rmse_W6"), ("quat_u6", "euler", "rmse_X13", "rmse_W6")]


@pytest.fixture(scope="module")
def B():
    import bluerov2_dynamics_b200 as b
    return b


@pytest.fixture(scope="module")
def EV():
    from bluerov2_dynamics_b200 import evaluators
    return evaluators


@pytest.fixture(scope="module")
def cg():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_cmp.npz")))


def cpu(t):
    return t.detach().cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------------------------------- double integrator
@pytest.mark.parametrize("tag,integ,xk,uk", DI_CASES)
def test_di_against_reference(EV, golden, cg, tag, integ, xk, uk):
    X, U, dt = golden[xk], golden[uk], float(cg["cmp_dt"])
    HS = [int(h) for h in cg["cmp_H"]]
    K_lin, K_ang = EV.estimate_di_gains(X[:100], U[:100], dt)
    assert np.allclose(K_lin, cg[f"di_{tag}_Klin"], rtol=1e-9, atol=1e-12)
    assert np.allclose(K_ang, cg[f"di_{tag}_Kang"], rtol=1e-9, atol=1e-12)
    K_lin, K_ang = cg[f"di_{tag}_Klin"], cg[f"di_{tag}_Kang"]
    traj = EV.simulate_double_integrator(X[0], U[:120], dt, K_lin, K_ang, integrator=integ)
    assert traj.shape == cg[f"di_{tag}_traj"].shape
    assert normwise(traj, cg[f"di_{tag}_traj"]) < TOL64
    got = EV.multistep_rmse_endpoint_di(X, U, HS, dt, K_lin, K_ang, integrator=integ)
    assert np.allclose(got, cg[f"di_{tag}_rmse"], rtol=1e-10)
    got32 = EV.multistep_rmse_endpoint_di(X, U, HS, dt, K_lin, K_ang, integrator=integ, dtype="f32")
    assert np.allclose(got32, cg[f"di_{tag}_rmse"], rtol=TOL32)


def test_di_edges(EV, golden, cg):
    dt = float(cg["cmp_dt"])
    r = EV.multistep_rmse_endpoint_di(golden["rmse_X12"][:5], golden["rmse_U8"][:5], 10, dt, cg["di_rk4_u8_Klin"],
                                      cg["di_rk4_u8_Kang"])
    assert np.isnan(r) and np.isnan(cg["di_rmse_nan_short"])
    traj = EV.simulate_double_integrator(cg["di_quat_u6_x0_scaled"], golden["rmse_W6"][3:43], dt, cg["di_quat_u6_Klin"],
                                         cg["di_quat_u6_Kang"], integrator="euler")
    assert normwise(traj, cg["di_quat_u6_traj_scaled"]) < TOL64
    with pytest.raises(ValueError):
        EV.simulate_double_integrator(np.zeros(12), np.zeros((4, 5)), dt, np.zeros((5, 3)), np.zeros((5, 3)))


@pytest.mark.parametrize("model,kind,integ,nx,nu", [("di12_u8", "di12", "rk4", 12, 8), ("di12_u6", "di12", "euler", 12, 6),
                                                    ("diq13_u6", "diq13", "euler", 13, 6)])
def test_di_ensemble_against_oracle(B, cg, model, kind, integ, nx, nu):
    rng = np.random.default_rng(31)
    n, T, dt = 1000, 60, 0.02
    x0 = rng.uniform(-1, 1, (n, nx))
    U = rng.uniform(-1, 1, (T, n, nu))
    K_lin, K_ang = rng.normal(0, 0.5, (nu, 3)), rng.normal(0, 0.5, (nu, 3))
    snaps, xT = CN.di_rollout(kind, integ, x0, U, dt, K_lin, K_ang, stride=20)
    for dtype, tol in (("f64", TOL64), ("f32", TOL32)):
        e = B.Engine(model, dtype)
        e.set_di_gains(K_lin, K_ang)
        r = e.rollout(x0, U, dt=dt, integrator=integ, stride=20)
        assert normwise(cpu(r.xT), xT) < tol and normwise(cpu(r.traj), snaps) < tol
