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


# ------------------------------------------------------------------------------------------------------ Koopman
@pytest.fixture(scope="module")
def KM(cg):
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=116, gamma=float(cg["koop_gamma"]), ridge=1e-1)
    K.centers_, K.A_, K.B_ = cg["koop_centers"], cg["koop_A"], cg["koop_B"]
    return K


def test_koopman_against_reference(KM, cg):
    X, U = cg["koop_X"], cg["koop_U"]
    assert normwise(KM._lift(X[:16]), cg["koop_lift"]) < 1e-13
    assert KM._lift(X[3]).shape == (128,)
    assert np.isclose(KM.evaluate(X, U), float(cg["koop_evaluate"]), rtol=TOL64)
    got = [KM.multistep_rmse(X, U, int(h)) for h in cg["cmp_H"]]
    assert np.allclose(got, cg["koop_rmse"], rtol=TOL64)
    sim = KM.simulate(X[0], U[:160])
    assert sim.shape == cg["koop_simulate"].shape and normwise(sim, cg["koop_simulate"]) < TOL64
    assert np.isnan(KM.multistep_rmse(X[:5], U[:5], 10))


def test_koopman_full_size_against_oracle():
    """d = 512 (500 RBFs, the reference's configuration), H = 100: oracle = sequential propagation of the lifted state."""
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    rng = np.random.default_rng(8)
    n, r, k, T = 12, 8, 500, 700
    C = rng.uniform(-1, 1, (k, n))
    X = np.cumsum(0.02 * rng.standard_normal((T, n)), axis=0)
    U = rng.uniform(-1, 1, (T, r))
    A = 0.98 * np.linalg.qr(rng.standard_normal((n + k, n + k)))[0] + 0.01 * rng.standard_normal((n + k, n + k)) / np.sqrt(n + k)
    Bm = 0.05 * rng.standard_normal((n + k, r))
    K = KoopmanEDMDc(state_dim=n, input_dim=r, n_rbfs=k, gamma=0.7)
    K.centers_, K.A_, K.B_ = C, A, Bm
    for H in (1, 10, 100):
        ref = CN.koop_multistep_se(X, U, H, C, 0.7, A, Bm)[2]
        assert np.isclose(K.multistep_rmse(X, U, H), ref, rtol=TOL64), H
    ref = CN.koop_simulate(X[0], U[:300], C, 0.7, A, Bm)
    assert normwise(K.simulate(X[0], U[:300]), ref) < TOL64
    sims = K.simulate_batch(X[:5], U[:50])
    for b in range(5):
        assert normwise(cpu(sims[:, b]), CN.koop_simulate(X[b], U[:50], C, 0.7, A, Bm)[1:]) < TOL64


def test_koopman_fit_matches_reference(cg):
    """fit() = scikit-learn k-means + ridge normal equations on the host: same centres, (A, B) to solver rounding."""
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    # the golden model was fitted on rows 0..599 of a series whose rows 600.. are stored; refit on the stored part
    X, U = cg["koop_X"], cg["koop_U"]
    K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=20, gamma=3.0, ridge=1e-1)
    K.fit(X, U)
    assert K.A_.shape == (32, 32) and K.B_.shape == (32, 8) and K.lift_dim_ == 32
    Z, Zp = CN.koop_lift(X[:-1], K.centers_, 3.0), CN.koop_lift(X[1:], K.centers_, 3.0)
    Gm = np.hstack([Z, U[:-1]])
    M = (np.linalg.pinv(Gm.T @ Gm + 0.1 * np.eye(40)) @ (Gm.T @ Zp)).T
    assert np.allclose(K.A_, M[:, :32], atol=1e-9) and np.allclose(K.B_, M[:, 32:], atol=1e-9)
    assert K.evaluate(X, U) < 0.5


# --------------------------------------------------------------------------------------------------------- PINc
@pytest.fixture(scope="module")
def PM(cg):
    from bluerov2_dynamics_b200 import pinc
    sd = {k[len("pinc_sd_"):]: v for k, v in cg.items() if k.startswith("pinc_sd_")}
    return pinc, pinc.PincModel(sd)


def test_pinc_forward_against_reference(PM, cg):
    pinc, model = PM
    out = cpu(model(cg["pinc_dataset_zin"][:64]))
    assert normwise(out, cg["pinc_forward"]) < 2e-6     # float32 network, torch CPU vs CUDA rounding
    with pytest.raises(ValueError):
        model(np.zeros((3, 13), np.float32))


def test_pinc_rollout_and_rmse_against_reference(PM, golden, cg):
    pinc, model = PM
    from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2
    X12, U8, dt = golden["rmse_X12"], golden["rmse_U8"], float(cg["cmp_dt"])
    HS = [int(h) for h in cg["cmp_H"]]
    rov = BlueROV2(dt=dt)
    traj = pinc.simulate_pinc(X12[0], U8[:120], dt, model, rov, None)
    assert traj.shape == cg["pinc_traj"].shape
    assert normwise(traj, cg["pinc_traj"]) < TOL32
    # the rollout left rov's lag where 120 reference thruster-map calls leave it
    from oracle import fossen_np as O
    m = O.Model("thruster8", dt)
    lag = m.zero_lag(1)
    for u in U8[:120]:
        _, lag = CN.thruster_map4(u[None], lag, m.Ad, m.Bd, m.alloc)
    assert normwise(np.stack([l._x for l in rov.thruster_lags]), lag[0]) < 1e-12
    got = pinc.multistep_rmse_endpoint_pinc(X12, U8, HS, dt, model, None, None, lag_mode="reset")
    assert np.allclose(got, cg["pinc_rmse_reset"], rtol=TOL32)
    got = [pinc.multistep_rmse_endpoint_pinc(X12, U8, h, dt, model, BlueROV2(dt=dt), None) for h in HS]
    assert np.allclose(got, cg["pinc_rmse_carry"], rtol=TOL32)
    assert np.isnan(pinc.multistep_rmse_endpoint_pinc(X12[:5], U8[:5], 10, dt, model))


def test_pinc_helpers(PM, golden, cg):
    pinc, _ = PM
    X12 = golden["rmse_X12"]
    assert normwise(pinc.batch12_to_9(X12)[:-1], cg["pinc_dataset_zin"][:, :9]) < 1e-15
    assert np.allclose(pinc.dataset12_to_9(X12[3]), pinc.batch12_to_9(X12[3:4])[0])
    x9 = pinc.batch12_to_9(X12)
    back = pinc.batch9_to_12(x9)
    assert np.allclose(back[:, [0, 1, 2, 6, 7, 8, 11]], X12[:, [0, 1, 2, 6, 7, 8, 11]])
    assert np.allclose(np.cos(back[:, 5]), np.cos(X12[:, 5])) and np.all(back[:, [3, 4, 9, 10]] == 0)
    assert np.allclose(pinc.state9_to_12(x9[5]), back[5])
    from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2
    rov = BlueROV2(dt=0.02)
    u4 = np.stack([pinc.thrusters_to_body_wrenches(u, 0.02, rov) for u in golden["rmse_U8"][:20]])
    assert normwise(u4, cg["pinc_U4_carry"][:20]) < 1e-12


def test_pinc_ensemble_against_oracle(PM, cg):
    """4096 windows x 60 steps with per-window inputs and non-zero initial lag vs the numpy float32 oracle."""
    pinc, model = PM
    rng = np.random.default_rng(17)
    n, T, dt = 4096, 60, 0.02
    x0 = np.zeros((n, 12))
    x0[:, :3] = rng.uniform(-2, 2, (n, 3))
    x0[:, 5] = rng.uniform(-3, 3, n)
    x0[:, 6:9] = rng.uniform(-0.3, 0.3, (n, 3))
    x0[:, 11] = rng.uniform(-0.3, 0.3, n)
    U = rng.uniform(-0.6, 0.6, (T, n, 8))
    lag0 = rng.normal(0, 0.05, (n, 8, 3))
    snaps, x9, _ = CN.pinc_rollout(x0, U, dt, CN.pinc_weights(cg), lag0=lag0, stride=20)
    traj, x9g, _ = model.rollout(x0, U, dt, lag0=lag0, stride=20)
    assert normwise(cpu(traj), snaps) < TOL32
    assert normwise(cpu(x9g), x9) < TOL32


# ------------------------------------------------------------------------------------- data formats around the path
def test_generate_sim_dataset_matches_reference(golden):
    """training/train_sim_brov2_koopmanEDMDc.py:153-197 with seed 42: same inputs (bit for bit), same Euler rollout."""
    from bluerov2_dynamics_b200.datasets import generate_sim_dataset
    true, noisy, U = generate_sim_dataset(1500, dt=0.05, seed=42)
    assert np.array_equal(U, golden["simgen_inputs"])
    assert normwise(true[::10], golden["simgen_states_true_s10"]) < TOL64
    assert normwise(noisy[::10], golden["simgen_states_noisy_s10"]) < TOL64


def test_csv_series_through_the_evaluators(tmp_path, golden):
    from bluerov2_dynamics_b200.datasets import load_dataset, save_dataset
    from bluerov2_dynamics_b200.evaluators import multistep_rmse_endpoint_physics
    p = tmp_path / "series.csv"
    save_dataset(p, golden["rmse_X12"], golden["rmse_U8"], 0.02)
    X, U, dt = load_dataset(p, verbose=False)
    got = multistep_rmse_endpoint_physics(X, U, [1, 10, 100], dt, integrator="rk4", lag_mode="carry")
    assert np.allclose(got, golden["rmse_thr_rk4_carry"], rtol=1e-9)


# ------------------------------------------------------------------------------------------------ edge cases
def test_koopman_long_horizon_taps_from_global_memory():
    """d = 512 with H = 170: the FIR taps (170 x 12 x 8 doubles) no longer fit shared memory beside the model and are
    read from global memory; ragged window count (not a multiple of the 512-window tile)."""
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    rng = np.random.default_rng(3)
    n, r, k, T = 12, 8, 500, 170 + 777
    C = rng.uniform(-1, 1, (k, n))
    X = np.cumsum(0.02 * rng.standard_normal((T, n)), axis=0)
    U = rng.uniform(-1, 1, (T, r))
    A = 0.97 * np.linalg.qr(rng.standard_normal((n + k, n + k)))[0]
    Bm = 0.05 * rng.standard_normal((n + k, r))
    K = KoopmanEDMDc(state_dim=n, input_dim=r, n_rbfs=k, gamma=0.9)
    K.centers_, K.A_, K.B_ = C, A, Bm
    assert np.isclose(K.multistep_rmse(X, U, 170), CN.koop_multistep_se(X, U, 170, C, 0.9, A, Bm)[2], rtol=TOL64)
    # a second, shorter horizon after the long one reuses the cached decoder rows
    assert np.isclose(K.multistep_rmse(X, U, 3), CN.koop_multistep_se(X, U, 3, C, 0.9, A, Bm)[2], rtol=TOL64)


def test_pinc_ragged_windows_and_tail(PM, golden, cg):
    """Window counts that are not a multiple of the 512-window tile, a single window, and horizons longer than what is
    left of the series for the last windows (those horizons simply do not count them)."""
    pinc, model = PM
    dt = float(cg["cmp_dt"])
    L = CN.pinc_weights(cg)
    X12, U8 = golden["rmse_X12"], golden["rmse_U8"]
    Xr, Ur = np.tile(X12, (5, 1))[:613], np.tile(U8, (5, 1))[:613]
    for hs in ([1], [2, 5, 40]):
        ref = CN.pinc_multistep_se(Xr, Ur, hs, dt, L, "reset")
        se, cnt = model.multistep_se(Xr, Ur, hs, dt, "reset")
        se = se.cpu().numpy()
        assert cnt == [ref[h][1] for h in hs]
        assert np.allclose(se[:len(hs)], [ref[h][0] for h in hs], rtol=2e-4)
    traj, x9, _ = model.rollout(X12[7:8], U8[:9], dt)
    snaps, x9r, _ = CN.pinc_rollout(X12[7:8], U8[:9], dt, L)
    assert normwise(cpu(traj), snaps) < TOL32 and normwise(cpu(x9), x9r) < TOL32
    with pytest.raises(ValueError):
        model.multistep_se(Xr, Ur, [5, 2], dt)


def test_di_shard_invariance(B):
    """Rows of a sub-ensemble equal the same rows of the whole ensemble, bit for bit (no cross-vehicle coupling)."""
    rng = np.random.default_rng(2)
    n, T = 1500, 40
    x0, U = rng.uniform(-1, 1, (n, 12)), rng.uniform(-1, 1, (T, n, 8))
    e = B.Engine("di12_u8", "f64")
    e.set_di_gains(rng.normal(0, 0.3, (8, 3)), rng.normal(0, 0.3, (8, 3)))
    whole = cpu(e.rollout(x0, U, dt=0.02).xT)
    part = cpu(e.rollout(x0[700:1333], np.ascontiguousarray(U[:, 700:1333]), dt=0.02).xT)
    assert np.array_equal(whole[700:1333], part)


def test_koopman_model_edited_in_place_is_reuploaded(cg):
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=116, gamma=float(cg["koop_gamma"]))
    K.centers_, K.A_, K.B_ = cg["koop_centers"].copy(), cg["koop_A"].copy(), cg["koop_B"].copy()
    X, U = cg["koop_X"], cg["koop_U"]
    assert np.isclose(K.multistep_rmse(X, U, 10), cg["koop_rmse"][1], rtol=TOL64)
    K.A_ *= 0.5            # same array object, new contents
    ref = CN.koop_multistep_se(X, U, 10, K.centers_, K.gamma, K.A_, K.B_)[2]
    assert np.isclose(K.multistep_rmse(X, U, 10), ref, rtol=TOL64) and not np.isclose(ref, cg["koop_rmse"][1])


def test_koopman_multi_horizon_single_lift(KM, cg):
    """All horizons from one lift pass = the per-horizon calls = the reference."""
    X, U = cg["koop_X"], cg["koop_U"]
    hs = [int(h) for h in cg["cmp_H"]]
    got = KM.multistep_rmse_multi(X, U, hs)
    assert np.allclose(got, cg["koop_rmse"], rtol=TOL64)
    assert np.allclose(got, [KM.multistep_rmse(X, U, h) for h in hs], rtol=1e-13)
    five = KM.multistep_rmse_multi(X, U, [100, 1, 7, 10, 3, 250, 400])   # more than MAX_H, unsorted, one too long
    assert np.allclose(five[:2], [cg["koop_rmse"][2], cg["koop_rmse"][0]], rtol=TOL64) and np.isnan(five[-1])
    assert np.isclose(five[2], KM.multistep_rmse(X, U, 7), rtol=1e-13)


def test_koopman_random_shapes_against_oracle():
    """Random model sizes (n, r) in {(12,8), (12,6), (13,6)}, numbers of centres, series lengths and horizon sets."""
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    rng = np.random.default_rng(40)
    for case in range(12):
        n, r = [(12, 8), (12, 6), (13, 6)][case % 3]
        k = int(rng.integers(1, 300))
        T = int(rng.integers(5, 1500))
        C = rng.uniform(-1, 1, (k, n))
        X = np.cumsum(0.03 * rng.standard_normal((T, n)), axis=0)
        U = rng.uniform(-1, 1, (T, r))
        A = 0.95 * np.linalg.qr(rng.standard_normal((n + k, n + k)))[0]
        Bm = 0.1 * rng.standard_normal((n + k, r))
        gam = float(rng.uniform(0.1, 2.0))
        K = KoopmanEDMDc(state_dim=n, input_dim=r, n_rbfs=k, gamma=gam)
        K.centers_, K.A_, K.B_ = C, A, Bm
        hs = sorted(set(int(h) for h in rng.integers(1, min(T + 3, 90), 3)))
        got = K.multistep_rmse_multi(X, U, hs)
        for h, g in zip(hs, got):
            ref = CN.koop_multistep_se(X, U, h, C, gam, A, Bm)[2]
            assert (np.isnan(g) and np.isnan(ref)) or np.isclose(g, ref, rtol=TOL64), (case, n, r, k, T, h, g, ref)
        h = hs[0]
        if T > h:
            assert np.isclose(K.multistep_rmse(X, U, h), CN.koop_multistep_se(X, U, h, C, gam, A, Bm)[2], rtol=TOL64)


def test_pinc_random_shapes_against_oracle(PM, cg):
    pinc, model = PM
    rng = np.random.default_rng(41)
    L = CN.pinc_weights(cg)
    for case in range(6):
        n, T = int(rng.integers(1, 1300)), int(rng.integers(1, 40))
        stride = int(rng.choice([1, 2, 5, T]))
        x0 = np.zeros((n, 12))
        x0[:, :3] = rng.uniform(-2, 2, (n, 3))
        x0[:, 5] = rng.uniform(-3, 3, n)
        x0[:, 6:9] = rng.uniform(-0.2, 0.2, (n, 3))
        x0[:, 11] = rng.uniform(-0.2, 0.2, n)
        shared = bool(rng.random() < 0.4)
        U = rng.uniform(-0.5, 0.5, (T, 8) if shared else (T, n, 8))
        lag0 = rng.normal(0, 0.05, (n, 8, 3)) if rng.random() < 0.5 else None
        snaps, x9, _ = CN.pinc_rollout(x0, U, 0.02, L, lag0=lag0, stride=stride)
        traj, x9g, _ = model.rollout(x0, U, 0.02, lag0=lag0, stride=stride)
        assert traj.shape == snaps.shape, (case, traj.shape, snaps.shape)
        assert normwise(cpu(traj), snaps) < TOL32 and normwise(cpu(x9g), x9) < TOL32, (case, n, T, stride, shared)


def test_make_pinc_dataset_series_thruster_map(PM, golden, cg):
    """make_pinc_dataset: the thruster map along the whole series with a carried lag state in ONE launch equals the
    reference's loop of stateful compute_thruster_forces calls, and leaves the model object's lag where the loop does."""
    pinc, _ = PM
    from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2
    X12, U8, dt = golden["rmse_X12"], golden["rmse_U8"], float(cg["cmp_dt"])
    rov = BlueROV2(dt=dt)
    z_in, y, U4 = pinc.make_pinc_dataset(X12, U8, dt, rov)
    assert normwise(U4, cg["pinc_U4_carry"]) < TOL64
    assert normwise(z_in, cg["pinc_dataset_zin"]) < TOL64 and normwise(y, cg["pinc_dataset_y"]) < TOL64
    ref = BlueROV2(dt=dt)
    for u in U8:
        ref.compute_thruster_forces(u, dt)
    assert normwise(np.stack([l._x for l in rov.thruster_lags]), np.stack([l._x for l in ref.thruster_lags])) < 1e-12
    # a long series (far beyond the replay depth) and a non-zero initial lag state, fp32 and fp64, against the oracle
    import bluerov2_dynamics_b200 as B
    from oracle import fossen_np as O
    rng = np.random.default_rng(6)
    T = 5000
    U = O.smooth_inputs(rng, T, 8, sigma=0.05)
    lag0 = rng.normal(0, 0.1, (8, 3))
    m = O.Model("thruster8", dt)
    lag, rows = lag0[None].copy(), []
    for u in U:
        tau, lag = O.thruster_wrench(u[None], lag, m.Ad, m.Bd, m.alloc)
        rows.append(tau[0])
    for dtype, tol in (("f64", TOL64), ("f32", TOL32)):
        tau_g, lag_g = B.Engine("thruster8", dtype).thruster_wrench_series(U, lag0=lag0.reshape(24), dt=dt)
        assert normwise(cpu(tau_g), np.array(rows)) < tol and normwise(cpu(lag_g).reshape(8, 3), lag[0]) < tol
