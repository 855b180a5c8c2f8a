"""GPU parity tests: the CUDA engine (through the C ABI, via bluerov2_dynamics_b200.Engine) against
 (a) the frozen outputs of the unmodified reference (tests/golden/reference_vectors.npz), and
 (b) the numpy oracle (oracle/fossen_np.py) on seeded inputs at sizes the oracle finishes in seconds.

Tolerances (BASELINE.json north_star): fp64 <= 1e-10 normwise relative, ||a-b||_inf / max(||b||_inf, 1);
fp32 <= 1e-4 after 1000 steps."""
import os

import numpy as np
import pytest
import torch

from conftest import normwise
from oracle import fossen_np as O

pytestmark = pytest.mark.gpu

DT = 0.02
TOL64 = 1e-10
TOL32 = 1e-4


@pytest.fixture(scope="module")
def B():
    import bluerov2_dynamics_b200 as b
    return b


def cpu(t):
    return t.detach().cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------------------------------------ RHS
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
@pytest.mark.parametrize("tag", ["", "_cur"])
def test_rhs_against_reference(B, golden, dtype, tol, tag):
    cur = golden["rhs_current"] if tag else None
    e = B.Engine("thruster8", dtype, current=cur)
    lag = e.tensor(golden["rhs_thr_lag0"].reshape(-1, 24)).clone()
    xd = e.rhs(golden["rhs_thr_x"], golden["rhs_thr_u"], lag=lag, dt=DT)
    assert normwise(cpu(xd), golden[f"rhs_thr{tag}_xdot"]) < tol
    assert normwise(cpu(lag).reshape(-1, 8, 3), golden[f"rhs_thr{tag}_lag1"]) < tol
    e = B.Engine("wrench12", dtype, current=cur)
    assert normwise(cpu(e.rhs(golden["rhs_thr_x"], golden["rhs_w_tau"])), golden[f"rhs_w12{tag}_xdot"]) < tol
    e = B.Engine("quat13", dtype, current=cur)
    assert normwise(cpu(e.rhs(golden["rhs_q13_x"], golden["rhs_w_tau"])), golden[f"rhs_q13{tag}_xdot"]) < tol


def test_rhs_without_lag_buffer_is_fresh_lag(B, golden):
    e = B.Engine("thruster8", "f64")
    n = 32  # first half of the golden set starts from zero lag
    xd = e.rhs(golden["rhs_thr_x"][:n], golden["rhs_thr_u"][:n], lag=None, dt=DT)
    assert normwise(cpu(xd), golden["rhs_thr_xdot"][:n]) < 1e-12


def test_thruster_wrench_matches_oracle(B, golden):
    e = B.Engine("thruster8", "f64")
    u = golden["rhs_thr_u"]
    lag0 = golden["rhs_thr_lag0"]
    lag = e.tensor(lag0.reshape(-1, 24)).clone()
    tau = cpu(e.thruster_wrench(u, lag=lag, dt=DT))
    Ad, Bd = O.lag_zoh(DT)
    tau_o, lag_o = O.thruster_wrench(u, lag0, Ad, Bd, O.thruster_geometry()[2])
    assert normwise(tau, tau_o) < 1e-12
    assert normwise(cpu(lag).reshape(-1, 8, 3), lag_o) < 1e-12


def test_trig_range_reduction_large_angles(B):
    """yaw winds up without bound in long rollouts; the kernels' own sincos must stay accurate far from 0."""
    rng = np.random.default_rng(7)
    n = 4096
    x = np.zeros((n, 12))
    x[:, 3:6] = rng.uniform(-1, 1, (n, 3)) * np.array([1.0, 1.0, 1.0]) * 10.0 ** rng.uniform(-3, 4, (n, 1))
    x[:, 6:] = rng.uniform(-1, 1, (n, 6))
    tau = rng.uniform(-10, 10, (n, 6))
    ref = O.rhs_wrench12(x, tau, O.default_params())
    ok = np.abs(np.cos(x[:, 4])) > 1e-2  # away from the Euler-angle singularity, where the RHS is ill-conditioned
    got = cpu(B.Engine("wrench12", "f64").rhs(x, tau))
    assert normwise(got[ok], ref[ok]) < 1e-12
    x32 = x.astype(np.float32).astype(np.float64)  # same rounded angles for both sides
    ref32 = O.rhs_wrench12(x32, tau.astype(np.float32).astype(np.float64), O.default_params())
    got32 = cpu(B.Engine("wrench12", "f32").rhs(x32, tau))
    assert normwise(got32[ok], ref32[ok]) < 2e-5


# ------------------------------------------------------------------------------------------------ cfg 1
@pytest.mark.parametrize("dtype,tol", [("f64", TOL64), ("f32", TOL32)])
def test_cfg1_thruster_rk4_1000_steps(B, golden, dtype, tol):
    e = B.Engine("thruster8", dtype)
    x0 = golden["cfg1_x0"][None]
    U = np.tile(golden["cfg1_u_const"], (1000, 1))
    r = e.rollout(x0, U, dt=DT, integrator="rk4", stride=10)
    assert normwise(cpu(r.traj)[:, 0], golden["cfg1_const_traj_s10"][1:]) < tol
    assert normwise(cpu(r.lag).reshape(8, 3), golden["cfg1_const_lagT"]) < tol
    # same thing with one constant input row per vehicle
    r2 = e.rollout(x0, (golden["cfg1_u_const"][None], 1000), dt=DT, u_layout="const")
    assert torch.equal(r2.xT, r.xT)
    r = e.rollout(x0, golden["cfg1_U_var"], dt=DT, integrator="rk4", stride=10)
    assert normwise(cpu(r.traj)[:, 0], golden["cfg1_var_traj_s10"][1:]) < tol
    assert normwise(cpu(r.lag).reshape(8, 3), golden["cfg1_var_lagT"]) < tol


def test_cfg1_euler_dt001(B, golden):
    e = B.Engine("thruster8", "f64")
    U = np.tile(golden["cfg1_u_const"], (500, 1))
    r = e.rollout(golden["cfg1_x0"][None], U, dt=0.01, integrator="euler", stride=10)
    assert normwise(cpu(r.traj)[:, 0], golden["cfg1_euler_dt001_traj_s10"][1:]) < TOL64
    assert normwise(cpu(r.lag).reshape(8, 3), golden["cfg1_euler_dt001_lagT"]) < TOL64


def test_scipy_zoh_override_is_equivalent(B, golden):
    e = B.Engine("thruster8", "f64")
    e.set_lag_discrete(DT, golden["const_lag_Ad_0.02"], golden["const_lag_Bd_0.02"])
    r = e.rollout(golden["cfg1_x0"][None], golden["cfg1_U_var"], dt=DT, stride=10)
    assert normwise(cpu(r.traj)[:, 0], golden["cfg1_var_traj_s10"][1:]) < TOL64


# ------------------------------------------------------------------------------------------------ ensembles
ENS = [("thruster8", "rk4", "ens_x0", "ens_U8", "ens_thr_rk4_s20"),
       ("thruster8", "euler", "ens_x0", "ens_U8", "ens_thr_euler_s20"),
       ("wrench12", "rk4", "ens_x0", "ens_W6", "ens_w12_rk4_s20"),
       ("wrench12", "euler", "ens_x0", "ens_W6", "ens_w12_euler_s20"),
       ("quat13", "rk4", "ens_x0_q13", "ens_W6", "ens_q13_rk4_s20"),
       ("quat13", "euler", "ens_x0_q13", "ens_W6", "ens_q13_euler_s20")]


@pytest.mark.parametrize("dtype,tol", [("f64", TOL64), ("f32", TOL32)])
@pytest.mark.parametrize("kind,integ,x0k,uk,outk", ENS)
def test_ensembles_against_reference(B, golden, dtype, tol, kind, integ, x0k, uk, outk):
    e = B.Engine(kind, dtype)
    U = np.ascontiguousarray(np.transpose(golden[uk], (1, 0, 2)))
    r = e.rollout(golden[x0k], U, dt=DT, integrator=integ, stride=20)
    ref = np.transpose(golden[outk], (1, 0, 2))[1:]
    assert normwise(cpu(r.traj), ref) < tol
    assert normwise(cpu(r.xT), ref[-1]) < tol
    if kind == "thruster8":
        assert normwise(cpu(r.lag).reshape(-1, 8, 3), golden[f"ens_thr_{integ}_lagT"]) < tol


def _mc_phys(B, scales):
    ph = np.tile(B.default_physical(), (scales.shape[0], 1))
    from bluerov2_dynamics_b200 import _lib as L
    ph[:, L.PH_ADDED:L.PH_ADDED + 6] *= scales[:, :6]
    ph[:, L.PH_LIN:L.PH_LIN + 6] *= scales[:, 6:12]
    ph[:, L.PH_QUAD:L.PH_QUAD + 6] *= scales[:, 12:18]
    m = ph[:, L.PH_M:L.PH_M + 1]
    ph[:, L.PH_MINV:L.PH_MINV + 3] = 1.0 / (m - ph[:, L.PH_ADDED:L.PH_ADDED + 3])
    ph[:, L.PH_MINV + 3:L.PH_MINV + 6] = 1.0 / (ph[:, L.PH_I:L.PH_I + 3] - ph[:, L.PH_ADDED + 3:L.PH_ADDED + 6])
    return ph


@pytest.mark.parametrize("dtype,tol", [("f64", TOL64), ("f32", TOL32)])
def test_monte_carlo_vehicle_params(B, golden, dtype, tol):
    ph = _mc_phys(B, golden["mc_scales"])
    U = np.ascontiguousarray(np.transpose(golden["mc_W6"], (1, 0, 2)))
    for kind, x0k, outk in (("wrench12", "mc_x0", "mc_w12_rk4_xT"), ("quat13", "mc_x0_q13", "mc_q13_rk4_xT")):
        e = B.Engine(kind, dtype)
        e.set_vehicle_physical(ph)
        r = e.rollout(golden[x0k], U, dt=DT, integrator="rk4")
        assert normwise(cpu(r.xT), golden[outk]) < tol
        e.set_vehicle_physical(None)


# ------------------------------------------------------------------------------------------------ evaluators
def test_multistep_rmse_against_reference(B, golden):
    HS = [int(h) for h in golden["rmse_H"]]
    X, U8, W6, Xq = golden["rmse_X12"], golden["rmse_U8"], golden["rmse_W6"], golden["rmse_X13"]
    e = B.Engine("thruster8", "f64")
    assert np.allclose(e.multistep_rmse(X, U8, HS, dt=DT, integrator="rk4"), golden["rmse_thr_rk4_reset"], rtol=1e-10)
    assert np.allclose(e.multistep_rmse(X, U8, HS, dt=DT, integrator="euler"), golden["rmse_thr_euler_reset"], rtol=1e-10)
    assert np.isclose(e.multistep_rmse(X, U8, 1, dt=DT, integrator="euler"), golden["rmse_thr_onestep_reset"], rtol=1e-10)
    e = B.Engine("wrench12", "f64")
    assert np.allclose(e.multistep_rmse(X, W6, HS, dt=DT, integrator="euler"), golden["rmse_w12_euler"], rtol=1e-10)
    assert np.isclose(e.multistep_rmse(X, W6, 1, dt=DT, integrator="euler"), golden["rmse_w12_onestep"], rtol=1e-10)
    e = B.Engine("quat13", "f64")
    assert np.allclose(e.multistep_rmse(Xq, W6, HS, dt=DT, integrator="euler"), golden["rmse_q13_euler"], rtol=1e-10)
    assert np.isclose(e.multistep_rmse(Xq, W6, 1, dt=DT, integrator="euler"), golden["rmse_q13_onestep"], rtol=1e-10)
    # horizon by horizon == all horizons in one pass; too-short series -> NaN as the reference
    e = B.Engine("wrench12", "f64")
    one_pass = e.multistep_rmse(X, W6, HS, dt=DT, integrator="euler")
    assert one_pass == [e.multistep_rmse(X, W6, h, dt=DT, integrator="euler") for h in HS]
    assert np.isnan(e.multistep_rmse(X[:5], W6[:5], 10, dt=DT, integrator="euler"))
    # fp32 evaluator
    e = B.Engine("thruster8", "f32")
    assert np.allclose(e.multistep_rmse(X, U8, HS, dt=DT, integrator="rk4"), golden["rmse_thr_rk4_reset"], rtol=1e-4)


def test_multistep_rmse_lag_carry_via_lag0(B, golden):
    """The reference's literal semantics (one model object, lag state leaking across windows, trap T3) are
    reproduced by handing the evaluator the carried lag state of every window, computed by the oracle's
    sequential recurrence (it depends on the inputs only)."""
    X, U = golden["rmse_X12"], golden["rmse_U8"]
    H = 10
    ns = len(X) - H
    Ad, Bd = O.lag_zoh(DT)
    lag = np.zeros((1, 8, 3))
    lag0 = np.zeros((ns, 8, 3))
    for k in range(ns):
        lag0[k] = lag[0]
        for j in range(H):
            F = O.thrust_poly(U[k + j:k + j + 1])
            for _ in range(4):
                lag, _ = O.lag_step(lag, F, Ad, Bd)
    e = B.Engine("thruster8", "f64")
    se, cnt = e.multistep_se(X, U, [H], dt=DT, integrator="rk4", lag0=lag0.reshape(ns, 24), n_windows=ns)
    got = float(np.sqrt(se[0].item() / (cnt[0] * 12)))
    assert np.isclose(got, golden["rmse_thr_rk4_carry"][1], rtol=1e-10)


def test_multistep_rmse_lag_carry_matches_unmodified_reference(B, golden):
    """lag_mode="carry": the evaluator reproduces the numbers of the reference's OWN functions (one model object for
    all windows, trap T3): train_tank_brov2_rk4.multistep_rmse_endpoint_physics (RK4),
    train_tank_brov2_full_comparison.multistep_rmse_endpoint_physics (Euler),
    train_tank_brov2_koopmanEDMDc.one_step_rmse_physics."""
    X, U = golden["rmse_X12"], golden["rmse_U8"]
    HS = [int(h) for h in golden["rmse_H"]]
    e = B.Engine("thruster8", "f64")
    assert 30 < e.carry_steps(DT, "rk4") < 80 and 120 < e.carry_steps(DT, "euler") < 320
    got = e.multistep_rmse(X, U, HS, dt=DT, integrator="rk4", lag_mode="carry")
    assert np.allclose(got, golden["rmse_thr_rk4_carry"], rtol=1e-10), (got, golden["rmse_thr_rk4_carry"])
    got = e.multistep_rmse(X, U, HS, dt=DT, integrator="euler", lag_mode="carry")
    assert np.allclose(got, golden["rmse_thr_euler_carry"], rtol=1e-10)
    got = e.multistep_rmse(X, U, 1, dt=DT, integrator="euler", lag_mode="carry")
    assert np.isclose(got, golden["rmse_thr_onestep_carry"], rtol=1e-10)
    # the carried and the reset semantics really differ on this series
    assert not np.allclose(golden["rmse_thr_rk4_carry"], golden["rmse_thr_rk4_reset"], rtol=1e-6)
    # a shard of the windows (global window / row offsets, history rows before the first window) adds up to the whole
    H = 10
    ns = len(X) - H
    se_all, _ = e.multistep_se(X, U, [H], dt=DT, integrator="rk4", lag_mode="carry")
    depth = e.carry_steps(DT, "rk4")
    w0 = 57
    r0 = max(0, w0 - depth - 1)
    se_a, ca = e.multistep_se(X[:w0 + H], U[:w0 + H], [H], dt=DT, integrator="rk4", lag_mode="carry", n_windows=w0)
    se_b, cb = e.multistep_se(X[r0:], U[r0:], [H], dt=DT, integrator="rk4", lag_mode="carry", n_windows=ns - w0,
                              window0=w0, row0=r0)
    assert ca[0] + cb[0] == ns
    assert np.isclose(se_a[0].item() + se_b[0].item(), se_all[0].item(), rtol=1e-12)
    # wrench models have no hidden state: carry == reset
    w = B.Engine("wrench12", "f64")
    assert w.multistep_rmse(X, golden["rmse_W6"], HS, dt=DT, integrator="euler", lag_mode="carry") == \
        w.multistep_rmse(X, golden["rmse_W6"], HS, dt=DT, integrator="euler")


def test_sim_data_generator(B, golden):
    """training/train_sim_brov2_koopmanEDMDc.py:179-182: Euler rollout, dt = 0.05, 1500 stored inputs."""
    e = B.Engine("thruster8", "f64")
    r = e.rollout(np.zeros((1, 12)), golden["simgen_inputs"], dt=0.05, integrator="euler", stride=1)
    assert normwise(cpu(r.traj)[::10, 0], golden["simgen_states_true_s10"]) < TOL64


# ------------------------------------------------------------------------------------------------ reduced model
def test_reduced9_against_reference(B, golden):
    x, u = golden["red9_x"], golden["red9_u"]
    out = B.reduced9_rhs(torch.tensor(x, device="cuda"), torch.tensor(u, device="cuda"))
    assert normwise(cpu(out), golden["red9_xdot_f64"]) < 1e-14
    out = B.reduced9_rhs(torch.tensor(x, device="cuda", dtype=torch.float32), torch.tensor(u, device="cuda", dtype=torch.float32))
    assert out.dtype == torch.float32
    assert normwise(cpu(out), golden["red9_xdot_f32"]) < 2e-6
    # ragged size (not a multiple of the block) and an unaligned view
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((1000 + 1, 9))
    ub = rng.standard_normal((1000 + 1, 4))
    got = B.reduced9_rhs(torch.tensor(xb, device="cuda")[1:], torch.tensor(ub, device="cuda")[1:])
    assert normwise(cpu(got), O.rhs_reduced9(xb[1:], ub[1:])) < 1e-14


# ------------------------------------------------------------------------------------------------ oracle, larger
@pytest.mark.parametrize("kind,nu,scale", [("thruster8", 8, 1.0), ("wrench12", 6, O.np.array([40, 40, 40, 5, 5, 5.0])),
                                           ("quat13", 6, O.np.array([40, 40, 40, 5, 5, 5.0]))])
def test_fp64_against_oracle_4096_vehicles(B, kind, nu, scale):
    rng = np.random.default_rng(11)
    n, T = 4096, 150
    x0 = np.zeros((n, 12))
    x0[:, :3] = rng.uniform(-2, 2, (n, 3))
    x0[:, 3:5] = rng.uniform(-0.2, 0.2, (n, 2))
    x0[:, 5] = rng.uniform(-np.pi, np.pi, n)
    U = O.smooth_inputs(rng, T, nu, n=n, scale=scale, sigma=0.05)
    if kind == "quat13":
        q = O.euler_to_quat(x0[:, 3], x0[:, 4], x0[:, 5])
        x0 = np.concatenate([x0[:, :3], q, x0[:, 6:]], axis=1)
    m = O.Model(kind, DT)
    snaps, xT, lagT = O.rollout(m, "rk4", x0, U, stride=1)
    r = B.Engine(kind, "f64").rollout(x0, U, dt=DT, integrator="rk4", stride=1)
    # Hard inputs pitch a few vehicles through theta = +-pi/2, where the Euler-angle kinematics are singular and a
    # 1e-15 perturbation of x0 grows to 1e-4 in the REFERENCE itself; parity is asserted on the well-conditioned rest.
    ok = np.ones(n, bool) if kind == "quat13" else np.abs(np.cos(snaps[:, :, 4])).min(axis=0) > 0.2
    assert ok.mean() > 0.8
    assert normwise(cpu(r.traj)[:, ok], snaps[:, ok]) < TOL64
    assert normwise(cpu(r.xT)[ok], xT[ok]) < TOL64
    if kind == "thruster8":
        assert normwise(cpu(r.lag).reshape(n, 8, 3), lagT) < TOL64


def test_fp32_tolerance_after_1000_steps(B):
    """north_star: fp32 kernel within 1e-4 relative of the float64 reference algorithm after 1000 steps."""
    rng = np.random.default_rng(12)
    n, T = 256, 1000
    x0 = np.zeros((n, 12))
    x0[:, :3] = rng.uniform(-2, 2, (n, 3))
    x0[:, 3:5] = rng.uniform(-0.2, 0.2, (n, 2))
    x0[:, 5] = rng.uniform(-np.pi, np.pi, n)
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.02).astype(np.float32)  # the reference generator's sigma
    snaps, xT, _ = O.rollout(O.Model("thruster8", DT), "rk4", x0.astype(np.float32).astype(np.float64),
                             U.astype(np.float64), stride=1)
    ok = np.abs(np.cos(snaps[:, :, 4])).min(axis=0) > 0.2  # see test_fp64_against_oracle_4096_vehicles
    assert ok.mean() > 0.9
    r = B.Engine("thruster8", "f32").rollout(x0, U, dt=DT, integrator="rk4")
    assert normwise(cpu(r.xT)[ok], xT[ok]) < TOL32


def test_wrench_lag1_extension_against_oracle(B):
    """First-order wrench lag (north-star extension; no reference counterpart -> parity unpinned, oracle only)."""
    rng = np.random.default_rng(13)
    n, T = 64, 200
    x0 = np.zeros((n, 12))
    x0[:, 2] = 1.0
    x0[:, 5] = rng.uniform(-3, 3, n)
    scale = np.array([40, 40, 40, 5, 5, 5.0])
    U = O.smooth_inputs(rng, T, 6, n=n, scale=scale, sigma=0.05)
    for integ in ("rk4", "euler"):
        m = O.Model("wrench12", DT, lag1_T=0.15)
        xa = np.concatenate([x0, np.zeros((n, 6))], axis=1)
        _, xT, _ = O.rollout(m, integ, xa, U)
        e = B.Engine("wrench12", "f64")
        e.set_wrench_lag1(True, T_lag=0.15)
        r = e.rollout(x0, U, dt=DT, integrator=integ)
        assert normwise(cpu(r.xT), xT[:, :12]) < TOL64
        assert normwise(cpu(r.lag), xT[:, 12:]) < TOL64


# ------------------------------------------------------------------------------------------------ invariances
def test_chunking_strides_and_host_path_are_bit_identical(B):
    rng = np.random.default_rng(14)
    n, T = 333, 64  # ragged: partial warp, partial block
    x0 = rng.uniform(-1, 1, (n, 12)) * 0.3
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.05)
    for dtype in ("f64", "f32"):
        e = B.Engine("thruster8", dtype)
        full = e.rollout(x0, U, dt=DT, stride=4)
        fullp = e.rollout(x0, U, dt=DT, stride=4, lag_repr="projected")
        assert torch.equal(fullp.xT, full.xT) and torch.equal(fullp.traj, full.traj)     # one kernel behind both
        # chunked: 3 launches, carrying state and the kernel's own (allocation-projected) lag: same bits
        x, lag = e.tensor(x0), None
        parts = []
        for a, b in ((0, 10), (10, 37), (37, 64)):
            r = e.rollout(x, U[a:b], dt=DT, stride=4, lag0=lag, step0=a, lag_repr="projected")
            x, lag = r.xT, r.lag
            parts.append(r.traj)
        assert torch.equal(x, full.xT) and torch.equal(lag, fullp.lag)
        assert torch.equal(torch.cat(parts), full.traj)
        # carrying the per-thruster states instead re-projects them at every chunk start: equal to rounding
        x, lag = e.tensor(x0), None
        for a, b in ((0, 10), (10, 37), (37, 64)):
            r = e.rollout(x, U[a:b], dt=DT, lag0=lag, step0=a)
            x, lag = r.xT, r.lag
        rt = 1e-13 if dtype == "f64" else 2e-5
        assert normwise(cpu(x), cpu(full.xT)) < rt and normwise(cpu(lag), cpu(full.lag)) < rt
        # host-buffer path (pageable and pinned), tiny chunks so the double buffering cycles
        x0h = x0.astype(e.ndtype)
        Uh = np.ascontiguousarray(U.astype(e.ndtype))
        xT, lagT, traj = e.rollout_host(x0h, Uh, dt=DT, stride=4, chunk_steps=8)
        assert np.array_equal(xT, full.xT.cpu().numpy()) and normwise(lagT, cpu(full.lag)) < rt
        assert np.array_equal(traj, full.traj.cpu().numpy())
        Up = B.pinned_empty(Uh.shape, e.ndtype)
        Up[...] = Uh
        xT2, _, traj2 = e.rollout_host(x0h, Up, dt=DT, stride=4, chunk_steps=0)
        assert np.array_equal(xT2, xT) and np.array_equal(traj2, traj)
        # a shard of the ensemble gives the same rows as the whole (no cross-vehicle coupling)
        half = e.rollout(x0[100:200], np.ascontiguousarray(U[:, 100:200]), dt=DT)
        assert torch.equal(half.xT, full.xT[100:200])


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-11), ("f32", 5e-5)])
def test_projected_lag_representation(B, golden, dtype, tol):
    """Allocation-projected lag state (18 values) == per-thruster lag state (24 values) up to rounding: same
    trajectories, Z_out = alloc . lag_out, chunks carry Z, un-projection is refused."""
    rng = np.random.default_rng(16)
    n, T = 200, 90
    x0 = rng.uniform(-1, 1, (n, 12)) * 0.1
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.02)
    lag0 = rng.uniform(-0.05, 0.05, (n, 24))
    e = B.Engine("thruster8", dtype)
    thr = e.rollout(x0, U, dt=DT, lag0=lag0, stride=30)
    z0 = e.project_lag(lag0)
    prj = e.rollout(x0, U, dt=DT, lag0=z0, stride=30, lag_repr="projected")
    assert prj.lag.shape == (n, 18)
    assert normwise(cpu(prj.xT), cpu(thr.xT)) < tol and normwise(cpu(prj.traj), cpu(thr.traj)) < tol
    assert normwise(cpu(prj.lag), cpu(e.project_lag(thr.lag))) < tol
    # no lag wanted back: thruster-coordinate lag_in is projected inside the kernel
    nol = e.rollout(x0, U, dt=DT, lag0=lag0, want_lag=False)
    assert nol.lag is None and normwise(cpu(nol.xT), cpu(thr.xT)) < tol
    # chunked carry of Z is bit-identical to one launch
    a = e.rollout(x0, U[:40], dt=DT, lag0=z0, lag_repr="projected")
    b = e.rollout(a.xT, U[40:], dt=DT, lag0=a.lag, lag_repr="projected", step0=40)
    assert torch.equal(b.xT, prj.xT) and torch.equal(b.lag, prj.lag)
    # host path with the projected carry
    x0h, Uh = x0.astype(e.ndtype), np.ascontiguousarray(U.astype(e.ndtype))
    xT, lagT, _ = e.rollout_host(x0h, Uh, dt=DT, lag0=cpu(z0).astype(e.ndtype), chunk_steps=16, lag_repr="projected")
    assert np.array_equal(xT, prj.xT.cpu().numpy()) and np.array_equal(lagT, prj.lag.cpu().numpy())
    xT2, lag2, _ = e.rollout_host(x0h, Uh, dt=DT, lag0=lag0.astype(e.ndtype), chunk_steps=16, want_lag=False)
    assert lag2 is None and normwise(xT2, cpu(thr.xT)) < tol
    # against the reference's golden ensemble too
    Ug = np.ascontiguousarray(np.transpose(golden["ens_U8"], (1, 0, 2)))
    r = e.rollout(golden["ens_x0"], Ug, dt=DT, stride=20, lag_repr="projected")
    assert normwise(cpu(r.traj), np.transpose(golden["ens_thr_rk4_s20"], (1, 0, 2))[1:]) < (TOL64 if dtype == "f64" else TOL32)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_temporal_tiling_is_bit_identical(B, dtype):
    """time_slices (ticketed time slices per vehicle block, hand-over through xT / lag_out) never changes a bit."""
    rng = np.random.default_rng(17)
    n, T = 5000, 53   # ragged: partial last block, steps not divisible by the slice counts
    x0 = rng.uniform(-1, 1, (n, 12)) * 0.2
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.03)
    e = B.Engine("thruster8", dtype)
    for repr_ in ("thruster", "projected"):
        ref = e.rollout(x0, U, dt=DT, stride=7, time_slices=1, lag_repr=repr_)
        for q in (2, 3, 4, 7, 53):
            r = e.rollout(x0, U, dt=DT, stride=7, time_slices=q, lag_repr=repr_)
            assert torch.equal(r.xT, ref.xT) and torch.equal(r.lag, ref.lag) and torch.equal(r.traj, ref.traj), (repr_, q)
    # in place (x0 aliases xT, lag_in aliases lag_out), as the chunked bench loop runs it
    x = e.tensor(x0).clone()
    lag = torch.zeros((n, 18), device="cuda", dtype=e.tdtype)
    e.rollout(x, U, dt=DT, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected", time_slices=4)
    ref = e.rollout(x0, U, dt=DT, lag_repr="projected", time_slices=1)
    assert torch.equal(x, ref.xT) and torch.equal(lag, ref.lag)
    # wrench model (no lag buffers) and the automatic choice at BASELINE cfg2 width
    w = B.Engine("quat13", dtype)
    xq = np.zeros((n, 13)); xq[:, 3] = 1.0
    W = O.smooth_inputs(rng, T, 6, n=n, scale=5.0)
    a, b = w.rollout(xq, W, dt=DT, time_slices=1), w.rollout(xq, W, dt=DT, time_slices=5)
    assert torch.equal(a.xT, b.xT)
    if dtype == "f64":
        n2 = 65536
        g = torch.Generator(device="cuda").manual_seed(9)
        U2 = (torch.rand((32, n2, 8), device="cuda", dtype=torch.float64, generator=g) * 0.8 - 0.4)
        x2 = torch.zeros((n2, 12), device="cuda", dtype=torch.float64)
        a, b = e.rollout(x2, U2, dt=DT, time_slices=0), e.rollout(x2, U2, dt=DT, time_slices=1)
        assert torch.equal(a.xT, b.xT) and torch.equal(a.lag, b.lag)


def test_quat13_odd_n_unaligned_snapshots(B):
    rng = np.random.default_rng(15)
    n, T = 37, 12
    x0 = np.zeros((n, 13))
    x0[:, 3] = 1.0
    x0[:, :3] = rng.uniform(-1, 1, (n, 3))
    U = O.smooth_inputs(rng, T, 6, n=n, scale=10.0, sigma=0.1)
    snaps, xT, _ = O.rollout(O.Model("quat13", DT), "euler", x0, U, stride=1)
    for dtype, tol in (("f64", TOL64), ("f32", TOL32)):
        r = B.Engine("quat13", dtype).rollout(x0, U, dt=DT, integrator="euler", stride=1)
        assert normwise(cpu(r.traj), snaps) < tol


def test_edge_cases(B):
    e = B.Engine("wrench12", "f64")
    x0 = np.zeros((1, 12))
    r = e.rollout(x0, np.zeros((0, 6)), dt=DT)  # zero steps: identity
    assert torch.equal(r.xT, e.tensor(x0))
    r = e.rollout(np.zeros((0, 12)), np.zeros((5, 0, 6)), dt=DT)  # empty ensemble
    assert r.xT.shape == (0, 12)
    with pytest.raises(ValueError):
        e.rollout(np.zeros((2, 11)), np.zeros((5, 6)), dt=DT)
    with pytest.raises(ValueError):
        e.rollout(np.zeros((2, 12)), np.zeros((5, 3, 6)), dt=DT)
    with pytest.raises(B.BrovError):
        e.rollout(np.zeros((2, 12)), np.zeros((5, 6)), dt=-1.0)
    with pytest.raises(B.BrovError):
        B.Engine("thruster8", "f64").set_wrench_lag1(True)
    with pytest.raises(ValueError):
        e.multistep_se(np.zeros((10, 12)), np.zeros((10, 6)), [10, 1])


def test_full_size_cfg2_properties(B):
    """BASELINE config 2 at full width (65,536 vehicles, fp64, thruster model), short horizon: determinism, chunking
    invariance and agreement with the oracle on a random subset of vehicles."""
    n, T = 65536, 40
    g = torch.Generator(device="cuda").manual_seed(2)
    e = B.Engine("thruster8", "f64")
    x0 = torch.zeros((n, 12), device="cuda", dtype=torch.float64)
    x0[:, :3] = torch.rand((n, 3), device="cuda", dtype=torch.float64, generator=g) * 4 - 2
    x0[:, 5] = torch.rand(n, device="cuda", dtype=torch.float64, generator=g) * 6 - 3
    U = (torch.rand((T, n, 8), device="cuda", dtype=torch.float64, generator=g) * 0.8 - 0.4).contiguous()
    a = e.rollout(x0, U, dt=DT)
    b = e.rollout(x0, U, dt=DT)
    assert torch.equal(a.xT, b.xT) and torch.equal(a.lag, b.lag)
    m1 = e.rollout(x0, U[:17], dt=DT, lag_repr="projected")
    m2 = e.rollout(m1.xT, U[17:], dt=DT, lag0=m1.lag, step0=17, lag_repr="projected")
    assert torch.equal(m2.xT, a.xT)
    idx = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(3))
    _, xT, _ = O.rollout(O.Model("thruster8", DT), "rk4", cpu(x0[idx]), cpu(U[:, idx]))
    assert normwise(cpu(a.xT[idx]), xT) < TOL64


def test_full_size_cfg3_properties(B):
    """BASELINE config 3 at full width (1,048,576 vehicles, fp32, stride-10 writeback), short horizon."""
    n, T = 1 << 20, 20
    g = torch.Generator(device="cuda").manual_seed(4)
    e = B.Engine("thruster8", "f32")
    x0 = torch.zeros((n, 12), device="cuda", dtype=torch.float32)
    x0[:, 5] = torch.rand(n, device="cuda", generator=g) * 6 - 3
    U = (torch.rand((T, n, 8), device="cuda", generator=g) * 0.8 - 0.4).contiguous()
    r = e.rollout(x0, U, dt=DT, stride=10)
    assert r.traj.shape == (2, n, 12)
    assert torch.equal(r.traj[1], r.xT)
    mid = e.rollout(x0, U[:10], dt=DT)
    assert torch.equal(r.traj[0], mid.xT)
    assert bool(torch.isfinite(r.xT).all())
    idx = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(5))
    _, xT, _ = O.rollout(O.Model("thruster8", DT), "rk4", cpu(x0[idx]), cpu(U[:, idx]))
    assert normwise(cpu(r.xT[idx]), xT) < TOL32


# ------------------------------------------------------------------------------------------------ reference API mirror
def test_fossen_mirror_classes(B, golden):
    from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2, ThrusterLag
    from bluerov2_dynamics_b200.fossen.BlueROV2_thrust import BlueROV2 as W12
    from bluerov2_dynamics_b200.fossen import BlueROV2_wrench as QW
    from bluerov2_dynamics_b200.evaluators import simulate_physics, multistep_rmse_endpoint_physics, one_step_rmse_physics

    rov = BlueROV2(dt=0.02)
    assert np.allclose(np.diag(rov.Minv), golden["const_Minv_diag"], rtol=1e-15)
    assert np.allclose(np.stack([t["r"] for t in rov.thrusters_r]), golden["const_thr_r"], atol=1e-16)
    # dynamics() is stateful: the second call sees the lag state left by the first
    x, u = golden["rhs_thr_x"][0], golden["rhs_thr_u"][0]
    xd1 = rov.dynamics(x, u, DT)
    assert normwise(xd1, golden["rhs_thr_xdot"][0]) < 1e-12
    assert normwise(np.stack([l._x for l in rov.thruster_lags]), golden["rhs_thr_lag1"][0]) < 1e-12
    xd2 = rov.dynamics(x, u, DT)
    assert not np.allclose(xd1, xd2)
    # simulate_physics: RK4 loop of train_tank_brov2_rk4.py, trajectory incl. row 0, lag state left in rov
    rov = BlueROV2()
    traj = simulate_physics(golden["cfg1_x0"], golden["cfg1_U_var"], DT, rov, integrator="rk4")
    assert traj.shape == (1001, 12) and np.array_equal(traj[0], golden["cfg1_x0"])
    assert normwise(traj[::10], golden["cfg1_var_traj_s10"]) < TOL64
    assert normwise(np.stack([l._x for l in rov.thruster_lags]), golden["cfg1_var_lagT"]) < TOL64
    # wrench models: shape validation by reshape (ValueError) as the reference
    w = W12()
    assert normwise(w.dynamics(golden["rhs_thr_x"][3], golden["rhs_w_tau"][3]), golden["rhs_w12_xdot"][3]) < 1e-12
    with pytest.raises(ValueError):
        w.dynamics(np.zeros(11), np.zeros(6))
    q = QW.BlueROV2(current_speed=golden["rhs_current"])
    assert normwise(q.dynamics(golden["rhs_q13_x"][40], golden["rhs_w_tau"][40]), golden["rhs_q13_cur_xdot"][40]) < 1e-12
    # attribute edits after construction follow the reference's semantics (trap T5): C/D change, Minv does not
    w2 = W12()
    w2.Xu_abs *= 2.0
    ref = O.default_params()
    ref["Xu_abs"] = ref["Xu_abs"] * 2.0
    assert normwise(w2.dynamics(golden["rhs_thr_x"][5], golden["rhs_w_tau"][5]),
                    O.rhs_wrench12(golden["rhs_thr_x"][5:6], golden["rhs_w_tau"][5:6], ref)[0]) < 1e-12
    # evaluators
    HS = [int(h) for h in golden["rmse_H"]]
    got = multistep_rmse_endpoint_physics(golden["rmse_X12"], golden["rmse_U8"], HS, DT, integrator="rk4")
    assert np.allclose(got, golden["rmse_thr_rk4_carry"], rtol=1e-10)  # default = the reference's literal semantics
    got = multistep_rmse_endpoint_physics(golden["rmse_X12"], golden["rmse_U8"], HS, DT, integrator="rk4", lag_mode="reset")
    assert np.allclose(got, golden["rmse_thr_rk4_reset"], rtol=1e-10)
    assert np.isclose(one_step_rmse_physics(golden["rmse_X12"], golden["rmse_U8"], DT), golden["rmse_thr_onestep_carry"], rtol=1e-10)
    assert np.isclose(one_step_rmse_physics(golden["rmse_X12"], golden["rmse_W6"], DT, model="wrench12"),
                      golden["rmse_w12_onestep"], rtol=1e-10)
    # helpers
    assert np.allclose(QW.quat_to_rotation_matrix(golden["quat_q"][0]), golden["quat_to_R"][0], atol=1e-15)
    assert np.allclose(QW.quat_multiply(golden["quat_q"][1], golden["quat_q2"][1]), golden["quat_multiply"][1], atol=1e-15)
    assert np.allclose(QW.euler_to_quat(*golden["quat_euler_in"][2]), golden["quat_euler_to_quat"][2], atol=1e-15)
    lag = ThrusterLag()
    y = [lag.step(1.0, DT) for _ in range(3)]
    Ad, Bd = O.lag_zoh(DT)
    xs = np.zeros(3)
    for k in range(3):
        xs = Ad @ xs + Bd
    assert np.isclose(y[-1], O.LAG_CC @ xs, rtol=1e-13)


def test_bluerov_torch_mirror(B, golden):
    from bluerov2_dynamics_b200.fossen.bluerov_torch import bluerov_compute, ssa
    x = torch.tensor(golden["red9_x"])
    u = torch.tensor(golden["red9_u"])
    out = bluerov_compute(0.0, x, u)  # CPU tensors in -> CPU tensor out, computed on the GPU
    assert out.device.type == "cpu" and out.dtype == torch.float64
    assert normwise(out.numpy(), golden["red9_xdot_f64"]) < 1e-14
    one = bluerov_compute(0.0, x[0].cuda().float(), u[0].cuda().float())  # 1-D promoted to a batch of one
    assert one.shape == (1, 9) and one.is_cuda and one.dtype == torch.float32
    assert np.allclose(ssa(torch.tensor(golden["ssa_in"])).numpy(), golden["ssa_out"], atol=1e-14)


def test_monte_carlo_with_trajectory_fp64_shared_memory_opt_in(B, golden):
    """fp64 per-vehicle table (36 KB) + snapshot tiles (12 KB) + static shared memory exceed the 48 KB default:
    the launcher has to opt in (found by profiles/sanitize_smoke.py)."""
    rng = np.random.default_rng(5)
    n, T = 300, 12
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 9:27] *= rng.uniform(0.8, 1.2, (n, 18))
    ph[:, 27:30] = 1.0 / (ph[:, 0:1] - ph[:, 9:12])
    ph[:, 30:33] = 1.0 / (ph[:, 6:9] - ph[:, 12:15])
    x0 = rng.uniform(-0.3, 0.3, (n, 12))
    U = rng.uniform(-5, 5, (T, n, 6))
    p = O.default_params()
    names = ["Xu_dot", "Yv_dot", "Zw_dot", "Kp_dot", "Mq_dot", "Nr_dot", "Xu", "Yv", "Zw", "Kp", "Mq", "Nr",
             "Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs"]
    for j, k in enumerate(names):
        p[k] = ph[:, 9 + j]
    p["Minv"] = O.minv_diag(p)
    snaps, xT, _ = O.rollout(O.Model("wrench12", DT, p), "rk4", x0, U, stride=4)
    e = B.Engine("wrench12", "f64")
    e.set_vehicle_physical(ph)
    r = e.rollout(x0, U, dt=DT, stride=4)
    assert normwise(cpu(r.xT), xT) < TOL64 and normwise(cpu(r.traj), snaps) < TOL64


# ------------------------------------------------------------------------------------ full size, EVERY vehicle checked
def _c_oracle():
    import os
    import subprocess
    from conftest import ROOT
    from oracle import c_oracle
    if not c_oracle.available():
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    return c_oracle


def test_full_size_cfg2_every_vehicle_against_c_oracle(B):
    """BASELINE config 2 at full width: all 65,536 fp64 vehicles x 1000 RK4 steps (10 chunks of 100 steps carrying
    state and per-thruster lag) against the plain-C oracle, which runs the reference's sequential lag steps on every
    host core.  1e-10 normwise on the final state AND on the hidden lag state of every vehicle."""
    CO = _c_oracle()
    n, chunks, T = 65536, 10, 100
    g = torch.Generator(device="cuda").manual_seed(11)
    e = B.Engine("thruster8", "f64")
    x = torch.zeros((n, 12), device="cuda", dtype=torch.float64)
    x[:, :3] = torch.rand((n, 3), device="cuda", dtype=torch.float64, generator=g) * 4 - 2
    x[:, 3:5] = torch.rand((n, 2), device="cuda", dtype=torch.float64, generator=g) * 0.2 - 0.1
    x[:, 5] = torch.rand(n, device="cuda", dtype=torch.float64, generator=g) * 6 - 3
    lag = torch.zeros((n, 24), device="cuda", dtype=torch.float64)
    xo, lo = cpu(x), np.zeros((n, 8, 3))
    worst = 0.0
    for c in range(chunks):
        U = (torch.rand((T, n, 8), device="cuda", dtype=torch.float64, generator=g) * 0.8 - 0.4).contiguous()
        r = e.rollout(x, U, dt=DT, lag0=lag, step0=c * T)
        x, lag = r.xT, r.lag
        _, xo, lo = CO.rollout("thruster8", "rk4", DT, xo, cpu(U), lag0=lo)
        worst = max(worst, normwise(cpu(x), xo), normwise(cpu(lag).reshape(n, 8, 3), lo))
    assert worst < TOL64, worst


def test_full_size_cfg3_every_vehicle_against_c_oracle(B):
    """BASELINE config 3 at full width: all 1,048,576 fp32 vehicles x 100 RK4 steps with stride-10 snapshots against
    the float64 C oracle fed the same (float32-representable) inputs; 1e-4 normwise on every snapshot."""
    CO = _c_oracle()
    n, T = 1 << 20, 100
    g = torch.Generator(device="cuda").manual_seed(12)
    e = B.Engine("thruster8", "f32")
    x0 = torch.zeros((n, 12), device="cuda", dtype=torch.float32)
    x0[:, :3] = torch.rand((n, 3), device="cuda", generator=g) * 4 - 2
    x0[:, 5] = torch.rand(n, device="cuda", generator=g) * 6 - 3
    U = (torch.rand((T, n, 8), device="cuda", generator=g) * 0.8 - 0.4).contiguous()
    r = e.rollout(x0, U, dt=DT, stride=10)
    snaps, xT, _ = CO.rollout("thruster8", "rk4", DT, cpu(x0), cpu(U), stride=10)
    assert normwise(cpu(r.traj), snaps) < TOL32 and normwise(cpu(r.xT), xT) < TOL32


def test_install_as_fossen_serves_unmodified_reference_imports(golden):
    """The import lines and the loop of the reference's fossen/test_euler.py, run against the engine in a fresh
    interpreter: `import fossen...` / `import Koopman...` resolve to the mirrors, and ten stateful Euler steps through
    `rov.dynamics` land on the reference's trajectory."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    code = """
import json, sys
import numpy as np
sys.path.insert(0, %r)
import bluerov2_dynamics_b200 as brov
brov.install_as_fossen()
from fossen.BlueROV2 import BlueROV2                      # fossen/test_euler.py:2 (flat import there)
from fossen.BlueROV2_thrust import BlueROV2 as Wrench12
from fossen.BlueROV2_wrench import BlueROV2 as Quat13, euler_to_quat
from fossen.bluerov_torch import bluerov_compute, ssa
from Koopman.koopmanEDMDc import KoopmanEDMDc
import fossen.parameters as P
assert BlueROV2.__module__.startswith("bluerov2_dynamics_b200") and P.m == 11.4
rov = BlueROV2(dt=0.01)
x = np.zeros(12); x[2] = 5.0
u = np.array([0.1, 0.1, 0.1, 0.0, 0.5, 0.5, 0.5, 0.5])
for _ in range(10):
    x = x + 0.01 * rov.dynamics(x, u, 0.01)
print(json.dumps(x.tolist()))
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    x = np.array(json.loads(out.stdout.strip().splitlines()[-1]))
    assert normwise(x, golden["cfg1_euler_dt001_traj_s10"][1]) < TOL64


def test_step_entry_point_matches_rollout(B, golden):
    """brov_step = one iteration of the reference's simulate_physics loop; chaining it reproduces the rollout."""
    e = B.Engine("thruster8", "f64")
    x0, U = golden["ens_x0"], np.transpose(golden["ens_U8"], (1, 0, 2))[:25]
    ref = e.rollout(x0, U, dt=DT)
    x = e.tensor(x0)
    lag = torch.zeros((x0.shape[0], 24), device="cuda", dtype=torch.float64)
    for k in range(U.shape[0]):
        r = e.step(x, U[k], lag=lag, dt=DT)
        x = r.xT
    # every step re-projects the per-thruster lag it is handed: equal to the one-launch rollout to rounding
    assert normwise(cpu(x), cpu(ref.xT)) < 1e-13 and normwise(cpu(lag), cpu(ref.lag)) < 1e-13
    q = B.Engine("quat13", "f32")
    xq = golden["ens_x0_q13"]
    one = q.step(xq, golden["ens_W6"][:, 0], dt=DT, integrator="euler")
    assert torch.equal(one.xT, q.rollout(xq, golden["ens_W6"][:, :1].transpose(1, 0, 2), dt=DT, integrator="euler").xT)


@pytest.mark.parametrize("kind", ["wrench12", "quat13"])
def test_cfg4_monte_carlo_sweep_with_wrench_lag(B, kind):
    """BASELINE config 4 as a whole: per-vehicle added-mass / damping perturbations U(0.7, 1.3) (Minv rebuilt, trap T5)
    AND per-vehicle first-order wrench lag T_lag ~ U(0.05, 0.3) s, wrench inputs scaled (40,40,40,5,5,5), both wrench
    models, fp64 and fp32, against the oracle (the lag is an extension: parity unpinned by the reference)."""
    rng = np.random.default_rng(21)
    n, T = 777, 150
    nx = 13 if kind == "quat13" else 12
    x0 = np.zeros((n, nx))
    x0[:, :3] = rng.uniform(-1, 1, (n, 3))
    if kind == "quat13":
        q = rng.normal(size=(n, 4))
        x0[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    else:
        x0[:, 3:6] = rng.uniform(-0.3, 0.3, (n, 3))
    U = O.smooth_inputs(rng, T, 6, n=n, scale=np.array([40, 40, 40, 5, 5, 5.0]), sigma=0.05)
    scales = rng.uniform(0.7, 1.3, (n, 18))
    Tlag = rng.uniform(0.05, 0.3, n)
    p = O.default_params()
    names = ["Xu_dot", "Yv_dot", "Zw_dot", "Kp_dot", "Mq_dot", "Nr_dot", "Xu", "Yv", "Zw", "Kp", "Mq", "Nr",
             "Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs"]
    ph = np.tile(B.default_physical(), (n, 1))
    for j, k in enumerate(names):
        p[k] = p[k] * scales[:, j]
        ph[:, 9 + j] = p[k]
    p["Minv"] = O.minv_diag(p)
    ph[:, 27:33] = p["Minv"]
    ph[:, 36] = Tlag
    xa = np.concatenate([x0, np.zeros((n, 6))], axis=1)
    snaps, xT, _ = O.rollout(O.Model(kind, DT, p, lag1_T=Tlag), "rk4", xa, U, stride=50)
    for dtype, tol in (("f64", TOL64), ("f32", TOL32)):
        e = B.Engine(kind, dtype)
        e.set_wrench_lag1(True)
        e.set_vehicle_physical(ph)
        r = e.rollout(x0, U, dt=DT, stride=50)
        assert normwise(cpu(r.xT), xT[:, :nx]) < tol
        assert normwise(cpu(r.lag), xT[:, nx:]) < tol * 40      # filtered wrench, up to 40 N
        assert normwise(cpu(r.traj), snaps[:, :, :nx]) < tol


def test_monte_carlo_per_vehicle_current(B):
    """Per-vehicle ocean currents in the Monte-Carlo table (the relative-velocity terms are only compiled in when the
    caller flags a non-zero current) against the oracle; and a table without currents equals the shared-constant run."""
    rng = np.random.default_rng(8)
    n, T = 500, 60
    x0 = rng.uniform(-0.4, 0.4, (n, 12))
    U = rng.uniform(-10, 10, (T, n, 6))
    cur = rng.uniform(-0.3, 0.3, (n, 3))
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 33:36] = cur
    p = O.default_params()
    p["current"] = cur
    _, xT, _ = O.rollout(O.Model("wrench12", DT, p), "rk4", x0, U)
    e = B.Engine("wrench12", "f64")
    e.set_vehicle_physical(ph)
    assert normwise(cpu(e.rollout(x0, U, dt=DT).xT), xT) < TOL64
    ph[:, 33:36] = 0.0
    e.set_vehicle_physical(ph)
    plain = B.Engine("wrench12", "f64").rollout(x0, U, dt=DT).xT
    assert normwise(cpu(e.rollout(x0, U, dt=DT).xT), cpu(plain)) < 1e-14


def test_randomised_parity_sweep():
    """tests/tools/fuzz_parity.py: random model / integrator / size / horizon / stride / input layout / chunking / time
    slicing / initial lag against the C oracle (150 cases here; 4 x 800 were run on the box when it was written)."""
    import subprocess
    import sys
    from conftest import ROOT
    _c_oracle()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "fuzz_parity.py"), "150", "2"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "cases OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_randomised_evaluator_sweep():
    """tests/tools/fuzz_evaluator.py: random series lengths, horizon sets, window limits, models, integrators and
    precisions of the sliding-window evaluator against the C oracle; carried-lag mode against the numpy oracle."""
    import subprocess
    import sys
    from conftest import ROOT
    _c_oracle()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "fuzz_evaluator.py"), "80", "3"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "cases OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_index_space_beyond_32_bits(B):
    """4,194,304 fp32 vehicles x 160 steps: the per-vehicle input series has 5.4e9 elements (21 GB), so every offset
    past the first 2^32 elements needs 64-bit index arithmetic; first / middle / last vehicles against the C oracle."""
    CO = _c_oracle()
    n, T = 1 << 22, 160
    e = B.Engine("thruster8", "f32")
    g = torch.Generator(device="cuda").manual_seed(5)
    U = torch.rand((T, n, 8), device="cuda", generator=g) * 0.8 - 0.4
    assert U.numel() > 2 ** 32
    x0 = torch.zeros((n, 12), device="cuda")
    x0[:, 5] = torch.rand(n, device="cuda", generator=g) * 6 - 3
    r = e.rollout(x0, U, dt=DT, stride=80)
    idx = torch.cat([torch.arange(0, 64), torch.arange(n // 2 - 32, n // 2 + 32), torch.arange(n - 64, n)])
    snaps, xT, _ = CO.rollout("thruster8", "rk4", DT, cpu(x0[idx]), cpu(U[:, idx]), stride=80)
    assert r.traj.shape == (2, n, 12)
    assert normwise(cpu(r.xT[idx]), xT) < TOL32 and normwise(cpu(r.traj[:, idx]), snaps) < TOL32
    del U, r
    torch.cuda.empty_cache()
