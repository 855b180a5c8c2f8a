"""GPU parity tests added in round 2 (all through the C ABI via bluerov2_dynamics_b200.Engine):

 * per-thruster lag states out of the projected-lag rollout kernel + lag epilogue vs the reference golden / oracle
 * the in-kernel command-signal generator: Philox stream vs the numpy restatement, generated == materialised bit for
   bit, chunking / slicing / sharding invariance, oracle parity on the materialised inputs
 * the TMA input ring vs the plain-load path, bit for bit, on ragged sizes and short horizons
 * health counters and min |cos theta| vs the C oracle; the cos(theta) clamp against the reference golden
 * argument validation of the evaluator geometry (direct C-ABI callers)
 * BASELINE configs[1] on the EXACT call bench.py times (projected lag carried in place, 1000-step launches, automatic
   time slices), every vehicle against the C oracle, at 1000 steps and at the configuration's full 10,000 steps

Tolerances (BASELINE.json north_star): fp64 <= 1e-10 normwise relative; fp32 <= 1e-4 after 1000 steps."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import normwise
from oracle import fossen_np as O
from oracle import inputgen_np as G

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


def _c_oracle():
    import subprocess
    from conftest import ROOT
    from oracle import c_oracle
    if not c_oracle.available():
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    return c_oracle


# ------------------------------------------------------------------------------------------------ lag epilogue
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
@pytest.mark.parametrize("integ", ["rk4", "euler"])
def test_per_thruster_lag_against_reference(B, golden_r2, dtype, tol, integ):
    """`ThrusterLag._x` of every thruster after a rollout (fossen/BlueROV2.py:503-510), from non-zero initial lag
    states, for calls shorter (1, 7) and longer (60 RK4 / 260) than the filter's memory: the kernel integrates the
    projected lag, the epilogue rebuilds the per-thruster states."""
    g = golden_r2
    e = B.Engine("thruster8", dtype)
    depth = e.carry_steps(DT, integ)
    assert 20 < depth < 400
    for T in (1, 7, 60, 260):
        r = e.rollout(g["lagtail_x0"], g["lagtail_U"][:T], dt=DT, integrator=integ, lag0=g["lagtail_lag0"].reshape(-1, 24))
        assert r.lag.shape == (5, 24)
        assert normwise(cpu(r.xT), g[f"lagtail_{integ}_T{T}_x"]) < (TOL64 if dtype == "f64" else TOL32), T
        assert normwise(cpu(r.lag).reshape(5, 8, 3), g[f"lagtail_{integ}_T{T}_lag"]) < tol, (T, depth)
    # chunked with the per-thruster carry (every chunk shorter than the memory) and through the host path
    x, lag = e.tensor(g["lagtail_x0"]), e.tensor(g["lagtail_lag0"].reshape(-1, 24))
    for c0 in range(0, 260, 13):
        r = e.rollout(x, g["lagtail_U"][c0:c0 + 13], dt=DT, integrator=integ, lag0=lag, step0=c0)
        x, lag = r.xT, r.lag
    assert normwise(cpu(lag).reshape(5, 8, 3), g[f"lagtail_{integ}_T260_lag"]) < tol
    Uh = np.ascontiguousarray(g["lagtail_U"].astype(e.ndtype))
    for chunk in (13, 64, 1000):
        xT, lagT, _ = e.rollout_host(g["lagtail_x0"].astype(e.ndtype), Uh, dt=DT, integrator=integ,
                                     lag0=g["lagtail_lag0"].reshape(-1, 24).astype(e.ndtype), chunk_steps=chunk)
        assert normwise(lagT.reshape(5, 8, 3), g[f"lagtail_{integ}_T260_lag"]) < tol, chunk
        assert normwise(xT, g[f"lagtail_{integ}_T260_x"]) < (TOL64 if dtype == "f64" else TOL32)


def test_default_and_projected_rollouts_run_the_same_kernel(B):
    """lag_repr='thruster' (the default) and 'projected' integrate with the same kernel: states agree bit for bit,
    and the per-thruster states project onto the kernel's own lag (to rounding)."""
    rng = np.random.default_rng(3)
    n, T = 3001, 140
    x0 = rng.uniform(-1, 1, (n, 12)) * 0.2
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.03)
    lag0 = rng.uniform(-0.05, 0.05, (n, 24))
    e = B.Engine("thruster8", "f64")
    a = e.rollout(x0, U, dt=DT, lag0=lag0, stride=20)
    b = e.rollout(x0, U, dt=DT, lag0=e.project_lag(lag0), lag_repr="projected", stride=20)
    assert normwise(cpu(a.xT), cpu(b.xT)) < 1e-13 and normwise(cpu(a.traj), cpu(b.traj)) < 1e-13
    assert normwise(cpu(e.project_lag(a.lag)), cpu(b.lag)) < 1e-13
    _, xT, lagT = O.rollout(O.Model("thruster8", DT), "rk4", x0[:64], U[:, :64], lag0=lag0[:64])
    assert normwise(cpu(a.lag[:64]).reshape(64, 8, 3), lagT) < 1e-12 and normwise(cpu(a.xT[:64]), xT) < TOL64
    # a projected lag_in cannot give per-thruster states back unless the call outlives the filter's memory
    depth = e.carry_steps(DT, "rk4")
    with pytest.raises(B.BrovError, match="cannot be recovered"):
        d_short = U[:depth - 1]
        _rollout_mixed(B, e, x0, d_short, e.project_lag(lag0))
    r = _rollout_mixed(B, e, x0, U, e.project_lag(lag0))
    assert normwise(cpu(r), cpu(a.lag)) < 1e-12
    # set_allocation changes what project_lag projects with
    al = B.default_allocation()[0].copy()
    al[3:5] *= 1.1
    e.set_allocation(al)
    assert normwise(cpu(e.project_lag(lag0)).reshape(n, 6, 3), np.einsum("ci,nik->nck", al, lag0.reshape(n, 8, 3))) < 1e-15


def _rollout_mixed(B, e, x0, U, z0):
    """projected lag_in, per-thruster lag_out: straight through the C ABI."""
    from bluerov2_dynamics_b200 import _lib as L
    x0t, Ut = e.tensor(x0), e.tensor(U)
    n = x0t.shape[0]
    xT, lag = torch.empty_like(x0t), torch.empty((n, 24), device="cuda", dtype=e.tdtype)
    d = L.RolloutDesc()
    d.struct_size = C.sizeof(L.RolloutDesc)
    d.integrator, d.n, d.steps, d.dt = L.RK4, n, Ut.shape[0], DT
    d.x0_dev, d.xT_dev, d.u_dev = x0t.data_ptr(), xT.data_ptr(), Ut.data_ptr()
    d.u_stride_t, d.u_stride_n = n * 8, 8
    d.lag_in_dev, d.lag_out_dev = z0.data_ptr(), lag.data_ptr()
    d.lag_in_repr, d.lag_out_repr = L.LAG_PROJECTED, L.LAG_THRUSTER
    d.stride = 1
    L.check(L.lib.brov_rollout(e._h, C.byref(d), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return lag


# ------------------------------------------------------------------------------------------------ input load paths
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind,nu", [("thruster8", 8), ("wrench12", 6)])
def test_vector_input_loads_equal_scalar_loads(B, dtype, kind, nu):
    """Aligned per-vehicle time-major inputs are prefetched with 128-bit loads; a view shifted by one scalar is not
    16-byte aligned and takes the scalar-load path.  Same bits, for ragged ensembles (partial warps, odd counts of
    24-byte rows), short horizons, and time-sliced launches."""
    e = B.Engine(kind, dtype)
    rng = np.random.default_rng(8)
    for n, T in ((1, 1), (31, 2), (33, 3), (97, 4), (129, 5), (1000, 9), (4097, 37)):
        x0 = rng.uniform(-1, 1, (n, 12)) * 0.1
        Uh = rng.uniform(-0.4, 0.4, (T, n, nu)) * (1.0 if nu == 8 else 10.0)
        U = e.tensor(Uh)
        buf = torch.empty(U.numel() + 1, device="cuda", dtype=e.tdtype)
        shifted = buf[1:].view(T, n, nu)
        shifted.copy_(U)
        assert U.data_ptr() % 16 == 0 and shifted.data_ptr() % 16 != 0
        for q in (1, 3):
            a = e.rollout(x0, U, dt=DT, stride=2, time_slices=q, want_lag=False)
            b = e.rollout(x0, shifted, dt=DT, stride=2, time_slices=q, want_lag=False)
            assert torch.equal(a.xT, b.xT) and torch.equal(a.traj, b.traj), (n, T, q)
        # and against the oracle
        _, xT, _ = O.rollout(O.Model(kind, DT), "rk4", x0, Uh)
        assert normwise(cpu(a.xT), xT) < (TOL64 if dtype == "f64" else TOL32)


# ------------------------------------------------------------------------------------------------ generated inputs
def test_generator_stream_matches_numpy_restatement(B):
    """The device's Philox4x32-10 counter stream and Box-Muller transform against oracle/inputgen_np.py (whose Philox is
    pinned to Random123's known answers): raw deviates (rho = 0, sigma = 1, no clip) to fast-math accuracy, for
    vehicles beyond 2^32 and steps beyond 2^32, both channel counts."""
    for kind, nu in (("thruster8", 8), ("wrench12", 6)):
        e = B.Engine(kind, "f64")
        for v0, s0 in ((0, 0), (7, 123456), ((1 << 33) + 5, (1 << 34) + 3)):
            gen = B.InputGenerator(seed=0x1234567887654321, rho=0.0, sigma=1.0, clip=float("inf"), vehicle0=v0)
            U, _ = e.generate_inputs(gen, steps=50, step0=s0, first=3, vstride=5, n_sel=40)
            want = G.normals(gen.seed, v0 + 3 + 5 * np.arange(40), s0 + np.arange(50))[..., :nu]
            assert np.max(np.abs(cpu(U) - want)) < 2e-5, (kind, v0, s0)
    # the reference's signal (defaults) in both precisions, against the restatement
    for dtype, tol in (("f64", 2e-6), ("f32", 2e-6)):
        e = B.Engine("thruster8", dtype)
        gen = B.InputGenerator(seed=99)
        U, s = e.generate_inputs(gen, steps=400, n_sel=64)
        want, sw = G.command_signal(99, np.arange(64), 0, 400, dtype=e.ndtype)
        assert np.max(np.abs(cpu(U) - want)) < tol and np.max(np.abs(cpu(s) - sw)) < tol
        assert 0.05 < float(U[200:].std()) < 0.15


@pytest.mark.parametrize("dtype,tol", [("f64", TOL64), ("f32", TOL32)])
@pytest.mark.parametrize("kind,nu", [("thruster8", 8), ("quat13", 6)])
def test_generated_inputs_rollout(B, dtype, tol, kind, nu):
    """Inputs generated inside the kernel: identical bits to a rollout fed the materialised signal, to chunked /
    time-sliced / sharded / host-path rollouts of the same stream; the oracle consumes the materialised inputs."""
    e = B.Engine(kind, dtype)
    rng = np.random.default_rng(21)
    n, T = 2500, 120
    x0 = np.zeros((n, e.nx))
    x0[:, :3] = rng.uniform(-1, 1, (n, 3))
    if e.nx == 13:
        x0[:, 3] = 1.0
    else:
        x0[:, 5] = rng.uniform(-3, 3, n)
    scale = None if nu == 8 else [40, 40, 40, 5, 5, 5.0]
    gen = B.InputGenerator(seed=2026, sigma=0.03, scale=scale, vehicle0=1000)
    U, s_end = e.generate_inputs(gen, steps=T, n_sel=n)
    mc = torch.empty(n, device="cuda", dtype=e.tdtype)
    a = e.rollout(x0, gen=gen, steps=T, dt=DT, stride=10, min_abs_cos=mc)
    b = e.rollout(x0, U, dt=DT, stride=10)
    well = (mc > 0.05).cpu().numpy()     # these commands tumble a few vehicles through theta = +-pi/2 (engine's own account)
    assert well.mean() > 0.95
    assert torch.equal(a.xT, b.xT) and torch.equal(a.traj, b.traj)
    assert torch.equal(a.gen_state, s_end)
    if a.lag is not None:
        assert torch.equal(a.lag, b.lag)          # per-thruster lag from the regenerated tail == from the array
    # chunks carry (x, the kernel's own lag, generator state); slices and a ragged number of chunks change nothing
    x, lag, gs = e.tensor(x0), None, None
    for c0 in range(0, T, 37):
        m = min(37, T - c0)
        r = e.rollout(x, gen=gen, steps=m, step0=c0, dt=DT, lag0=lag, gen_state=gs, time_slices=2 if m > 16 else 1,
                      lag_repr="projected")
        x, lag, gs = r.xT, r.lag, r.gen_state
    assert torch.equal(x, a.xT) and torch.equal(gs, a.gen_state)
    if a.lag is not None:   # ... and with the per-thruster carry (re-projected at every chunk start) to rounding
        x, lag, gs = e.tensor(x0), None, None
        for c0 in range(0, T, 37):
            r = e.rollout(x, gen=gen, steps=min(37, T - c0), step0=c0, dt=DT, lag0=lag, gen_state=gs)
            x, lag, gs = r.xT, r.lag, r.gen_state
        rt = 1e-12 if dtype == "f64" else 2e-5
        assert normwise(cpu(x)[well], cpu(a.xT)[well]) < rt and normwise(cpu(lag), cpu(a.lag)) < rt
    for q in (1, 3, 5):
        r = e.rollout(x0, gen=gen, steps=T, dt=DT, stride=10, time_slices=q)
        assert torch.equal(r.xT, a.xT) and torch.equal(r.traj, a.traj) and torch.equal(r.gen_state, a.gen_state), q
    # a shard sees its own slice of the stream through vehicle0
    sh = e.rollout(x0[700:900], gen=B.InputGenerator(seed=2026, sigma=0.03, scale=scale, vehicle0=1700), steps=T, dt=DT)
    assert torch.equal(sh.xT, a.xT[700:900])
    # host path: nothing but the states crosses PCIe
    health = np.zeros(2, np.uint64)
    xT, lagT, traj = e.rollout_host(x0.astype(e.ndtype), gen=gen, steps=T, dt=DT, stride=10, chunk_steps=50, health=health)
    assert np.array_equal(xT, a.xT.cpu().numpy()) and np.array_equal(traj, a.traj.cpu().numpy())
    # oracle parity on the identical inputs
    sub = slice(0, 256)
    snaps, xT_o, lag_o = O.rollout(O.Model(kind, DT), "rk4", x0[sub], cpu(U[:, sub]), stride=10)
    w = well[sub]
    assert normwise(cpu(a.xT[sub])[w], xT_o[w]) < tol and normwise(cpu(a.traj[:, sub])[:, w], snaps[:, w]) < tol
    if a.lag is not None:
        assert normwise(cpu(a.lag[sub]).reshape(-1, 8, 3), lag_o) < (1e-12 if dtype == "f64" else 2e-5)


# ------------------------------------------------------------------------------------------------ health
def test_health_counters_and_min_abs_cos(B):
    CO = _c_oracle()
    rng = np.random.default_rng(31)
    n, T = 3000, 200
    x0 = np.zeros((n, 12))
    x0[:, 4] = rng.uniform(-1.4, 1.4, n)          # pitched vehicles: some tumble through theta = +-pi/2
    x0[:, 10] = rng.uniform(-2, 2, n)             # pitch rate
    U = O.smooth_inputs(rng, T, 8, n=n, sigma=0.05)
    e = B.Engine("thruster8", "f64")
    mc = torch.empty(n, device="cuda", dtype=torch.float64)
    r = e.rollout(x0, U, dt=DT, health=True, min_abs_cos=mc, singular_eps=0.05)
    mco = np.ones(n)
    CO.rollout("thruster8", "rk4", DT, x0, U, min_abs_cos=mco)
    ok = mco > 0.02                                 # away from the singularity both trajectories agree
    assert np.max(np.abs(cpu(mc)[ok] - mco[ok])) < 2e-7     # the engine keeps this health metric in float32
    hc = r.health.cpu().numpy()
    assert hc[0] == 0 and abs(int(hc[1]) - int((mco < 0.05).sum())) <= 2 and hc[1] > 0
    # chunked accumulation and time slices give the same per-vehicle minima
    mc2 = torch.empty(n, device="cuda", dtype=torch.float64)
    a = e.rollout(x0, U[:90], dt=DT, min_abs_cos=mc2, time_slices=3, lag_repr="projected")
    b = e.rollout(a.xT, U[90:], dt=DT, lag0=a.lag, step0=90, min_abs_cos=mc2, min_abs_cos_accumulate=True, health=True,
                  singular_eps=0.05, lag_repr="projected", time_slices=2)
    assert torch.equal(mc2, mc) and torch.equal(b.health, r.health)
    # non-finite states are counted, quaternion models never report a singularity
    xb = x0.copy()
    xb[5, 0] = np.inf
    xb[17, 7] = np.nan
    r = e.rollout(xb, U[:5], dt=DT, health=True)
    assert int(r.health[0]) == 2
    q = B.Engine("quat13", "f32")
    xq = np.zeros((n, 13)); xq[:, 3] = 1.0
    rq = q.rollout(xq, O.smooth_inputs(rng, 20, 6, n=n, scale=5.0), dt=DT, health=True)
    assert rq.health.tolist() == [0, 0]
    # evaluator: windows near the singularity are reported next to the squared errors
    X = np.zeros((300, 12)); X[:, 4] = np.linspace(1.50, 1.64, 300)
    hv = torch.zeros(2, device="cuda", dtype=torch.int64)
    e.multistep_se(X, O.smooth_inputs(rng, 300, 8), [1], dt=DT, health_out=hv, singular_eps=0.01)
    want = int((np.abs(np.cos(X[:299, 4])) < 0.01).sum())      # H = 1: only the window starts are evaluated
    assert hv[0] == 0 and abs(int(hv[1]) - want) <= 1 and hv[1] > 0
    e.multistep_se(X, O.smooth_inputs(rng, 300, 8), [1, 5], dt=DT, health_out=hv, singular_eps=0.01)
    assert hv[0] == 0 and int(hv[1]) >= want                   # longer windows drift into the band as well


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 5e-5)])
def test_cos_theta_clamp_against_reference(B, golden_r2, dtype, tol):
    """The reference clamps |cos theta| < 1e-7 to 1e-7 sign(cos theta) (fossen/BlueROV2.py:52-56): right-hand sides and
    one Euler step on and next to theta = +-pi/2, both signs of the tiny cosine, against the unmodified reference."""
    g = golden_r2

    def rel(a, b):
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
    e, w = B.Engine("thruster8", dtype), B.Engine("wrench12", dtype)
    x, u8, tau = g["clamp_x"], g["clamp_u8"], g["clamp_tau"]
    want = {k: g[k] for k in ("clamp_xdot_thr", "clamp_euler_thr", "clamp_xdot_wrench", "clamp_euler_wrench")}
    if dtype == "f32":
        # float32 cannot hold pi/2 +- 5e-8: the rounded angles land on the other side of the singularity, so the fp32
        # engine is compared with the oracle (pinned to the same golden rows in tests/test_oracle_r2_golden.py)
        # evaluated at the float32-rounded states, where both see the same tiny cosines
        x, u8, tau = (a.astype(np.float32).astype(np.float64) for a in (x, u8, tau))
        m = O.Model("thruster8", DT)
        want["clamp_xdot_thr"] = m.f(x, u8, m.zero_lag(len(x)))[0]
        want["clamp_euler_thr"] = x + DT * want["clamp_xdot_thr"]
        want["clamp_xdot_wrench"] = O.rhs_wrench12(x, tau, O.default_params())
        want["clamp_euler_wrench"] = x + DT * want["clamp_xdot_wrench"]
    assert rel(cpu(e.rhs(x, u8, dt=DT)), want["clamp_xdot_thr"]) < tol
    r = e.rollout(x, u8[None], dt=DT, integrator="euler", health=True, singular_eps=1e-6)
    assert rel(cpu(r.xT), want["clamp_euler_thr"]) < tol
    assert rel(cpu(w.rhs(x, tau)), want["clamp_xdot_wrench"]) < tol
    rw = w.rollout(x, tau[None], dt=DT, integrator="euler")
    assert rel(cpu(rw.xT), want["clamp_euler_wrench"]) < tol
    if dtype == "f64":
        assert r.health.tolist() == [0, 7]          # every clamped row is reported, the control row is not
    # host helper (sign(0) = 0 is only reachable through the helper: no double has cos exactly 0)
    from bluerov2_dynamics_b200.fossen.BlueROV2 import euler_kinematics_matrix
    for i in range(8):
        assert np.allclose(euler_kinematics_matrix(g["clamp_x"][i, 3], g["clamp_x"][i, 4]), g["clamp_J2"][i], rtol=1e-13)


# ------------------------------------------------------------------------------------------------ validation
def test_evaluator_geometry_is_validated(B):
    """Direct C-ABI callers cannot make the evaluator kernels read outside X / U."""
    from bluerov2_dynamics_b200 import _lib as L
    e = B.Engine("thruster8", "f64")
    rows = 400
    X = torch.zeros((rows, 12), device="cuda", dtype=torch.float64)
    U = torch.zeros((rows, 8), device="cuda", dtype=torch.float64)
    se = torch.zeros(4, device="cuda", dtype=torch.float64)
    ws = torch.empty(L.lib.brov_se_workspace_bytes(rows), device="cuda", dtype=torch.uint8)

    def call(n_windows, window0, row0, carry, H=10):
        d = L.SeDesc()
        d.struct_size = C.sizeof(L.SeDesc)
        d.integrator, d.rows, d.n_windows, d.dt = L.RK4, rows, n_windows, DT
        d.X_dev, d.U_dev, d.se_out_dev = X.data_ptr(), U.data_ptr(), se.data_ptr()
        d.n_horizons, d.horizons[0] = 1, H
        d.workspace_dev, d.workspace_bytes = ws.data_ptr(), ws.numel()
        d.lag_carry, d.window0, d.row0 = carry, window0, row0
        return L.lib.brov_multistep_se(e._h, C.byref(d), None)

    depth = e.carry_steps(DT, "rk4")
    assert call(390, 0, 0, 0) == 0
    # windows start at local row window0 - row0: 395 windows from local row 10 on do not fit 400 rows
    assert call(395, 1000, 990, 1) == -1 and b"exceed" in L.lib.brov_last_error()
    # a carried-lag shard must hold the rows its first window replays
    assert call(380, 1000, 1000, 1) == -1 and b"must start at row" in L.lib.brov_last_error()
    need = (1000 * 10 - depth) // 10
    assert call(380 - (1000 - need), 1000, need, 1) == 0 and call(380 - (1000 - need), 1000, need + 1, 1) == -1
    assert call(10, 5, 0, 0) == -1                           # rows before the first window need lag_carry
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ the bench call, full size
def _cfg_x0(n, g, dtype):
    """Initial states of SURVEY 8(d) cfg2: positions in a tank-sized box, small roll / pitch, any yaw, at rest."""
    x = torch.zeros((n, 12), device="cuda", dtype=dtype)
    x[:, :2] = torch.rand((n, 2), device="cuda", dtype=dtype, generator=g) * 4 - 2
    x[:, 2] = torch.rand(n, device="cuda", dtype=dtype, generator=g) * 3
    x[:, 3:5] = torch.rand((n, 2), device="cuda", dtype=dtype, generator=g) * 0.4 - 0.2
    x[:, 5] = torch.rand(n, device="cuda", dtype=dtype, generator=g) * 6.28 - 3.14
    return x


# A vehicle that misses the tolerance is accepted as ill-conditioned — and only then — if the ORACLE itself does not
# determine its trajectory to that tolerance: its closest approach to theta = +-pi/2 (where the reference divides by
# cos theta clamped at 1e-7, fossen/BlueROV2.py:52-56) is below EPS_COS, or K_PERT re-runs of the oracle from x0
# perturbed by one unit in the last place of every component move its result by more than the tolerance.
EPS_COS = 0.01
K_PERT = 4
MAX_EXCLUDED_FRACTION = 5e-4


def _ill_conditioned(CO, x0_sub, inputs_of, steps, chunk, x_ref, tol, seed):
    """True per vehicle of the subset if K_PERT one-ulp perturbations of x0 move the oracle's own result by > tol."""
    rng = np.random.default_rng(seed)
    moved = np.zeros(len(x0_sub), bool)
    for _ in range(K_PERT):
        xp = np.nextafter(x0_sub, x0_sub + rng.choice([-1.0, 1.0], x0_sub.shape))
        lp = np.zeros((len(x0_sub), 8, 3))
        for c0 in range(0, steps, chunk):
            _, xp, lp = CO.rollout("thruster8", "rk4", DT, xp, inputs_of(c0), lag0=lp)
        err = np.max(np.abs(xp - x_ref), axis=1) / max(1.0, float(np.max(np.abs(x_ref))))
        moved |= err > tol
    return moved


@pytest.mark.parametrize("total_steps", [1000, 10000])
def test_bench_call_every_vehicle_against_c_oracle(B, total_steps):
    """BASELINE configs[1] on the call bench.py times: 65,536 fp64 vehicles, the reference's smooth random thrust
    commands generated in the kernel, 1000 RK4 steps per launch in place (xT aliases x0, allocation-projected lag and
    generator state carried in place, automatic time slices) — at 1000 steps and at the configuration's full 10,000.
    EVERY vehicle and its hidden lag state against the plain-C oracle fed the materialised inputs; the engine's own
    health counters say which vehicles came near the Euler-angle singularity."""
    CO = _c_oracle()
    n, chunk = 65536, 1000
    e = B.Engine("thruster8", "f64")
    g = torch.Generator(device="cuda").manual_seed(2026)
    x = _cfg_x0(n, g, torch.float64)
    x0 = cpu(x)
    gen = B.InputGenerator(seed=1)
    lag = torch.zeros((n, 18), device="cuda", dtype=torch.float64)
    gs = torch.zeros((n, 8), device="cuda", dtype=torch.float64)
    mc = torch.empty(n, device="cuda", dtype=torch.float64)
    xo, lo, mco = x0.copy(), np.zeros((n, 8, 3)), np.ones(n)
    gso = None
    for c0 in range(0, total_steps, chunk):
        r = e.rollout(x, gen=gen, steps=chunk, step0=c0, dt=DT, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected",
                      gen_state=gs, gen_state_out=gs, time_slices=0, health=True, min_abs_cos=mc,
                      min_abs_cos_accumulate=c0 > 0, singular_eps=EPS_COS)
        U, gso = e.generate_inputs(gen, steps=chunk, step0=c0, n_sel=n, state_in=gso)
        Uh = cpu(U)
        del U
        _, xo, lo = CO.rollout("thruster8", "rk4", DT, xo, Uh, lag0=lo, min_abs_cos=mco)
    assert torch.equal(gs, gso)
    xg = cpu(x)
    scale = max(1.0, float(np.max(np.abs(xo))))
    err = np.max(np.abs(xg - xo), axis=1) / scale
    lag_err = normwise(cpu(lag).reshape(n, 6, 3), np.einsum("ci,nik->nck", B.default_allocation()[0], lo))
    bad = np.flatnonzero(err > TOL64)
    # the engine's own singularity accounting agrees with the oracle's wherever the trajectories agree
    hc = r.health.cpu().numpy()
    good = err <= TOL64
    assert np.max(np.abs(cpu(mc)[good] - mco[good])) < 2e-7 and hc[0] == 0     # float32 health metric
    assert abs(int(hc[1]) - int((mco < EPS_COS).sum())) <= len(bad) + 2
    near = mco[bad] < EPS_COS
    rest = bad[~near]
    if len(rest):       # not explained by the singularity: ask the oracle how well it knows its own answer
        # generator states of the subset at each chunk start (regenerated: counter-based streams need no storage)
        st_all = {0: [None] * len(rest)}
        for c0 in range(0, total_steps - chunk, chunk):
            st_all[c0 + chunk] = [e.generate_inputs(gen, steps=chunk, step0=c0, first=int(v), n_sel=1,
                                                    state_in=st_all[c0][j])[1] for j, v in enumerate(rest)]

        def inputs_of(c0):
            cols = [cpu(e.generate_inputs(gen, steps=chunk, step0=c0, first=int(v), n_sel=1, state_in=st_all[c0][j])[0])
                    for j, v in enumerate(rest)]
            return np.concatenate(cols, axis=1)
        ill = _ill_conditioned(CO, x0[rest], inputs_of, total_steps, chunk, xo[rest], TOL64, seed=5)
        assert ill.all(), (f"{int((~ill).sum())} vehicles miss {TOL64:g} (worst {err[rest][~ill].max():.2e}) although the "
                           f"oracle is well conditioned on them; min |cos theta| {mco[rest][~ill]}")
    assert len(bad) <= MAX_EXCLUDED_FRACTION * n, (len(bad), float(err.max()))
    assert lag_err < 1e-12, lag_err
    print(f"[{total_steps} steps] {n - len(bad)} of {n} vehicles within {TOL64:g} (median {np.median(err):.1e}, worst accepted "
          f"{err[good].max():.2e}); {len(bad)} ill-conditioned ({int(near.sum())} within |cos theta| < {EPS_COS}); "
          f"projected lag {lag_err:.1e}; engine health {hc.tolist()}")


def test_bench_call_fp32_twin(B):
    """BASELINE configs[2] on the call bench.py times: 1,048,576 fp32 vehicles, generated commands, 100 steps per launch
    with stride-10 snapshots, state / projected lag / generator state carried in place — 1000 steps (the horizon the
    fp32 tolerance is stated at).  Every vehicle's snapshots of the last launch and final state against the float64 C
    oracle fed the same float32 inputs: 1e-4, with the singularity criterion of the fp64 test."""
    CO = _c_oracle()
    n, chunk, total = 1 << 20, 100, 1000
    e = B.Engine("thruster8", "f32")
    g = torch.Generator(device="cuda").manual_seed(2027)
    x = _cfg_x0(n, g, torch.float32)
    gen = B.InputGenerator(seed=2)
    lag = torch.zeros((n, 18), device="cuda", dtype=torch.float32)
    gs = torch.zeros((n, 8), device="cuda", dtype=torch.float32)
    traj = torch.empty((chunk // 10, n, 12), device="cuda", dtype=torch.float32)
    xo, lo, mco = cpu(x), np.zeros((n, 8, 3)), np.ones(n)
    gso, snaps = None, None
    for c0 in range(0, total, chunk):
        e.rollout(x, gen=gen, steps=chunk, step0=c0, dt=DT, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected",
                  gen_state=gs, gen_state_out=gs, stride=10, traj_out=traj)
        U, gso = e.generate_inputs(gen, steps=chunk, step0=c0, n_sel=n, state_in=gso)
        snaps, xo, lo = CO.rollout("thruster8", "rk4", DT, xo, cpu(U), lag0=lo, stride=10, min_abs_cos=mco)
        del U
    scale = max(1.0, float(np.max(np.abs(xo))))
    err = np.maximum(np.max(np.abs(cpu(x) - xo), axis=1), np.max(np.abs(cpu(traj) - snaps), axis=(0, 2))) / scale
    bad = err > TOL32
    # fp32 resolves cos theta to ~1e-7 relative to 1: the amplification 1 / cos theta eats the tolerance sooner
    assert np.all(mco[bad] < 0.05), (int(bad.sum()), np.sort(mco[bad])[-5:], float(err[mco >= 0.05].max()))
    assert bad.sum() <= MAX_EXCLUDED_FRACTION * n, int(bad.sum())
    print(f"fp32 twin: {int((~bad).sum())} of {n} vehicles within {TOL32:g} (median {np.median(err):.1e}); {int(bad.sum())} "
          f"near the singularity (max of their min |cos theta| {mco[bad].max() if bad.any() else 0:.3f})")


# ------------------------------------------------------------------------------------------------ sharded evaluator
@pytest.mark.parametrize("use_graph", [True, False])
def test_sharded_evaluator_matches_reference_golden(B, golden, use_graph):
    """dist.ShardedEvaluator (what bench.py's rmse leg and a multi-GPU caller run): `reset` scores every horizon in one
    pass, `carry` — the reference's literal semantics (training/train_tank_brov2_rk4.py:399-417, one model object for
    all windows) — one pass per horizon on concurrent streams inside one captured graph.  Both against the numbers the
    unmodified reference produced; rank r of a pretended world of 3 adds up to the same sums."""
    from bluerov2_dynamics_b200 import dist as D
    X, U = golden["rmse_X12"], golden["rmse_U8"]
    HS = [int(h) for h in golden["rmse_H"]]
    e = B.Engine("thruster8", "f64")
    for mode, key in (("reset", "rmse_thr_rk4_reset"), ("carry", "rmse_thr_rk4_carry")):
        ev = D.ShardedEvaluator(e, X, U, HS, DT, "rk4", 0, 1, lag_mode=mode, use_graph=use_graph)
        assert (ev.graph is not None) == use_graph, getattr(ev, "capture_error", None)
        for _ in range(3):          # replays give the same answer
            ev.run()
        torch.cuda.synchronize()
        r, hc = ev.rmse()
        assert np.allclose(r, golden[key], rtol=1e-10), (mode, r, golden[key])
        assert hc == [0, 0]
        # shards: the per-rank vectors (no process group here: the all-reduce is the identity) add up to the whole
        tot = np.zeros(4)
        for rank in range(3):
            es = D.ShardedEvaluator(e, X, U, HS, DT, "rk4", rank, 3, lag_mode=mode, use_graph=False)
            es.run()
            tot += es.buf.cpu().numpy()[:4]
        cnt = D.global_counts(len(X), HS)
        rs = [float(np.sqrt(tot[i] / (cnt[i] * 12))) for i in range(len(HS))]
        assert np.allclose(rs, golden[key], rtol=1e-10), (mode, rs)


@pytest.mark.parametrize("model,nu", [("thruster8", 8), ("wrench12", 6)])
def test_evaluator_time_slices_do_not_change_the_result(B, model, nu):
    """Temporal tiling of the evaluator (brov_se_desc.time_slices): the windows' states pass through memory between
    slices exactly, so forcing 2 / 3 / 4 slices, the automatic choice and the plain launch agree to summation order;
    health counters included.  3000 windows = 24 blocks, ragged tail (the last windows run out of rows)."""
    rng = np.random.default_rng(11)
    T, hs = 3100, [1, 10, 100]
    e = B.Engine(model, "f64")
    U = rng.uniform(-0.3, 0.3, (T, nu)) * (1.0 if nu == 8 else np.array([20, 20, 20, 2, 2, 2.0]))
    X = np.cumsum(rng.normal(0, 0.002, (T, 12)), axis=0)
    X[:, 4] = np.clip(X[:, 4], -1.0, 1.0)
    X[5, 4] = np.pi / 2 - 1e-4                  # one window starts next to the singularity: health counter
    ref = None
    for q in (1, 0, 2, 3, 4):
        hc = torch.zeros(2, dtype=torch.int64, device="cuda")
        se, cnt = e.multistep_se(X, U, hs, dt=DT, integrator="rk4", time_slices=q, health_out=hc, singular_eps=1e-3)
        got = (se.cpu().numpy()[:3], cnt, hc.cpu().numpy().tolist())
        if ref is None:
            ref = got
            assert ref[2][1] >= 1 and np.all(np.isfinite(ref[0]))
        assert got[1] == ref[1] and got[2] == ref[2], (q, got, ref)
        assert np.allclose(got[0], ref[0], rtol=1e-13, atol=0), (q, got[0], ref[0])
    with pytest.raises(Exception):
        e.multistep_se(X, U, hs, dt=DT, time_slices=9)
