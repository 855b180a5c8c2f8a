"""Randomised sweep of the sliding-window evaluator (brov_multistep_se) against the plain-C oracle: random model,
integrator, precision, series length, horizon sets, window limits; the carried-lag mode against the numpy oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bluerov2_dynamics_b200 as B  # noqa: E402
from oracle import c_oracle as CO, fossen_np as O  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
DT = 0.02
engines = {}
worst = {"f64": 0.0, "f32": 0.0}
for case in range(cases):
    model = rng.choice(["thruster8", "wrench12", "quat13"])
    integ = rng.choice(["rk4", "euler"])
    dtype = rng.choice(["f64", "f32"])
    T = int(rng.integers(3, 500))
    nx, nu = (13 if model == "quat13" else 12), (8 if model == "thruster8" else 6)
    amp = np.full(8, 0.4) if nu == 8 else np.array([8.0, 8.0, 8.0, 0.3, 0.3, 0.3])
    U = O.smooth_inputs(rng, T, nu, sigma=0.05) * amp / 0.4 * (0.4 if nu == 8 else 0.4)
    x0 = np.zeros((1, nx))
    if nx == 13:
        x0[0, 3] = 1.0
    snaps, _, _ = CO.rollout(model, "rk4", DT, x0, U, stride=1)
    X = np.vstack([x0, snaps[:, 0]])[:T] + 1e-3 * rng.standard_normal((T, nx))   # a "recorded" series with sensor noise
    nd = np.float32 if dtype == "f32" else np.float64
    X, U = X.astype(nd).astype(np.float64), U.astype(nd).astype(np.float64)
    nh = int(rng.integers(1, 5))
    hs = sorted(set(int(h) for h in rng.integers(1, max(2, min(T + 5, 120)), nh)))
    e = engines.get((model, dtype)) or engines.setdefault((model, dtype), B.Engine(model, dtype))
    nwin = None if rng.random() < 0.6 else int(rng.integers(0, max(1, T - hs[0]) + 1))
    se, cnt = e.multistep_se(X, U, hs, dt=DT, integrator=integ, n_windows=nwin)
    se = se.cpu().numpy()
    for i, h in enumerate(hs):
        full = max(T - h, 0)
        want = full if nwin is None else min(nwin, full)
        assert cnt[i] == want, (case, hs, cnt, nwin, T)
        if want == 0:
            assert se[i] == 0.0
            continue
        ref = CO.multistep_se(model, integ, DT, X[:want + h], U[:want + h], h)[0]
        err = abs(se[i] - ref) / max(ref, 1e-30)
        tol = 1e-9 if dtype == "f64" else 2e-3        # squared errors of fp32 endpoints: relative 1e-4 .. 1e-3
        worst[dtype] = max(worst[dtype], err)
        assert err < tol, (case, model, integ, dtype, T, hs, h, se[i], ref, err)
    if model == "thruster8" and dtype == "f64" and T <= 160 and rng.random() < 0.5:
        h = hs[0]
        if T - h > 0:
            got = e.multistep_rmse(X, U, h, dt=DT, integrator=integ, lag_mode="carry")
            ref = O.multistep_se(O.Model("thruster8", DT), integ, X, U, [h], lag_mode="carry")[h][2]
            assert abs(got - ref) <= 1e-9 * max(ref, 1e-30), (case, "carry", T, h, got, ref)
    if case % 20 == 0:
        print(f"case {case:4d} {model:9s} {integ:5s} {dtype} T={T:3d} H={hs} n_windows={nwin} ok", flush=True)
print(f"fuzz_evaluator: {cases} cases OK; worst relative error of the squared-error sums fp64 {worst['f64']:.2e}, fp32 {worst['f32']:.2e}")
