"""BASELINE configs[1] at full length: 65,536 fp64 vehicles x 10,000 RK4 steps (dt = 0.02, per-vehicle random thrust),
every vehicle checked against the plain-C oracle (all host threads) — in 100 chunks of 100 steps carrying state and the
per-thruster lag on both sides.  Prints the normwise error every 1000 steps."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bluerov2_dynamics_b200 as B  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402

chunks, T, dt = int(sys.argv[1]) if len(sys.argv) > 1 else 100, 100, 0.02
dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"          # "f32": BASELINE configs[2] (1,048,576 vehicles), tolerance 1e-4
n, tol = (65536, 1e-10) if dtype == "f64" else (1 << 20, 1e-4)
g = torch.Generator(device="cuda").manual_seed(2026)
e = B.Engine("thruster8", dtype)
x = torch.zeros((n, 12), device="cuda", dtype=torch.float64)
x[:, :2] = torch.rand((n, 2), device="cuda", dtype=torch.float64, generator=g) * 4 - 2
x[:, 2] = torch.rand(n, device="cuda", dtype=torch.float64, generator=g) * 3
x[:, 3:5] = torch.rand((n, 2), device="cuda", dtype=torch.float64, generator=g) * 0.4 - 0.2
x[:, 5] = torch.rand(n, device="cuda", dtype=torch.float64, generator=g) * 6.28 - 3.14
lag = torch.zeros((n, 24), device="cuda", dtype=torch.float64)
xo, lo = x.cpu().numpy(), np.zeros((n, 8, 3))
# conditioning reference: the SAME C code compiled with FMA contraction (oracle/libbrov_oracle_fma.so) — a second, equally
# valid float64 evaluation of the reference's formulas whose roundings differ in the last bit at every step
xp, lp = xo.copy(), np.zeros((n, 8, 3))
u_prev = torch.zeros((n, 8), device="cuda", dtype=torch.float64)
min_cos = np.ones(n)
t_gpu = t_cpu = 0.0
worst = 0.0
for c in range(chunks):
    # the reference's smooth random command generator, per vehicle: u = clip(0.98 u + 0.02 N(0,1), -1, 1)
    noise = torch.randn((T, n, 8), device="cuda", dtype=torch.float64, generator=g) * 0.02
    U = torch.empty((T, n, 8), device="cuda", dtype=torch.float64)
    for k in range(T):
        u_prev = torch.clamp(0.98 * u_prev + noise[k], -1.0, 1.0)
        U[k] = u_prev
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if dtype == "f32":
        U = U.float().double()        # both sides see the same float32-representable inputs
    r = e.rollout(x.to(e.tdtype), U.to(e.tdtype), dt=dt, lag0=lag.to(e.tdtype), step0=c * T)
    torch.cuda.synchronize()
    t_gpu += time.perf_counter() - t0
    x, lag = r.xT, r.lag
    Uh = U.cpu().numpy()
    t0 = time.perf_counter()
    sn, xo, lo = CO.rollout("thruster8", "rk4", dt, xo, Uh, lag0=lo, stride=1)
    t_cpu += time.perf_counter() - t0
    min_cos = np.minimum(min_cos, np.abs(np.cos(sn[:, :, 4])).min(axis=0))   # closest approach to the pitch singularity
    _, xp, lp = CO.rollout("thruster8", "rk4", dt, xp, Uh, lag0=lp, fma=True)
    if (c + 1) % 10 == 0 or c == chunks - 1:
        xg = x.double().cpu().numpy()
        per = np.max(np.abs(xg - xo), axis=1) / max(1.0, float(np.max(np.abs(xo))))
        el = float(np.max(np.abs(lag.double().cpu().numpy().reshape(n, 8, 3) - lo)) / max(1.0, float(np.max(np.abs(lo)))))
        sens = np.max(np.abs(xp - xo), axis=1) / max(1.0, float(np.max(np.abs(xo))))
        bad, ill = per > tol, sens > tol
        worst = max(worst, float(per[~ill].max()), el)
        print(f"step {(c + 1) * T:6d}: GPU vs oracle: median vehicle {np.median(per):.1e}, worst well-conditioned vehicle "
              f"{per[~ill].max():.2e}, vehicles above {tol:g}: {int(bad.sum())} (of which the two CPU evaluations of the "
              f"reference formulas — with / without FMA — also differ by more than {tol:g}: {int((bad & ill).sum())}; "
              f"such vehicles in total: {int(ill.sum())}, worst {sens.max():.1e}); lag {el:.1e}", flush=True)
print(f"closest approach to the Euler-angle singularity, min |cos theta| along the trajectory: deviating vehicles "
      f"{np.sort(min_cos[bad]).round(4).tolist()}; all others: min {min_cos[~bad].min():.4f}, median {np.median(min_cos[~bad]):.3f}")
print(f"configs[{1 if dtype == 'f64' else 2}] ({dtype}): {n} vehicles x {chunks * T} RK4 steps; {int((~bad).sum())} vehicles within {tol:g} "
      f"(worst {per[~bad].max():.2e}); worst normwise error over vehicles on which the two CPU evaluations agree {worst:.3e}; "
      f"GPU {t_gpu:.2f} s, C oracle on {CO.threads()} threads {t_cpu:.1f} s")
