"""Times the 1M-window endpoint-RMSE evaluator (configs[4]) in both lag semantics through dist.ShardedEvaluator on one
GPU: `reset` (one pass, all horizons) and `carry` (the reference's literal semantics: one pass per horizon, concurrent
streams in one graph).  Usage: python profiles/evaluator_timing.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bluerov2_dynamics_b200 as B  # noqa: E402
from bluerov2_dynamics_b200 import dist as D  # noqa: E402

T, HS = 1_000_100, [1, 10, 100]
eng = B.Engine("thruster8", "f64")
X, U = bench.tank_series(torch, eng.device, T)


def time_ev(evs, reps=10):
    for _ in range(3):
        for ev in evs:
            ev.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for ev in evs:
            ev.run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ev = D.ShardedEvaluator(eng, X, U, HS, 0.02, "rk4", 0, 1, lag_mode="reset")
ms_r = time_ev([ev])
print(f"reset, one pass:                      {ms_r:7.3f} ms  rmse {ev.rmse()[0]}")
evs = [D.ShardedEvaluator(eng, X, U, [h], 0.02, "rk4", 0, 1, lag_mode="carry") for h in HS]
for h, e1 in zip(HS, evs):
    print(f"carry, H = {h:3d} alone:                 {time_ev([e1]):7.3f} ms")
ms_s = time_ev(evs)
print(f"carry, three passes back to back:     {ms_s:7.3f} ms  ({ms_s / ms_r:.3f} x reset)")
evc = D.ShardedEvaluator(eng, X, U, HS, 0.02, "rk4", 0, 1, lag_mode="carry")
ms_c = time_ev([evc])
print(f"carry, three passes concurrent:       {ms_c:7.3f} ms  ({ms_c / ms_r:.3f} x reset)  rmse {evc.rmse()[0]}")
print("back-to-back rmse", [e1.rmse()[0][0] for e1 in evs])

# the 8-GPU shard of the same series on ONE GPU: 125,000 windows = 977 blocks on 296 resident slots = 3.3 rounds
n8 = 125_000
Xs, Us = X[:n8 + 100].contiguous(), U[:n8 + 100].contiguous()
for q, label in ((1, "plain launch      "), (0, "automatic slicing "), (2, "2 slices          "), (3, "3 slices          "),
                 (4, "4 slices          ")):
    fn = lambda: eng.multistep_se(Xs, Us, HS, dt=0.02, integrator="rk4", n_windows=n8, time_slices=q)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        se, _ = fn()
    b.record()
    torch.cuda.synchronize()
    print(f"125,000 windows (1/8 shard), {label}: {a.elapsed_time(b) / 20:6.3f} ms  se {se.cpu().numpy()[:3]}")
