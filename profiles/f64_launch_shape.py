"""fp64 rollout, 65,536 vehicles: time per RK4 step as a function of steps per launch and of the temporal-tiling
factor (time_slices; 0 = automatic) — separates per-launch / tail effects from the steady-state step cost."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bluerov2_dynamics_b200 as B  # noqa: E402

n = 65536
e = B.Engine("thruster8", "f64")
g = torch.Generator(device="cuda").manual_seed(1)
for steps in (25, 100, 400):
    U = (torch.rand((steps, n, 8), device="cuda", dtype=torch.float64, generator=g) * 0.8 - 0.4)
    for q in (1, 2, 4, 8, 0):
        x = torch.zeros((n, 12), device="cuda", dtype=torch.float64)
        lag = torch.zeros((n, 18), device="cuda", dtype=torch.float64)

        def one():
            e.rollout(x, U, dt=0.02, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected", time_slices=q)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            one()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"steps/launch {steps:4d}  time_slices {q}:  {ms:7.3f} ms  {1e3 * ms / steps:6.3f} us/step  {n * steps / ms / 1e6:6.2f}e9 steps/s", flush=True)
for nn in (37888, 75776, 148 * 2 * 128 * 3):   # whole multiples of the 296 resident blocks x 128 threads
    U = (torch.rand((100, nn, 8), device="cuda", dtype=torch.float64, generator=g) * 0.8 - 0.4)
    x = torch.zeros((nn, 12), device="cuda", dtype=torch.float64)
    lag = torch.zeros((nn, 18), device="cuda", dtype=torch.float64)
    for q in (1, 0):
        def one():
            e.rollout(x, U, dt=0.02, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected", time_slices=q)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            one()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"n {nn:6d} (whole waves)  time_slices {q}: {ms:7.3f} ms  {nn * 100 / ms / 1e6:6.2f}e9 steps/s", flush=True)
