"""Short target for `ncu --set full -k regex:rollout_kernel`: the wrench-input rollout kernels (BASELINE configs[3]
and the wrench rows of the throughput matrix), 100 RK4 steps per launch, inputs resident in HBM.

    python profiles/prof_wrench.py [launches]

Launch order (each `launches` times, default 2):
  wrench12 fp64, quat13 fp64, wrench12 fp64 + per-vehicle table + first-order wrench lag (Monte-Carlo, fp64),
  wrench12 fp32 + per-vehicle table + first-order wrench lag (Monte-Carlo, fp32: the bench's `monte_carlo` leg),
  wrench12 fp32, thruster8 fp64 (reference point).
Prints CUDA-event times so the ncu durations can be compared with warm timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bluerov2_dynamics_b200 as B  # noqa: E402

T = 100
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(3)


def mc_table(n):
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 9:27] *= rng.uniform(0.7, 1.3, (n, 18))
    ph[:, 27:30] = 1.0 / (ph[:, 0:1] - ph[:, 9:12])
    ph[:, 30:33] = 1.0 / (ph[:, 6:9] - ph[:, 12:15])
    ph[:, 36] = rng.uniform(0.05, 0.3, n)
    return ph


def run(model, dtype, n, mc):
    e = B.Engine(model, dtype)
    nx, nu = e.nx, e.nu
    if mc:
        e.set_wrench_lag1(True)
        e.set_vehicle_physical(mc_table(n))
    g = torch.Generator(device="cuda").manual_seed(1)
    if nu == 8:
        U = (torch.rand((T, n, nu), device="cuda", dtype=e.tdtype, generator=g) * 0.8 - 0.4).contiguous()
    else:
        scale = torch.tensor([40, 40, 40, 5, 5, 5.0], device="cuda", dtype=e.tdtype)
        U = ((torch.rand((T, n, nu), device="cuda", dtype=e.tdtype, generator=g) * 2 - 1) * scale).contiguous()
    x = torch.zeros((n, nx), device="cuda", dtype=e.tdtype)
    if nx == 13:
        x[:, 3] = 1.0
    lag = None
    if model == "thruster8":
        lag = torch.zeros((n, 18), device="cuda", dtype=e.tdtype)
    elif mc:
        lag = torch.zeros((n, 6), device="cuda", dtype=e.tdtype)

    def one(k):
        e.rollout(x, U, dt=0.02, integrator="rk4", lag0=lag, xT_out=x, lag_out=lag,
                  lag_repr="projected" if model == "thruster8" else "thruster", step0=k * T)
    one(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps):
        one(1 + k)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{model:10s} {dtype} mc={int(mc)} n={n}: {ms:8.4f} ms per {T} steps = {n * T / ms / 1e6:8.2f}e9 steps/s "
          f"finite={bool(torch.isfinite(x).all())}", flush=True)


run("wrench12", "f64", 1 << 16, False)
run("quat13", "f64", 1 << 16, False)
run("wrench12", "f64", 1 << 16, True)
run("wrench12", "f32", 1 << 20, True)
run("wrench12", "f32", 1 << 20, False)
run("thruster8", "f64", 1 << 16, False)
