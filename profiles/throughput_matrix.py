"""Vehicle-steps/s of every model kind x integrator x precision on one B200 (inputs resident in HBM, per-vehicle input
series, 100 steps per launch, no trajectory writeback): fp32 with 1,048,576 vehicles, fp64 with 65,536."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bluerov2_dynamics_b200 as B  # noqa: E402

T = 100
out = {}
rng = np.random.default_rng(0)
for model, nx, nu in (("thruster8", 12, 8), ("wrench12", 12, 6), ("quat13", 13, 6), ("di12_u8", 12, 8), ("diq13_u6", 13, 6)):
    for dtype, n in (("f32", 1 << 20), ("f64", 1 << 16)):
        e = B.Engine(model, dtype)
        if e.is_di:
            e.set_di_gains(rng.normal(0, 0.05, (nu, 3)), rng.normal(0, 0.05, (nu, 3)))
        g = torch.Generator(device="cuda").manual_seed(1)
        scale = 0.4 if nu == 8 else 10.0
        U = [((torch.rand((T, n, nu), device="cuda", dtype=e.tdtype, generator=g) * 2 - 1) * scale).contiguous() for _ in range(2)]
        x = torch.zeros((n, nx), device="cuda", dtype=e.tdtype)
        if nx == 13:
            x[:, 3] = 1.0
        lag = torch.zeros((n, 18), device="cuda", dtype=e.tdtype) if model == "thruster8" else None
        for integ in ("rk4", "euler"):
            if model == "diq13_u6" and integ == "rk4":
                continue

            def one(k):
                e.rollout(x, U[k % 2], dt=0.02, integrator=integ, lag0=lag, xT_out=x, lag_out=lag,
                          lag_repr="projected" if lag is not None else "thruster", step0=k * T)
            for k in range(3):
                one(k)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for k in range(10):
                one(k)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            out[f"{model}/{integ}/{dtype}"] = {"ms": round(ms, 4), "steps_per_s": n * T / (ms * 1e-3)}
            print(f"{model:10s} {integ:5s} {dtype}: {ms:8.3f} ms  {n * T / (ms * 1e-3) / 1e9:7.2f}e9 steps/s", flush=True)
            x.zero_()
            if nx == 13:
                x[:, 3] = 1.0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "throughput_matrix.json"), "w"), indent=1)
