import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bluerov2_dynamics_b200 as B
rng = np.random.default_rng(3)
n, T = 1 << 20, 100
def table(n):
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 9:27] *= rng.uniform(0.7, 1.3, (n, 18))
    ph[:, 27:30] = 1.0 / (ph[:, 0:1] - ph[:, 9:12]); ph[:, 30:33] = 1.0 / (ph[:, 6:9] - ph[:, 12:15])
    ph[:, 36] = rng.uniform(0.05, 0.3, n)
    return ph
g = torch.Generator(device="cuda").manual_seed(1)
amp = torch.tensor([40, 40, 40, 5, 5, 5.0], device="cuda")
U = [((torch.rand((T, n, 6), device="cuda", generator=g) * 2 - 1) * amp).contiguous() for _ in range(2)]
ph = table(n)
for name, lag1, pv in (("plain", False, False), ("lag1", True, False), ("pv", False, True), ("pv+lag1", True, True)):
    e = B.Engine("wrench12", "f32")
    if lag1: e.set_wrench_lag1(True)
    if pv: e.set_vehicle_physical(ph)
    x = torch.zeros((n, 12), device="cuda")
    lag = torch.zeros((n, 6), device="cuda") if lag1 else None
    fn = lambda k: e.rollout(x, U[k % 2], dt=0.02, lag0=lag, xT_out=x, lag_out=lag, step0=k * T)
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(10): fn(k)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name:8s} {ms:7.3f} ms  {n * T / ms / 1e6:6.2f}e9 steps/s", flush=True)
