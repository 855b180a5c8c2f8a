"""Times the PINc evaluator (1,000,100-row series, H = 1/10/100, every window) on the tensor-core path and — with
BROV_PINC_TC=0 in the environment — on the CUDA-core path.  Usage: python profiles/pinc_timing.py [H]  (one horizon
only, e.g. 10, for a short run under ncu)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluerov2_dynamics_b200 import pinc as P  # noqa: E402

T, dt = 1_000_100, 0.02
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(4)
U = torch.rand((T, 8), device=dev, dtype=torch.float64, generator=g) * 0.8 - 0.4
X = 0.5 * torch.randn((T, 12), device=dev, dtype=torch.float64, generator=g)
cg = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_cmp.npz"))
M = P.PincModel({kk[len("pinc_sd_"):]: cg[kk] for kk in cg.files if kk.startswith("pinc_sd_")})
out = M(cg["pinc_dataset_zin"][:64]).cpu().numpy().astype(np.float64)
ref = cg["pinc_forward"]
print(f"forward vs the reference's torch network: normwise {np.max(np.abs(out - ref)) / max(np.max(np.abs(ref)), 1.0):.2e}", flush=True)
for hs in ([[int(sys.argv[1])]] if len(sys.argv) > 1 else ([1, 10, 100], [100], [10])):
    se, cnt = M.multistep_se(X, U, hs, dt, "reset")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        se, cnt = M.multistep_se(X, U, hs, dt, "reset")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    steps = float((T - hs[-1]) * hs[-1])
    flop = steps * 2.0 * (14 * 64 + 3 * 64 * 64 + 64 * 9)
    print(f"tc={os.environ.get('BROV_PINC_TC', '1')} H={hs}: {ms:8.3f} ms  {flop / ms / 1e9:7.2f} TFLOP/s algorithmic  se={se.cpu().numpy()[:len(hs)]}", flush=True)
