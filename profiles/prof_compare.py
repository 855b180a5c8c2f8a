"""Runs the Koopman and PINc scoring kernels once each on a 1,000,100-row synthetic series (the bench's comparison
workload) — a short target for `ncu --set full -k regex:"koop_se|pinc_se"`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluerov2_dynamics_b200 import pinc as P  # noqa: E402
from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc  # noqa: E402

T, dt = 1_000_100, 0.02
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(4)
U = torch.rand((T, 8), device=dev, dtype=torch.float64, generator=g) * 0.8 - 0.4
X = 0.5 * torch.randn((T, 12), device=dev, dtype=torch.float64, generator=g)
rng = np.random.default_rng(12)
k = 500
K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=k, gamma=3.0)
K.centers_ = X[:k].cpu().numpy()
K.A_ = 0.98 * np.linalg.qr(rng.standard_normal((12 + k, 12 + k)))[0]
K.B_ = 0.02 * rng.standard_normal((12 + k, 8))
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "koop"):
    for H in (1, 100):
        print("koopman H", H, K.multistep_rmse(X, U, H))
    print("koopman multi", K.multistep_rmse_multi(X, U, [1, 10, 100]))
if which in ("both", "pinc"):
    cg = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_cmp.npz"))
    M = P.PincModel({kk[len("pinc_sd_"):]: cg[kk] for kk in cg.files if kk.startswith("pinc_sd_")})
    se, cnt = M.multistep_se(X, U, [1, 10, 100], dt, "reset")
    print("pinc", se.cpu().numpy()[:3], cnt)
torch.cuda.synchronize()
