import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
rng = np.random.default_rng(1)
T, k = 50_000, 500
dev = torch.device("cuda")
X = torch.randn((T, 12), device=dev, dtype=torch.float64) * 0.5
U = torch.rand((T, 8), device=dev, dtype=torch.float64) - 0.5
K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=k, gamma=3.0)
K.centers_ = X[:k].cpu().numpy(); K.A_ = 0.98 * np.linalg.qr(rng.standard_normal((512, 512)))[0]; K.B_ = 0.02 * rng.standard_normal((512, 8))
K.multistep_rmse(X, U, 1); torch.cuda.synchronize()
for tag in ("first H=100 (builds 99 decoder-row powers)", "second H=100 (cached)"):
    t0 = time.perf_counter(); r = K.multistep_rmse(X, U, 100); torch.cuda.synchronize(); print(tag, f"{1e3*(time.perf_counter()-t0):.2f} ms", r)
