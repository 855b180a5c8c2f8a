import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
rng = np.random.default_rng(1)
T, k = 1_000_100, 500
dev = torch.device("cuda")
X = torch.randn((T, 12), device=dev, dtype=torch.float64) * 0.5
U = torch.rand((T, 8), device=dev, dtype=torch.float64) - 0.5
K = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=k, gamma=3.0)
K.centers_ = X[:k].cpu().numpy(); K.A_ = 0.98 * np.linalg.qr(rng.standard_normal((512, 512)))[0]; K.B_ = 0.02 * rng.standard_normal((512, 8))
def tm(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for H in (1, 10, 100):
    print(H, "fused %.2f ms" % tm(lambda: K.multistep_rmse(X, U, H)), " lift+fir %.2f ms" % tm(lambda: K.multistep_rmse_multi(X, U, [H])))
print("all three: %.2f ms" % tm(lambda: K.multistep_rmse_multi(X, U, [1, 10, 100])))
