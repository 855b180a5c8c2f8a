"""torchrun check of the sharded carried-lag evaluator on N GPUs against the unmodified reference's numbers:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/check_sharded_carry.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import bluerov2_dynamics_b200 as B  # noqa: E402
from bluerov2_dynamics_b200 import dist as D  # noqa: E402

rank, world, local = D.init_from_env()
torch.cuda.set_device(local)
g = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
X, U, HS = g["rmse_X12"], g["rmse_U8"], [int(h) for h in g["rmse_H"]]
eng = B.Engine("thruster8", "f64", device=local)
ok = True
for integ in ("rk4", "euler"):
    got = [D.sharded_multistep_rmse_carry(eng, X, U, h, 0.02, integ, rank, world) for h in HS]
    ref = g[f"rmse_thr_{integ}_carry"]
    good = bool(np.allclose(got, ref, rtol=1e-9))
    ok &= good
    if rank == 0:
        print(integ, "sharded over", world, "GPUs:", got, "reference:", ref.tolist(), "OK" if good else "MISMATCH")
# windows sharded in reset mode through the generic helper, all horizons in one pass
got = D.sharded_multistep_rmse(lambda Xr, Ur, hs, n: eng.multistep_se(Xr, Ur, hs, dt=0.02, integrator="rk4", n_windows=n)[0],
                               X, U, HS, rank, world)
good = bool(np.allclose(got, g["rmse_thr_rk4_reset"], rtol=1e-9))
ok &= good
if rank == 0:
    print("reset mode:", got, "OK" if good else "MISMATCH")
    print("SHARDED_CARRY_OK" if ok else "SHARDED_CARRY_FAILED")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
