"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on B200s, gloo in CPU tests).

The path shards without any data-path exchange: vehicles of a rollout and windows of the RMSE evaluator are
independent, so each rank takes a contiguous index range.  The only collective is the sum of the per-rank squared
error vectors (<= 4 doubles) of the evaluator."""
from __future__ import annotations

import os
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of range(n): [lo, hi) of `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def window_shard(T: int, horizons: Sequence[int], rank: int, world: int):
    """Rows and windows of one rank for the sliding-window evaluator.
    Global windows k = 0..T-Hmin-1 are block-partitioned; the rank needs series rows [lo, min(T, hi + Hmax)) — an
    Hmax-row halo replicated at load time, no exchange.  Returns (row_lo, row_hi, n_windows_local)."""
    hmin, hmax = min(horizons), max(horizons)
    nwin = max(T - hmin, 0)
    lo, hi = shard_range(nwin, rank, world)
    return lo, min(T, hi + hmax) if hi > lo else lo, hi - lo


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def global_counts(T: int, horizons: Sequence[int]) -> list[int]:
    return [max(T - int(h), 0) for h in horizons]


def sharded_multistep_rmse(local_se: Callable[[np.ndarray, np.ndarray, Sequence[int], int], torch.Tensor],
                           X: np.ndarray, U: np.ndarray, horizons: Sequence[int], rank: int, world: int):
    """Endpoint RMSE per horizon of the whole series, computed by `world` ranks.

    local_se(X_rows, U_rows, horizons, n_windows) -> tensor of per-horizon squared-error sums for the rank's
    windows (on the GPU: Engine.multistep_se; in CPU tests: the oracle).  The sums are all-reduced, then
    rmse_h = sqrt(se_h / ((T - H_h) * n_states))."""
    T = len(X)
    hs = [int(h) for h in horizons]
    lo, hi, nloc = window_shard(T, hs, rank, world)
    if nloc > 0:
        se = local_se(X[lo:hi], U[lo:hi], hs, nloc)
    else:
        se = None
    dev = se.device if se is not None else (torch.device("cuda", torch.cuda.current_device())
                                            if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu"))
    vec = torch.zeros(len(hs), dtype=torch.float64, device=dev)
    if se is not None:
        vec += se[:len(hs)].to(torch.float64)
    allreduce_sum_(vec)
    cnt = global_counts(T, hs)
    vals = vec.cpu().numpy()
    nx = X.shape[1]
    return [float(np.sqrt(vals[i] / (cnt[i] * nx))) if cnt[i] > 0 else float("nan") for i in range(len(hs))]
