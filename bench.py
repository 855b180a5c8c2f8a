#!/usr/bin/env python3
"""bench.py — RK4 vehicle-steps/s of the batched BlueROV2 Fossen rollout (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one batch of synthetic input: ONE rollout-kernel launch that advances
every vehicle of the ensemble by a chunk of RK4 steps (1000 for configs[1], 100 for configs[2]; the inputs of a whole
10,000-step rollout do not fit HBM for configs[2]) under per-vehicle random thrust inputs.  K = 100 steps is
the full 10,000-step rollout of BASELINE configs[1] / configs[2].

Primary line (`value`, `roofline`, `e2e`): configs[1] — 65,536 vehicles per GPU, fp64, 8-thruster model with the
3rd-order lag, dt = 0.02.  The `fp32` object of the same line carries configs[2] — 1,048,576 vehicles per GPU, fp32,
trajectory writeback every 10 steps.  N > 1: the ensemble is sharded by vehicle, each rank runs its own shard
(weak scaling, no data-path collective); `rmse` times the multi-horizon evaluator, whose per-rank squared-error
sums are combined with one NCCL all-reduce.

Inputs are resident in HBM for `value` (a ring of two chunk buffers, each far larger than the 126 MB L2, used
alternately) and in pinned host memory for `e2e` (every step copies its chunk host->device and reads the step's
final state back, through Engine.rollout_host = brov_rollout_host of the C ABI).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there on every rank), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a saved
# copy of the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

METRIC = "rk4_vehicle_steps_per_s"
UNIT = "vehicle-steps/s"
DT = 0.02
CHUNK = 100                      # RK4 steps per launch ("step" of the bench)
FLOP_PER_STEP = {"thruster8": 1756.0, "wrench12": 676.0, "quat13": 805.0}   # SURVEY 8(d), algorithmic
# chunk = RK4 steps per launch.  cfg2: 65,536 vehicles are 1.73 waves of resident blocks, so longer launches amortise the
# tail (100 steps 14.0e9, 250 steps 14.8e9, 1000 steps 15.1e9 vehicle-steps/s); cfg3 is flat in the chunk length.
CFG2 = dict(name="cfg2", model="thruster8", dtype="f64", n_per_gpu=65536, stride=0, chunk=1000)
CFG3 = dict(name="cfg3", model="thruster8", dtype="f32", n_per_gpu=1 << 20, stride=10, chunk=100)


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle; the only place bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One host core: `nveh` vehicles, one at a time, `steps` RK4 steps each — numpy float64, one vehicle per Python
    call, the way the reference executes simulate_physics (training/train_tank_brov2_rk4.py:375-396)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    seed, nveh, steps = job
    from oracle import fossen_np as O
    rng = np.random.default_rng(seed)
    m = O.Model("thruster8", DT)
    t0 = time.perf_counter()
    for _ in range(nveh):
        x = np.zeros((1, 12))
        x[0, 2] = 1.0
        U = rng.uniform(-0.4, 0.4, (steps, 8))
        O.rollout(m, "rk4", x, U)
    return time.perf_counter() - t0


def cpu_reference_rate(cores: int, steps_per_core: int, seed: int = 0):
    """Aggregate RK4 vehicle-steps/s of `cores` processes each rolling one vehicle for steps_per_core steps."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    jobs = [(seed + i, 1, steps_per_core) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return cores * steps_per_core / wall, wall


def cpu_batched_rate(n: int = 4096, steps: int = 20):
    """The same oracle vectorised over n vehicles in one process (numpy's own threading): a stronger CPU baseline
    than the reference's one-vehicle-per-call structure."""
    from oracle import fossen_np as O
    rng = np.random.default_rng(1)
    x = np.zeros((n, 12))
    U = rng.uniform(-0.4, 0.4, (steps, n, 8))
    m = O.Model("thruster8", DT)
    O.rollout(m, "rk4", x, U[:2])
    t0 = time.perf_counter()
    O.rollout(m, "rk4", x, U)
    return n * steps / (time.perf_counter() - t0)


def cpu_c_oracle_rate(n: int = 8192, steps: int = 400):
    """The plain-C oracle (oracle/brov_oracle.c: scalar float64, reference operation structure, OpenMP over vehicles)
    on all host cores: the strongest CPU baseline — compiled code instead of the reference's Python."""
    from oracle import c_oracle as CO
    if not CO.available():
        return None, 0
    rng = np.random.default_rng(2)
    x = np.zeros((n, 12))
    x[:, 2] = 1.0
    U = rng.uniform(-0.4, 0.4, (steps, n, 8))
    CO.rollout("thruster8", "rk4", DT, x[:64], U[:10, :64])
    t0 = time.perf_counter()
    CO.rollout("thruster8", "rk4", DT, x, U)
    return n * steps / (time.perf_counter() - t0), CO.threads()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        """windows: list of (t0, t1) perf_counter intervals that were timed."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if windows and not any(a - 0.05 <= t <= b + 0.15 for a, b in windows):
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def make_inputs(torch, eng, n, chunk, ring, seed):
    """Per-vehicle random thrust inputs [ring][chunk][n][8], uniform in [-0.4, 0.4] (SURVEY 8(d) cfg2 range),
    generated on the device by torch's Philox generator; x0 = scattered positions, random yaw, at rest."""
    g = torch.Generator(device=eng.device).manual_seed(seed)
    U = [(torch.rand((chunk, n, 8), device=eng.device, dtype=eng.tdtype, generator=g) * 0.8 - 0.4).contiguous()
         for _ in range(ring)]
    x0 = torch.zeros((n, 12), device=eng.device, dtype=eng.tdtype)
    x0[:, 0:2] = torch.rand((n, 2), device=eng.device, dtype=eng.tdtype, generator=g) * 4 - 2
    x0[:, 2] = torch.rand(n, device=eng.device, dtype=eng.tdtype, generator=g) * 3
    x0[:, 5] = torch.rand(n, device=eng.device, dtype=eng.tdtype, generator=g) * 6.2 - 3.1
    return U, x0


def run_rollout_leg(torch, dist, B, cfg, steps, warmup, local, world, windows):
    eng = B.Engine(cfg["model"], cfg["dtype"], device=local)
    n, stride, chunk = cfg["n_per_gpu"], cfg["stride"], cfg["chunk"]
    ring = 2
    U, x0 = make_inputs(torch, eng, n, chunk, ring, seed=1000 + int(os.environ.get("RANK", "0")))
    x = x0.clone()
    lag = torch.zeros((n, 18), device=eng.device, dtype=eng.tdtype)  # allocation-projected lag carried between chunks
    traj = [torch.empty((chunk // stride, n, 12), device=eng.device, dtype=eng.tdtype) for _ in range(ring)] if stride else None

    def one(k):
        eng.rollout(x, U[k % ring], dt=DT, integrator="rk4", lag0=lag, stride=stride, step0=k * chunk, xT_out=x,
                    lag_out=lag, traj_out=traj[k % ring] if stride else None, lag_repr="projected")

    for k in range(warmup):
        one(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(steps):
        one(warmup + k)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    windows.append((t0, t1))
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=eng.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    finite = bool(torch.isfinite(x).all().item())
    sz = 8 if cfg["dtype"] == "f64" else 4
    bytes_per_launch = n * chunk * 8 * sz + (n * 12 * sz * (chunk // stride) if stride else 0) + 2 * n * 30 * sz
    return dict(ms_total=ms, ms_per_step=ms / steps, vehicle_steps=float(n) * world * chunk * steps, finite=finite,
                bytes_per_launch=bytes_per_launch, n=n, eng=eng, U=U, x0=x0)


def run_e2e_leg(torch, dist, B, cfg, leg, steps, warmup, world, windows):
    """Same workload through the host-buffer API: per step, the chunk's inputs + x0 + lag go host->device from
    pinned memory and the step's final state + lag come back."""
    eng, n = leg["eng"], leg["n"]
    chunk = min(cfg["chunk"], 250)   # 1.07 GB of inputs per call from pinned memory (keeps 8 ranks' pinned pools small)
    ring = 2
    Uh = [B.pinned_empty((chunk, n, 8), eng.ndtype) for _ in range(ring)]
    for r in range(ring):
        Uh[r][...] = leg["U"][r][:chunk].cpu().numpy()
    xh = B.pinned_empty((n, 12), eng.ndtype)
    xh[...] = leg["x0"].cpu().numpy()
    lagh = B.pinned_empty((n, 18), eng.ndtype)
    lagh[...] = 0

    def one(k):
        eng.rollout_host(xh, Uh[k % ring], dt=DT, integrator="rk4", lag0=lagh, out_xT=xh, out_lag=lagh,
                         chunk_steps=max(chunk // 10, 25), lag_repr="projected")

    for k in range(max(1, min(warmup, 3))):
        one(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    windows.append((t0, t1))
    sec = t1 - t0
    if world > 1:
        t = torch.tensor([sec], device=eng.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    h2d = Uh[0].nbytes + xh.nbytes + lagh.nbytes
    d2h = xh.nbytes + lagh.nbytes
    return dict(value=float(n) * world * chunk * steps / sec, unit=UNIT, h2d_bytes_per_step=int(h2d),
                d2h_bytes_per_step=int(d2h), steps=steps, ms_per_step=1e3 * sec / steps,
                h2d_gbs=h2d * steps / sec / 1e9,
                api="Engine.rollout_host -> brov_rollout_host (C ABI), pinned host buffers, 250 RK4 steps per call in 10 sub-chunks "
                    "double-buffered on a copy stream")


def run_rmse_leg(torch, dist, B, local, rank, world, windows, T=1_000_100, horizons=(1, 10, 100)):
    """cfg5: multi-horizon endpoint RMSE over ~1M sliding windows of a synthetic 50 Hz series; windows sharded over
    ranks (H-row halo), per-rank squared-error sums all-reduced with NCCL."""
    from bluerov2_dynamics_b200 import dist as D
    eng = B.Engine("thruster8", "f64", device=local)
    g = torch.Generator(device=eng.device).manual_seed(4)
    U = (torch.rand((T, 8), device=eng.device, dtype=torch.float64, generator=g) * 0.8 - 0.4)
    # "recorded" states: a smooth bounded synthetic series in a tank-sized box plus sensor-like noise
    t = torch.arange(T, device=eng.device, dtype=torch.float64) * DT
    X = torch.zeros((T, 12), device=eng.device, dtype=torch.float64)
    X[:, 0] = 2.0 * torch.sin(0.05 * t); X[:, 1] = 2.0 * torch.cos(0.04 * t); X[:, 2] = 1.5 + torch.sin(0.03 * t)
    X[:, 5] = 0.5 * torch.sin(0.02 * t)
    X[:, 6] = 0.1 * torch.cos(0.05 * t); X[:, 7] = -0.08 * torch.sin(0.04 * t); X[:, 8] = 0.03 * torch.cos(0.03 * t)
    X += 1e-3 * torch.randn((T, 12), device=eng.device, dtype=torch.float64, generator=g)
    hs = list(horizons)
    lo, hi, nloc = D.window_shard(T, hs, rank, world)
    Xl, Ul = X[lo:hi].contiguous(), U[lo:hi].contiguous()

    def one():
        se, _ = eng.multistep_se(Xl, Ul, hs, dt=DT, integrator="rk4", n_windows=nloc)
        vec = se[:len(hs)].clone()
        D.allreduce_sum_(vec)
        return vec

    one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    vec = one()
    e1.record()
    torch.cuda.synchronize()
    windows.append((t0, time.perf_counter()))
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=eng.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    cnt = D.global_counts(T, hs)
    # window k runs min(Hmax, T-1-k) steps; all horizons are read off the same rollout
    hm = hs[-1]
    vsteps = float(max(T - hm, 0) * hm + hm * (hm - 1) // 2) if T > hm else float(T * (T - 1) // 2)
    rm = [float(np.sqrt(v / (c * 12))) for v, c in zip(vec.cpu().numpy(), cnt)]
    return dict(workload=f"cfg5: T={T} rows, H={hs}, RK4, 8-thruster fp64, windows sharded over {world} GPU(s), "
                         "NCCL all-reduce of the SE vector" if world > 1 else
                         f"cfg5: T={T} rows, H={hs}, RK4, 8-thruster fp64", windows=cnt, ms=ms,
                vehicle_steps_per_s=vsteps / (ms * 1e-3), rmse=rm)


def run_reduced9_leg(torch, B, local, peaks, peak_src, steps, warmup, windows, rows=1 << 24, traffic=None):
    """bluerov_compute (fossen/bluerov_torch.py:20-67) batched over `rows` states: 13 scalars in, 9 out per row —
    38 FLOP per 88 B in fp32, HBM-bound.  Two input/output sets (1.48 GB each, far larger than L2) used alternately."""
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(9)
    X = [torch.randn((rows, 9), device=dev, dtype=torch.float32, generator=g) for _ in range(2)]
    U = [torch.randn((rows, 4), device=dev, dtype=torch.float32, generator=g) for _ in range(2)]
    O = [torch.empty((rows, 9), device=dev, dtype=torch.float32) for _ in range(2)]
    for k in range(max(warmup, 3)):
        B.reduced9_rhs(X[k % 2], U[k % 2], out=O[k % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(10, min(steps, 50))
    t0 = time.perf_counter()
    e0.record()
    for k in range(n):
        B.reduced9_rhs(X[k % 2], U[k % 2], out=O[k % 2])
    e1.record()
    torch.cuda.synchronize()
    windows.append((t0, time.perf_counter()))
    ms = e0.elapsed_time(e1) / n
    nbytes = rows * (9 + 4 + 9) * 4
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"workload": f"bluerov_compute RHS, {rows} rows fp32, 88 B algorithmic per row", "ms": ms,
            "evals_per_s": rows / (ms * 1e-3), "launches": n,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": (traffic or {}).get("bytes"),
                         "traffic_source": (traffic or {}).get("capture"), "kernel": "brov::reduced9_kernel<float>",
                         "bytes_per_launch": nbytes, "peak_source": peak_src}}


def run_compare_leg(torch, B, local, windows, fp32_peak, fp64_peak, T=1_000_100, horizons=(1, 10, 100), cpu=True):
    """The reference's model-comparison table (training/train_tank_brov2_full_comparison.py:977-1009: endpoint RMSE at
    H = 1/10/100 of the Fossen, Koopman, double-integrator and PINc models) on one synthetic 50 Hz series of T rows:
    device time of each evaluator, and the oracle timed on a bounded prefix of the same series on the host cores."""
    from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc
    from bluerov2_dynamics_b200 import pinc as P
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(4)
    U = (torch.rand((T, 8), device=dev, dtype=torch.float64, generator=g) * 0.8 - 0.4)
    t = torch.arange(T, device=dev, dtype=torch.float64) * DT
    X = torch.zeros((T, 12), device=dev, dtype=torch.float64)
    X[:, 0] = 2.0 * torch.sin(0.05 * t); X[:, 1] = 2.0 * torch.cos(0.04 * t); X[:, 2] = 1.5 + torch.sin(0.03 * t)
    X[:, 5] = 0.5 * torch.sin(0.02 * t)
    X[:, 6] = 0.1 * torch.cos(0.05 * t); X[:, 7] = -0.08 * torch.sin(0.04 * t); X[:, 8] = 0.03 * torch.cos(0.03 * t)
    X += 1e-3 * torch.randn((T, 12), device=dev, dtype=torch.float64, generator=g)
    hs = list(horizons)
    rng = np.random.default_rng(12)
    out = {"workload": f"comparison table: endpoint RMSE at H={hs} over all windows of a {T}-row synthetic series",
           "models": {}}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        windows.append((t0, time.perf_counter()))
        return e0.elapsed_time(e1) / reps, r

    nwin = [T - h for h in hs]
    steps_all = float(sum((T - h) * h for h in hs))
    # double integrator (RK4, 8 inputs): one pass, all horizons
    e = B.Engine("di12_u8", "f64", device=local)
    K_lin, K_ang = rng.normal(0, 0.05, (8, 3)), rng.normal(0, 0.05, (8, 3))
    e.set_di_gains(K_lin, K_ang)
    ms, _ = timed(lambda: e.multistep_se(X, U, hs, dt=DT, integrator="rk4")[0])
    out["models"]["double_integrator_rk4_f64"] = {"ms": ms, "windows": nwin, "kernel": "brov::se_kernel<double, DI12_U8, RK4>"}
    # Koopman EDMDc, the reference's configuration: 500 RBFs -> d = 512
    k = 500
    Kc = X[torch.randint(0, T, (k,), device=dev, generator=g)].cpu().numpy()
    A = 0.98 * np.linalg.qr(rng.standard_normal((12 + k, 12 + k)))[0]
    Bm = 0.02 * rng.standard_normal((12 + k, 8))
    KM = KoopmanEDMDc(state_dim=12, input_dim=8, n_rbfs=k, gamma=3.0)
    KM.centers_, KM.A_, KM.B_ = Kc, A, Bm
    h_k = KM._handle()
    from bluerov2_dynamics_b200 import _lib as L

    arr_h = (__import__("ctypes").c_int * len(hs))(*hs)
    se_k = torch.zeros(len(hs), device=dev, dtype=torch.float64)

    def koop_all():   # all horizons from ONE lift pass (brov_koopman_multistep_se_multi)
        L.check(L.lib.brov_koopman_multistep_se_multi(h_k, X.data_ptr(), U.data_ptr(), T, len(hs), arr_h, se_k.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream))
        return se_k

    def koop_each():  # one call per horizon (brov_koopman_multistep_se), the lift repeated for each
        for i, h in enumerate(hs):
            L.check(L.lib.brov_koopman_multistep_se(h_k, X.data_ptr(), U.data_ptr(), T, T - h, h,
                                                    se_k.data_ptr() + 8 * i, torch.cuda.current_stream().cuda_stream))
        return se_k
    ms_each, _ = timed(koop_each)
    ms, _ = timed(koop_all)
    # algorithmic flop of the multi-horizon formulation: the lift (k RBFs: 24-flop distance + exp counted as 2) once per
    # window, then per horizon the decoder rows (2 n d) and the input FIR (2 n r H)
    flop_k = (T - hs[0]) * k * 26.0 + sum((T - h) * (2.0 * 12 * (12 + k) + 2.0 * 12 * 8 * h) for h in hs)
    out["models"]["koopman_d512_f64"] = {
        "ms": ms, "ms_one_launch_per_horizon": ms_each, "windows": nwin,
        "kernel": "koop_liftw_kernel<12, 3> + koop_fir_se_kernel<12, 8>", "algorithmic_tflops": flop_k / (ms * 1e-3) / 1e12,
        "frac_fp64_pipe": flop_k / (ms * 1e-3) / 1e12 / fp64_peak,
        "dense_equivalent_tflops": sum((T - h) * h * 2.0 * (12 + k) ** 2 for h in hs) / (ms * 1e-3) / 1e12,
        "note": "decoder-row formulation: 2 n (d + r H) + lift flop per window instead of the reference's 2 d^2 H"}
    # PINc with the reference's trained checkpoint (frozen in tests/golden)
    cg = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_cmp.npz"))
    PM = P.PincModel({kk[len("pinc_sd_"):]: cg[kk] for kk in cg.files if kk.startswith("pinc_sd_")}, device=local)
    ms, _ = timed(lambda: PM.multistep_se(X, U, hs, DT, "reset")[0], reps=2)
    hm = hs[-1]   # every window runs min(hm, rows left) network steps; the shorter horizons are read off on the way
    steps_p = float((T - hm) * hm + hm * (hm - 1) // 2)
    flop_p = steps_p * 2.0 * (14 * 64 + 3 * 64 * 64 + 64 * 9)
    out["models"]["pinc_f32"] = {"ms": ms, "windows": nwin, "kernel": "pinc_se_kernel", "network_steps": steps_p,
                                 "algorithmic_tflops": flop_p / (ms * 1e-3) / 1e12,
                                 "frac_fp32_pipe": flop_p / (ms * 1e-3) / 1e12 / fp32_peak}
    if cpu:
        from oracle import compare_np as CN
        Tc = 2100
        Xc, Uc = X[:Tc].cpu().numpy(), U[:Tc].cpu().numpy()
        t0 = time.perf_counter(); CN.di_multistep_se("di12", "rk4", Xc, Uc, hs, DT, K_lin, K_ang); t_di = time.perf_counter() - t0
        t0 = time.perf_counter()
        for h in hs:
            CN.koop_multistep_se(Xc, Uc, h, Kc, 3.0, A, Bm)
        t_k = time.perf_counter() - t0
        layers = CN.pinc_weights(cg)
        t0 = time.perf_counter(); CN.pinc_multistep_se(Xc, Uc, hs, DT, layers, "reset"); t_p = time.perf_counter() - t0
        scale = steps_all / float(sum((Tc - h) * h for h in hs))
        out["cpu_oracle"] = {"rows": Tc, "kind": "port", "note": "numpy oracle vectorised over windows (already a stronger baseline than "
                             "the reference's per-window Python loops), seconds on this prefix and extrapolated to T rows",
                             "double_integrator_s": t_di, "koopman_s": t_k, "pinc_s": t_p,
                             "extrapolated_s": {"double_integrator": t_di * scale, "koopman": t_k * scale, "pinc": t_p * scale}}
    return out


def run_monte_carlo_leg(torch, B, local, fp32_peak, windows, n=1 << 20, steps=5):
    """BASELINE configs[3]: Monte-Carlo sweep — per-vehicle added-mass / damping coefficients (table [36][N], staged per
    block into shared memory) with wrench input and the first-order wrench lag, fp32, 100 RK4 steps per launch."""
    e = B.Engine("wrench12", "f32", device=local)
    rng = np.random.default_rng(3)
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 9:27] *= rng.uniform(0.7, 1.3, (n, 18))
    ph[:, 27:30] = 1.0 / (ph[:, 0:1] - ph[:, 9:12])
    ph[:, 30:33] = 1.0 / (ph[:, 6:9] - ph[:, 12:15])
    ph[:, 36] = rng.uniform(0.05, 0.3, n)
    e.set_wrench_lag1(True)
    e.set_vehicle_physical(ph)
    g = torch.Generator(device=e.device).manual_seed(33)
    scale = torch.tensor([40, 40, 40, 5, 5, 5.0], device=e.device)
    U = [((torch.rand((CHUNK, n, 6), device=e.device, generator=g) * 2 - 1) * scale).contiguous() for _ in range(2)]
    x = torch.zeros((n, 12), device=e.device, dtype=torch.float32)
    lag = torch.zeros((n, 6), device=e.device, dtype=torch.float32)

    def one(k):
        e.rollout(x, U[k % 2], dt=DT, integrator="rk4", lag0=lag, xT_out=x, lag_out=lag, step0=k * CHUNK)
    for k in range(3):
        one(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(steps):
        one(3 + k)
    e1.record()
    torch.cuda.synchronize()
    windows.append((t0, time.perf_counter()))
    ms = e0.elapsed_time(e1) / steps
    rate = n * CHUNK / (ms * 1e-3)
    flop = FLOP_PER_STEP["wrench12"] + 4 * 12.0     # + first-order lag: 12 flop per RHS evaluation
    return {"workload": f"configs[3]: {n} vehicles, per-vehicle coefficients + first-order wrench lag, wrench-input "
                        "12-state model, fp32, 100 RK4 steps per launch",
            "ms": ms, "vehicle_steps_per_s": rate, "finite": bool(torch.isfinite(x).all().item()),
            "roofline": {"bound": "fp32_pipe", "achieved": rate * flop / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": rate * flop / 1e12 / fp32_peak, "flop_per_vehicle_step": flop,
                         "kernel": "brov::rollout_kernel<float, WRENCH12, RK4, lag1, per-vehicle>"}}


def main_ours(args):
    import torch
    import torch.distributed as dist
    import bluerov2_dynamics_b200 as B
    from bluerov2_dynamics_b200 import dist as D

    rank, world, local = D.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"[bench] note: WORLD_SIZE={world} but --gpus {args.gpus}; using {world}", file=sys.stderr)
    torch.cuda.set_device(local)
    cores = host_cores()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before the timed GPU region; forked workers never touch CUDA
        spc = args.cpu_steps
        rate, wall = cpu_reference_rate(cores, spc)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} processes x 1 vehicle x {spc} RK4 steps (thruster model, dt=0.02), numpy float64 "
                         f"oracle executed one vehicle per call as the reference does; wall {wall:.1f} s",
               "batched_numpy_value": cpu_batched_rate(),
               "batched_numpy_sample": "same oracle vectorised over 4096 vehicles x 20 RK4 steps in one process"}
        c_rate, c_thr = cpu_c_oracle_rate()
        if c_rate is not None:
            cpu["c_port_value"] = c_rate
            cpu["c_port_threads"] = c_thr
            cpu["c_port_sample"] = ("oracle/brov_oracle.c (scalar float64 C restatement, gcc -O2, OpenMP over vehicles): "
                                    "8192 vehicles x 400 RK4 steps")

    sampler = ClockSampler(local) if rank == 0 else None
    windows = []
    peaks, peak_src = measured_peaks()

    # in-run FP pipe peaks (MEASURED_PEAKS.json carries only HBM and bf16-GEMM numbers)
    fp64_peak, _ = B.fma_peak("f64", local, 2048)
    fp32_peak, _ = B.fma_peak("f32", local, 8192)

    leg2 = run_rollout_leg(torch, dist, B, CFG2, args.steps, args.warmup, local, world, windows)
    e2e = run_e2e_leg(torch, dist, B, CFG2, leg2, min(args.steps, args.e2e_steps), args.warmup, world, windows)
    del leg2["U"], leg2["x0"]
    torch.cuda.empty_cache()
    leg3 = run_rollout_leg(torch, dist, B, CFG3, args.steps, args.warmup, local, world, windows)
    del leg3["U"], leg3["x0"], leg3["eng"]
    torch.cuda.empty_cache()
    rmse = run_rmse_leg(torch, dist, B, local, rank, world, windows) if not args.no_rmse else None
    if rmse is not None:
        tf = rmse["vehicle_steps_per_s"] / world * FLOP_PER_STEP["thruster8"] / 1e12   # per GPU
        rmse["roofline"] = {"bound": "fp64_pipe", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": tf / fp64_peak, "kernel": "brov::se_kernel<double, THRUSTER8, RK4>",
                            "note": "per GPU; at N > 1 the timed region includes the NCCL all-reduce of the SE vector"}
    mc = run_monte_carlo_leg(torch, B, local, fp32_peak, windows) if rank == 0 and not args.no_compare else None
    compare = (run_compare_leg(torch, B, local, windows, fp32_peak, fp64_peak, cpu=not args.no_cpu_baseline)
               if rank == 0 and not args.no_compare else None)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except OSError:
        traffic = {}
    red9 = run_reduced9_leg(torch, B, local, peaks, peak_src, args.steps, args.warmup, windows,
                            traffic=traffic.get("reduced9_f32")) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    clocks = sampler.stop(windows)

    def roof(leg, cfg, peak):
        per_gpu_steps = leg["vehicle_steps"] / world
        tr = traffic.get("rollout_" + cfg["dtype"], {})
        tf = FLOP_PER_STEP[cfg["model"]] * per_gpu_steps / (leg["ms_total"] * 1e-3) / 1e12
        gbs = leg["bytes_per_launch"] / (leg["ms_per_step"] * 1e-3) / 1e9
        return {"bound": "fp64_pipe" if cfg["dtype"] == "f64" else "fp32_pipe", "achieved": tf, "peak": peak,
                "unit": "TFLOP/s", "frac": tf / peak, "traffic": tr.get("bytes"), "traffic_source": tr.get("capture"),
                "kernel": f"brov::rollout_kernel<{'double' if cfg['dtype'] == 'f64' else 'float'}, THRUSTER8, RK4>",
                "flop_per_vehicle_step": FLOP_PER_STEP[cfg["model"]],
                "peak_source": "in-run FMA-chain microbenchmark (brov_fma_peak), 2 flop per FMA",
                "kernel_ms": leg["ms_per_step"],
                "hbm": {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                        "bytes_per_launch": leg["bytes_per_launch"], "peak_source": peak_src}}

    v2 = leg2["vehicle_steps"] / (leg2["ms_total"] * 1e-3)
    v3 = leg3["vehicle_steps"] / (leg3["ms_total"] * 1e-3)
    line = {
        "metric": METRIC, "value": v2, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": leg2["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 65,536-vehicle ensemble per GPU, 8-thruster Fossen model with "
                               "3rd-order thruster lag, per-vehicle random thrust inputs U(-0.4,0.4), RK4, dt=0.02, "
                               f"{args.steps * CFG2['chunk']} steps ({CFG2['chunk']} per launch), fp64",
                   "vehicles_per_gpu": CFG2["n_per_gpu"], "rk4_steps_per_launch": CFG2["chunk"],
                   "l2_policy": "inputs larger than L2: two 4.19 GB input chunks used alternately, read once per launch",
                   "parallelism": f"vehicle-sharded x{world}, no data-path collective"},
        "roofline": roof(leg2, CFG2, fp64_peak),
        "e2e": e2e,
        "gpu_launches": args.steps,
        "clocks": clocks,
        "finite": leg2["finite"] and leg3["finite"],
        "fp32": {"value": v3, "unit": UNIT, "ms_per_step": leg3["ms_per_step"], "dtype": "f32",
                 "config": {"workload": "BASELINE configs[2]: 1,048,576-vehicle ensemble per GPU, same model, fp32, "
                                        f"{args.steps * CFG3['chunk']} RK4 steps ({CFG3['chunk']} per launch), trajectory writeback every 10 steps",
                            "vehicles_per_gpu": CFG3["n_per_gpu"],
                            "l2_policy": "two 3.36 GB input chunks used alternately"},
                 "roofline": roof(leg3, CFG3, fp32_peak), "gpu_launches": args.steps},
        "rmse": rmse,
        "reduced9": red9,
        "monte_carlo": mc,
        "comparison_models": compare,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main_reference(args):
    """Reference arm: the reference's algorithm (oracle port; the reference itself is Python under /root/reference and
    does not exist on the GPU box) on all host cores.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    spc = args.cpu_steps
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(cores, max(spc // 4, 10))
    rates, t0 = [], time.perf_counter()
    steps_done = 0
    for k in range(args.steps):
        r, _ = cpu_reference_rate(cores, spc, seed=100 * k)
        rates.append(r)
        steps_done += 1
        if time.perf_counter() - t0 > args.reference_budget_s:
            break
    wall = time.perf_counter() - t0
    value = cores * spc * steps_done / wall
    sample = (f"each step: {cores} processes x 1 vehicle x {spc} RK4 steps of configs[1]'s model (8-thruster + lag, "
              f"dt=0.02, random thrust), numpy float64, one vehicle per Python call as the reference executes; "
              f"{steps_done} of {args.steps} steps run within the {args.reference_budget_s:.0f} s budget")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps_done, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * wall / max(steps_done, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] model and inputs, bounded sample on the host CPU cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-steps", type=int, default=1500,
                    help="RK4 steps per core of the CPU baseline sample (1500 = about 25 core-seconds on 16 cores)")
    ap.add_argument("--reference-budget-s", type=float, default=120.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rmse", action="store_true")
    ap.add_argument("--no-compare", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
