"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/brov.h declares, the host-only
entry points (constants, geometry, zero-order hold) agree with the reference's golden values, argument checking
reports errors, and the multi-rank evaluator plumbing works over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, normwise


@pytest.fixture(scope="module")
def L():
    """ctypes binding; builds libbrov.so first if the checkout is fresh (nvcc cross-compiles without a GPU)."""
    from bluerov2_dynamics_b200.build import build_lib
    build_lib()
    from bluerov2_dynamics_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "brov.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(brov_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    raw = C.CDLL(L.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    assert declared == set(L._PROTOS), declared ^ set(L._PROTOS)
    assert L.lib.brov_abi_version() == L.ABI_VERSION
    # struct layouts match the header's field order and sizes (LP64)
    gen = 4 + 4 + 8 + 8 + 8 * 3 + 8 * 8 + 8 + 8
    assert C.sizeof(L.InputGen) == gen
    assert C.sizeof(L.RolloutDesc) == 4 + 4 + 8 * 14 + 4 + 4 + 4 + 4 + 8 * 3 + gen
    assert C.sizeof(L.GenInputsDesc) == 4 * 4 + 8 * 5 + gen + 8
    assert C.sizeof(L.SeDesc) == 4 + 4 + 8 * 3 + 8 * 3 + 4 + 4 * 4 + 4 + 8 * 4 + 4 + 4 + 8 + 8 + 8 + 8
    assert C.sizeof(L.RolloutHostDesc) == 4 + 4 + 8 * 3 + 8 * 3 + 4 + 4 + 8 * 5 + 4 + 4 + gen + 8 + 8
    assert C.sizeof(L.PincWeights) == 4 * 4 + 5 * 8 + 5 * 8 + 4 * 4 + 4 * 8 + 4 * 8
    assert C.sizeof(L.PincRolloutDesc) == 4 + 4 + 8 * 11
    assert C.sizeof(L.PincSeDesc) == 4 + 4 + 4 * 4 + 8 * 5 + 4 + 4 + 8 * 3


def test_header_compiles_as_c_and_layouts_match_the_ctypes_mirrors(L, tmp_path):
    """include/brov.h is plain C: gcc compiles it, and sizeof / offsetof of every descriptor a binding has to mirror
    agree with the ctypes structures of _lib.py (the binding INTEGRATION.md shows)."""
    pairs = [("brov_input_gen", L.InputGen), ("brov_rollout_desc", L.RolloutDesc),
             ("brov_gen_inputs_desc", L.GenInputsDesc), ("brov_se_desc", L.SeDesc),
             ("brov_rollout_host_desc", L.RolloutHostDesc), ("brov_pinc_weights", L.PincWeights),
             ("brov_pinc_rollout_desc", L.PincRolloutDesc), ("brov_pinc_se_desc", L.PincSeDesc)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "brov.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("abi %d\\n", BROV_ABI_VERSION);', "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["abi"]) == L.ABI_VERSION
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_constants_match_reference(L, golden):
    from bluerov2_dynamics_b200.engine import default_allocation, default_physical, derive_params, lag_discretize
    ph = default_physical()
    assert ph[L.PH_W] == golden["const_W"] and ph[L.PH_B] == golden["const_B"]
    assert np.allclose(ph[L.PH_MINV:L.PH_MINV + 6], golden["const_Minv_diag"], rtol=1e-15)
    alloc, r, d = default_allocation()
    assert np.allclose(alloc, golden["const_alloc"], atol=1e-16)
    assert np.allclose(r, golden["const_thr_r"], atol=1e-16) and np.allclose(d, golden["const_thr_dir"], atol=1e-16)
    for dt in (0.01, 0.02, 0.05):
        Ad, Bd = lag_discretize(dt)
        assert np.max(np.abs(Ad - golden[f"const_lag_Ad_{dt}"])) < 2e-15
        assert np.max(np.abs(Bd - golden[f"const_lag_Bd_{dt}"])) < 2e-15
    kp = derive_params(ph)
    assert kp.shape == (L.NKP,)
    assert np.isclose(kp[27], 132.57 - 131.588) and np.allclose(kp[6:9], [19.86, 20.62, 32.18])
    assert np.allclose(kp[15:21], [13.7, 0, 33, 0, 0.8, 0]) and np.allclose(kp[21:27], [141, 217, 190, 1.19, 0.47, 1.5])
    batch = derive_params(np.tile(ph, (5, 1)))
    assert batch.shape == (5, L.NKP) and np.array_equal(batch[3], kp)


def test_errors_are_reported_without_a_gpu(L):
    rc = L.lib.brov_default_physical(1000.0, None)
    assert rc == -1 and b"NULL" in L.lib.brov_last_error()
    Ad, Bd = np.zeros(9), np.zeros(3)
    assert L.lib.brov_lag_discretize(-1.0, L.dptr(Ad), L.dptr(Bd)) == -1
    with pytest.raises(L.BrovError):
        L.check(L.lib.brov_lag_discretize(float("nan"), L.dptr(Ad), L.dptr(Bd)))
    h = C.c_void_p()
    rc = L.lib.brov_create(7, L.F64, 0, C.byref(h))  # unknown model: rejected before any CUDA call
    assert rc == -1 and not h.value
    # comparison models: argument errors are reported before any CUDA call, too
    z = np.zeros(4)
    assert L.lib.brov_koopman_create(0, 99, 8, 4, 1.0, L.dptr(z), L.dptr(z), L.dptr(z), C.byref(h)) == -1
    assert b"unsupported dimensions" in L.lib.brov_last_error() and not h.value
    w = L.PincWeights()
    w.struct_size = 1
    assert L.lib.brov_pinc_create(0, C.byref(w), C.byref(h)) == -1 and b"size mismatch" in L.lib.brov_last_error()
    w.struct_size, w.n_hidden_layers, w.hidden = C.sizeof(L.PincWeights), 3, 64
    assert L.lib.brov_pinc_create(0, C.byref(w), C.byref(h)) == -4 and not h.value


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    """No GPU here: constructing an engine must fail loudly, and nothing under bluerov2_dynamics_b200/ may mention
    the oracle."""
    import torch
    pkg = os.path.join(ROOT, "bluerov2_dynamics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
    if not torch.cuda.is_available():
        import bluerov2_dynamics_b200 as B
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            B.Engine("thruster8", "f64")


def test_parameters_module_matches_reference_values():
    from bluerov2_dynamics_b200.fossen import parameters as P
    from oracle.fossen_np import RED
    for k, v in RED.items():
        assert getattr(P, k) == v, k
    assert P.F_bouy == 1026 * 0.0115 * 9.82 and P.z_b == -0.1 and P.I_xx == 0.21


def test_shard_helpers():
    from bluerov2_dynamics_b200 import dist as D
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = [D.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    lo, hi, nl = D.window_shard(140, [1, 10, 100], 1, 2)
    assert (lo, hi, nl) == (70, 140, 69)
    assert D.global_counts(140, [1, 10, 100]) == [139, 130, 40]
    assert D.global_counts(5, [10]) == [0]


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import torch.distributed as dist
from bluerov2_dynamics_b200 import dist as D
from oracle import fossen_np as O
rank, world, _ = D.init_from_env("gloo")
g = dict(np.load(os.path.join({root!r}, "tests", "golden", "reference_vectors.npz")))
X, U, HS = g["rmse_X12"], g["rmse_W6"], [1, 10, 100]
m = O.Model("wrench12", 0.02)
def local_se(Xr, Ur, hs, nwin):
    # per-rank squared-error sums from the oracle: windows [0, nwin) of the rank's rows
    out = []
    for h in hs:
        ns = min(nwin, len(Xr) - h)
        if ns <= 0:
            out.append(0.0); continue
        x = Xr[:ns].copy(); lag = m.zero_lag(ns)
        for j in range(h):
            x, lag = O.step(m, "euler", x, Ur[j:j + ns], lag)
        e = x - Xr[h:h + ns]
        out.append(float(np.sum(e * e)))
    return torch.tensor(out, dtype=torch.float64)
got = D.sharded_multistep_rmse(local_se, X, U, HS, rank, world)
if rank == 0:
    assert np.allclose(got, g["rmse_w12_euler"], rtol=1e-11), (got, g["rmse_w12_euler"])
    print("SHARDED_OK", got)
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_evaluator_world_size_2_gloo(tmp_path, golden):
    """N > 1 path on CPU: two gloo ranks shard the sliding windows (with the H-row halo), the oracle supplies each
    rank's squared-error sums, one all-reduce combines them; the result equals the reference's RMSE."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHARDED_OK" in r.stdout


def test_bench_reference_arm_prints_contract_line():
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-steps", "20"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    # the unmodified reference from oracle/_ref when it has been built (kind "reference"), else the labelled numpy port
    from oracle import ref_loader
    want = "reference" if ref_loader.available() else "port"
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == want
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "vehicle-steps/s"


def test_dataset_wire_format_roundtrip(tmp_path):
    """load_dataset: the reference's CSV contract (sort by t, dedupe, drop non-finite states, zero-fill inputs)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("brov_datasets", os.path.join(ROOT, "bluerov2_dynamics_b200", "datasets.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)   # loaded standalone: the package itself refuses to import without libbrov.so + GPU use
    import pandas as pd
    rng = np.random.default_rng(0)
    X, U = rng.normal(size=(50, 12)), rng.uniform(-1, 1, (50, 8))
    p = tmp_path / "ds.csv"
    D.save_dataset(p, X, U, 0.02)
    X2, U2, dt = D.load_dataset(p, verbose=False)
    assert np.allclose(X2, X) and np.allclose(U2, U) and abs(dt - 0.02) < 1e-12
    df = pd.read_csv(p)
    df = pd.concat([df.iloc[::-1], df.iloc[:3]])          # shuffled order + duplicate stamps
    df.loc[df.index[5], "x"] = np.inf                       # a non-finite state row is dropped
    df = df.drop(columns=["u7", "u8"])                      # missing input columns are zero-filled
    df.to_csv(p, index=False)
    X3, U3, dt3 = D.load_dataset(p, verbose=False)
    assert len(X3) == 49 and np.all(np.diff(pd.read_csv(p).sort_values("t")["t"].unique()) > 0)
    assert np.all(U3[:, 6:] == 0) and abs(dt3 - 0.02) < 1e-9
    Xq, Uw, _ = D.load_dataset(p, inputs="wrench", quaternion=True, verbose=False)   # legacy Euler file -> quaternions
    assert Xq.shape[1] == 13 and Uw.shape[1] == 6 and np.allclose(np.linalg.norm(Xq[:, 3:7], axis=1), 1.0)
    with pytest.raises(ValueError):
        D.load_dataset(p, inputs="voltages")
    pd.read_csv(p).drop(columns=["t"]).to_csv(p, index=False)
    with pytest.raises(ValueError):
        D.load_dataset(p, verbose=False)


def test_sim_generator_random_stream_matches_reference(golden):
    import importlib.util
    spec = importlib.util.spec_from_file_location("brov_datasets", os.path.join(ROOT, "bluerov2_dynamics_b200", "datasets.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    U, nz = D.smooth_random_inputs(1500, seed=42)
    assert np.array_equal(U, golden["simgen_inputs"])
    scale = np.repeat([0.0005, 0.001, 0.0005, 0.001], 3)
    noisy = golden["simgen_states_true_s10"] + (nz * scale)[::10]
    assert np.allclose(noisy, golden["simgen_states_noisy_s10"], rtol=0, atol=1e-15)


def test_window_shards_cover_every_window_once_and_hold_their_rows():
    """dist.window_shard / window_shard_carry (pure arithmetic): for random series lengths, horizons, world sizes and
    replay depths the ranks' windows partition 0..T-H-1, every rank holds the rows its windows read (H-row halo) and,
    in carry mode, the rows its first window's replay reaches back to — the condition brov_multistep_se enforces
    (`row0 <= max(0, (window0 * H - carry_steps) / H)`)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("brov_dist", os.path.join(ROOT, "bluerov2_dynamics_b200", "dist.py"))
    D = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(D)
    rng = np.random.default_rng(5)
    for _ in range(300):
        T = int(rng.integers(1, 5000))
        world = int(rng.integers(1, 9))
        hs = sorted(set(int(h) for h in rng.integers(1, 120, size=int(rng.integers(1, 4)))))
        depth = int(rng.integers(1, 250))
        # reset mode: one shard for all horizons (windows counted for the shortest one)
        nwin = max(T - hs[0], 0)
        seen = 0
        for r in range(world):
            lo, hi, n = D.window_shard(T, hs, r, world)
            assert n >= 0 and lo == seen if n else True
            if n:
                assert hi == min(T, lo + n + hs[-1]) and hi <= T
                seen = lo + n
        assert seen == nwin or nwin == 0
        # carry mode: one shard per horizon
        for H in hs:
            nw = max(T - H, 0)
            nxt = 0
            for r in range(world):
                row_lo, row_hi, n, w0 = D.window_shard_carry(T, H, r, world, depth)
                if not n:
                    continue
                assert w0 == nxt
                nxt = w0 + n
                assert row_hi == min(T, w0 + n + H) and row_lo <= w0
                assert row_lo <= max(0, (w0 * H - depth) // H)          # the C ABI's replay-history condition
                assert row_lo >= 0
            assert nxt == nw
