"""Builds tuning variants of libbrov.so (compile-time knobs of csrc/brov_kernels.cuh / brov_device.cuh) HERE, and on a
GPU times a matrix of rollout kernels with each.  Usage:  python profiles/tune_variants.py build | run [names...]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# name -> list of -D defines.  The knobs measured in round 2 (input path per kernel, register caps, nu_dot form, ring
# depth, Philox rounds, per-vehicle table placement) are decided and gone from the sources; their timings are kept in
# profiles/r02[d-t]_tune_variants.txt.  Add a knob to the sources and an entry here to measure the next one.
VARIANTS = {}
VDIR = os.path.join(ROOT, "bluerov2_dynamics_b200", "variants")


def build(names):
    from bluerov2_dynamics_b200.build import build_lib
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        if names and name not in names:
            continue
        out = os.path.join(VDIR, f"libbrov_{name}.so")
        t = time.time()
        build_lib(force=True, defines=defs, out=out, tag="_" + name)
        print(name, f"{time.time() - t:.0f}s", flush=True)


WORKER = r'''
import os, sys, json, torch, numpy as np
sys.path.insert(0, %r)
import bluerov2_dynamics_b200 as B
res = {}
rng = np.random.default_rng(3)

def timeit(fn, reps=10):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def mc_table(n):
    ph = np.tile(B.default_physical(), (n, 1))
    ph[:, 9:27] *= rng.uniform(0.7, 1.3, (n, 18))
    ph[:, 27:30] = 1.0 / (ph[:, 0:1] - ph[:, 9:12]); ph[:, 30:33] = 1.0 / (ph[:, 6:9] - ph[:, 12:15])
    ph[:, 36] = rng.uniform(0.05, 0.3, n)
    return ph

def case(tag, model, dtype, n, T, mode, mc=False, stride=0, lag_repr="projected"):
    if not tag.startswith(os.environ.get("BROV_CASES", "")):
        return
    e = B.Engine(model, dtype)
    if mc:
        e.set_wrench_lag1(True); e.set_vehicle_physical(mc_table(n))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.zeros((n, e.nx), device="cuda", dtype=e.tdtype)
    if e.nx == 13: x[:, 3] = 1.0
    nl = (18 if lag_repr == "projected" else 24) if model == "thruster8" else (6 if mc else 0)
    lag = torch.zeros((n, nl), device="cuda", dtype=e.tdtype) if nl else None
    traj = torch.empty((T // stride, n, e.nx), device="cuda", dtype=e.tdtype) if stride else None
    sc = None if e.nu == 8 else [40, 40, 40, 5, 5, 5.0]
    if mode == "gen":
        gen = B.InputGenerator(seed=5, scale=sc)
        gs = torch.zeros((n, e.nu), device="cuda", dtype=e.tdtype)
        fn = lambda k: e.rollout(x, gen=gen, steps=T, step0=k * T, dt=0.02, lag0=lag, xT_out=x, lag_out=lag, lag_repr=lag_repr,
                                 gen_state=gs, gen_state_out=gs, stride=stride, traj_out=traj)
    else:
        amp = torch.tensor([0.4] * 8 if e.nu == 8 else [40, 40, 40, 5, 5, 5.0], device="cuda", dtype=e.tdtype)
        U = [((torch.rand((T, n, e.nu), device="cuda", dtype=e.tdtype, generator=g) * 2 - 1) * amp).contiguous() for _ in range(2)]
        fn = lambda k: e.rollout(x, U[k %% 2], dt=0.02, lag0=lag, xT_out=x, lag_out=lag, lag_repr=lag_repr, stride=stride,
                                 traj_out=traj, step0=k * T)
    ms = timeit(fn, 5 if T >= 1000 else 10)
    res[tag] = {"ms": round(ms, 4), "gsteps": round(n * T / ms / 1e6, 3), "finite": bool(torch.isfinite(x).all())}
    print(f"  {tag:26s} {ms:9.4f} ms  {n * T / ms / 1e6:8.2f}e9 steps/s", file=sys.stderr, flush=True)

N64, N32 = 1 << 16, 1 << 20
case("thr_f64_tma_100", "thruster8", "f64", N64, 100, "tma")
case("thr_f64_tma_1000", "thruster8", "f64", N64, 1000, "tma")
case("thr_f64_gen_1000", "thruster8", "f64", N64, 1000, "gen")
case("thr_f64_tma_1000_lag24", "thruster8", "f64", N64, 1000, "tma", lag_repr="thruster")
case("w12_f64_tma_100", "wrench12", "f64", N64, 100, "tma")
case("w12_f64_tma_1000", "wrench12", "f64", N64, 1000, "tma")
case("w12_f64_gen_1000", "wrench12", "f64", N64, 1000, "gen")
case("q13_f64_tma_100", "quat13", "f64", N64, 100, "tma")
case("q13_f64_tma_1000", "quat13", "f64", N64, 1000, "tma")
case("w12_f64_mc_100", "wrench12", "f64", N64, 100, "tma", mc=True)
case("thr_f32_tma_100_s10", "thruster8", "f32", N32, 100, "tma", stride=10)
case("thr_f32_gen_100_s10", "thruster8", "f32", N32, 100, "gen", stride=10)
case("w12_f32_tma_100", "wrench12", "f32", N32, 100, "tma")
case("w12_f32_mc_100", "wrench12", "f32", N32, 100, "tma", mc=True)
case("q13_f32_mc_100", "quat13", "f32", N32, 100, "tma", mc=True)
print(json.dumps(res))
''' % ROOT


def run(names):
    out = {}
    libs = {"default": os.path.join(ROOT, "bluerov2_dynamics_b200", "libbrov.so")}
    libs.update({n: os.path.join(VDIR, f"libbrov_{n}.so") for n in VARIANTS})
    for name, lib in libs.items():
        if not os.path.exists(lib) or (names and name not in names):
            continue
        print(name, flush=True)
        r = subprocess.run([sys.executable, "-c", WORKER], env=dict(os.environ, BROV_LIB=lib), capture_output=True, text=True,
                           timeout=600)
        sys.stdout.write(r.stderr[-4000:])
        try:
            out[name] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(name, "FAILED", r.stdout[-500:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_variants.json"), "w"), indent=1)


if __name__ == "__main__":   # BROV_CASES=<prefix> restricts the timed cases
    {"build": build, "run": run}[sys.argv[1]](sys.argv[2:])
