"""Builds tuning variants of libbrov.so (compile-time knobs of csrc/brov_kernels.cuh) and, on a GPU, times the
fp64 / fp32 rollout kernel of each.  Usage:  python profiles/tune_variants.py build | run"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "f32_r96": ["BROV_F32_MAXREG=96"],
    "f32_r160": ["BROV_F32_MAXREG=160"],
    "f32_r255": ["BROV_F32_MAXREG=255"],
}
VDIR = os.path.join(ROOT, "bluerov2_dynamics_b200", "variants")

def build():
    from bluerov2_dynamics_b200.build import build_lib
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"libbrov_{name}.so")
        t = time.time()
        build_lib(force=True, defines=defs, out=out, tag="_" + name)
        log = open(os.path.join(ROOT, "bluerov2_dynamics_b200", "build", f"brov_kernels_f64_{name}.o.log")).read()
        i = log.find("rollout_kernelIdLi0ELi0ELb0ELb0ELb1E")
        print(name, f"{time.time()-t:.0f}s", log[i:i + 400].split("\n")[1:3])

WORKER = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import bluerov2_dynamics_b200 as B
res = {}
for dtype, n in (("f64", 65536), ("f32", 1 << 20)):
    e = B.Engine("thruster8", dtype)
    g = torch.Generator(device="cuda").manual_seed(1)
    U = [(torch.rand((100, n, 8), device="cuda", dtype=e.tdtype, generator=g) * 0.8 - 0.4) for _ in range(2)]
    x = torch.zeros((n, 12), device="cuda", dtype=e.tdtype); lag = torch.zeros((n, 18), device="cuda", dtype=e.tdtype)
    stride = 10 if dtype == "f32" else 0
    traj = torch.empty((10, n, 12), device="cuda", dtype=e.tdtype) if stride else None
    def one(k):
        e.rollout(x, U[k %% 2], dt=0.02, lag0=lag, xT_out=x, lag_out=lag, lag_repr="projected", stride=stride, traj_out=traj, step0=k * 100)
    for k in range(3): one(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(10): one(k)
    b.record(); torch.cuda.synchronize()
    res[dtype] = a.elapsed_time(b) / 10
print(json.dumps(res))
''' % ROOT

def run():
    out = {}
    libs = {"default": os.path.join(ROOT, "bluerov2_dynamics_b200", "libbrov.so")}
    libs.update({n: os.path.join(VDIR, f"libbrov_{n}.so") for n in VARIANTS})
    for name, lib in libs.items():
        if not os.path.exists(lib):
            continue
        r = subprocess.run([sys.executable, "-c", WORKER], env=dict(os.environ, BROV_LIB=lib), capture_output=True, text=True)
        try:
            ms = json.loads(r.stdout.strip().splitlines()[-1])
            out[name] = ms
            print(f"{name:28s} f64 {ms['f64']:.3f} ms ({65536*100/ms['f64']/1e6:.2f}e9/s)   f32 {ms['f32']:.3f} ms ({(1<<20)*100/ms['f32']/1e6:.2f}e9/s)", flush=True)
        except Exception:
            print(name, "FAILED", r.stderr[-500:])
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_variants.json"), "w"), indent=1)

if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
