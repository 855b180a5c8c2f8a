"""Loads the byte-compiled UNMODIFIED reference from oracle/_ref/ (built by oracle/build_ref.py) — TEST / BASELINE
INFRASTRUCTURE, never imported by the product package.

    R = ref_loader.load()
    rov = R.BlueROV2(dt=0.02)                                   # fossen/BlueROV2.py:79
    traj = R.simulate_physics(x0, U_seq, 0.02, rov)             # training/train_tank_brov2_rk4.py:375-396
    rmse = R.multistep_rmse_endpoint_physics(X, U, H, 0.02)     # training/train_tank_brov2_rk4.py:399-417

matplotlib (absent from the image, used only by the reference's plotting helpers) is stubbed.  The reference's
top-level package names `fossen` / `Koopman` are imported from oracle/_ref; a process that called
bluerov2_dynamics_b200.install_as_fossen() must not call this."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_cache = None


def available() -> bool:
    return os.path.exists(os.path.join(REF, "fossen", "BlueROV2.refc")) and \
        os.path.exists(os.path.join(REF, "training", "train_tank_brov2_rk4.refc"))


def _load_pyc(name: str, rel: str):
    path = os.path.join(REF, rel)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


def load():
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    mine = sys.modules.get("fossen")
    if mine is not None and not getattr(mine, "__file__", "").startswith(REF):
        raise RuntimeError("a different top-level `fossen` is already imported in this process")
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.lines",
              "matplotlib.cm", "matplotlib.colors"):
        sys.modules.setdefault(m, MagicMock())
    for pkg in ("fossen", "Koopman"):
        p = types.ModuleType(pkg)
        p.__path__ = [os.path.join(REF, pkg)]
        p.__file__ = os.path.join(REF, pkg, "__init__.refc")
        sys.modules[pkg] = p
    _load_pyc("fossen.parameters", "fossen/parameters.refc")
    b = _load_pyc("fossen.BlueROV2", "fossen/BlueROV2.refc")
    bt = _load_pyc("fossen.BlueROV2_thrust", "fossen/BlueROV2_thrust.refc")
    bw = _load_pyc("fossen.BlueROV2_wrench", "fossen/BlueROV2_wrench.refc")
    _load_pyc("fossen.bluerov_torch", "fossen/bluerov_torch.refc")
    _load_pyc("Koopman.koopmanEDMDc", "Koopman/koopmanEDMDc.refc")
    rk4 = _load_pyc("ref_train_tank_brov2_rk4", "training/train_tank_brov2_rk4.refc")
    cmp_ = _load_pyc("ref_train_tank_brov2_full_comparison", "training/train_tank_brov2_full_comparison.refc")
    ns = types.SimpleNamespace(BlueROV2=b.BlueROV2, ThrusterLag=b.ThrusterLag, BlueROV2_thrust=bt.BlueROV2,
                               BlueROV2_wrench=bw.BlueROV2, simulate_physics=rk4.simulate_physics,
                               multistep_rmse_endpoint_physics=rk4.multistep_rmse_endpoint_physics,
                               simulate_physics_euler=cmp_.simulate_physics,
                               multistep_rmse_endpoint_physics_euler=cmp_.multistep_rmse_endpoint_physics)
    _cache = ns
    return ns
