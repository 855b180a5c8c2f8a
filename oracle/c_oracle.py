"""ctypes loader of oracle/libbrov_oracle.so — the plain-C restatement of the reference hot path (brov_oracle.c).

TEST INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.  The C file is a second, independent checker next to oracle/fossen_np.py: scalar float64, one
vehicle at a time, in the reference's own operation structure (sequential ThrusterLag steps, 6x6 Coriolis matrix
times nu, libm pow/sin/cos), OpenMP over vehicles / windows only.  It is pinned against the outputs of the unmodified
reference in tests/test_c_oracle_golden.py.

Build: `make -C oracle` (also done by __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import fossen_np as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libbrov_oracle.so")
_PATH_FMA = os.path.join(_HERE, "libbrov_oracle_fma.so")   # same code, FMA contraction allowed (conditioning studies)
_lib = None
_libs = {}

MODEL_ID = {"thruster8": 0, "wrench12": 1, "quat13": 2}
_PH_NAMES_ADDED = ["Xu_dot", "Yv_dot", "Zw_dot", "Kp_dot", "Mq_dot", "Nr_dot"]
_PH_NAMES_LIN = ["Xu", "Yv", "Zw", "Kp", "Mq", "Nr"]
_PH_NAMES_QUAD = ["Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs"]


def available() -> bool:
    return os.path.exists(_PATH)


def lib(fma: bool = False):
    global _lib
    if fma:
        if "fma" not in _libs:
            if not os.path.exists(_PATH_FMA):
                raise RuntimeError(f"{_PATH_FMA} not built: run `make -C oracle`")
            _libs["fma"] = _bind(C.CDLL(_PATH_FMA))
        return _libs["fma"]
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_PATH} not built: run `make -C oracle`")
        _lib = _bind(C.CDLL(_PATH))
    return _lib


def _bind(L):
    if True:
        dp, ll, i, d = C.c_void_p, C.c_longlong, C.c_int, C.c_double
        L.brov_oracle_rollout.argtypes = [i, i, ll, ll, d, dp, i, dp, dp, dp, dp, dp, dp, i, dp, dp, ll, dp]
        L.brov_oracle_rollout.restype = i
        L.brov_oracle_multistep_se.argtypes = [i, i, ll, ll, d, dp, dp, dp, dp, dp, dp, dp]
        L.brov_oracle_multistep_se.restype = d
        L.brov_oracle_threads.restype = i
    return L


def threads() -> int:
    return int(lib().brov_oracle_threads())


def pack_phys(p: dict, n: int | None = None) -> np.ndarray:
    """Physical parameter dict of oracle/fossen_np.py -> double[37] (layout BROV_PH_* of include/brov.h), or
    [n][37] when any entry is an array of length n (Monte-Carlo table)."""
    per = n is not None
    out = np.zeros((n if per else 1, 37))

    def put(col, v):
        out[:, col] = np.asarray(v, float)

    put(0, p["m"]); put(1, p["W"]); put(2, p["B"])
    put(3, p["xb"]); put(4, p["yb"]); put(5, p["zb"])
    put(6, p["Ix"]); put(7, p["Iy"]); put(8, p["Iz"])
    for j, k in enumerate(_PH_NAMES_ADDED):
        put(9 + j, p[k])
    for j, k in enumerate(_PH_NAMES_LIN):
        put(15 + j, p[k])
    for j, k in enumerate(_PH_NAMES_QUAD):
        put(21 + j, p[k])
    out[:, 27:33] = np.asarray(p["Minv"], float)
    out[:, 33:36] = np.asarray(p["current"], float)
    return np.ascontiguousarray(out if per else out[0])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _model_consts(kind, dt):
    if kind != "thruster8":
        return None, None, None, None
    Ad, Bd = O.lag_zoh(dt)
    r, e, _ = O.thruster_geometry()
    return (np.ascontiguousarray(Ad), np.ascontiguousarray(Bd), np.ascontiguousarray(r), np.ascontiguousarray(e))


def rollout(kind: str, integ: str, dt: float, x0, U, params: dict | None = None, lag0=None, stride: int = 0,
            fma: bool = False, min_abs_cos=None):
    """Same contract as fossen_np.rollout: x0 [N,nx]; U [T,N,nu] or [T,nu]; returns (snaps [S,N,nx], xT, lagT).
    fma=True runs the FMA-contracted build of the same code (a second valid rounding, for conditioning studies).
    min_abs_cos: float64 [N] array updated in place with the running minimum of |cos theta| over the states the steps
    start from (initialise it to 1)."""
    L = lib(fma)
    x = np.array(x0, float, ndmin=2, order="C")
    N, nx = x.shape
    U = np.ascontiguousarray(U, float)
    T = U.shape[0]
    shared = int(U.ndim == 2)
    p = O.default_params() if params is None else params
    per = np.ndim(p["Minv"]) == 2 or any(np.ndim(v) == 1 for k, v in p.items() if k not in ("Minv", "current"))
    phys = pack_phys(p, N if per else None)
    Ad, Bd, r, e = _model_consts(kind, dt)
    lag = None
    if kind == "thruster8":
        lag = np.zeros((N, 24)) if lag0 is None else np.array(lag0, float).reshape(N, 24).copy()
    S = T // stride if stride else 0
    traj = np.zeros((S, N, nx)) if S else None
    if min_abs_cos is not None:
        assert min_abs_cos.dtype == np.float64 and min_abs_cos.shape == (N,) and min_abs_cos.flags.c_contiguous
    rc = L.brov_oracle_rollout(MODEL_ID[kind], int(integ == "euler"), N, T, float(dt), _ptr(phys), int(per), _ptr(Ad),
                               _ptr(Bd), _ptr(r), _ptr(e), _ptr(x), _ptr(U), shared, _ptr(lag), _ptr(traj),
                               int(stride), _ptr(min_abs_cos))
    assert rc == 0
    return (traj if S else np.zeros((0, N, nx))), x, (lag.reshape(N, 8, 3) if lag is not None else None)


def multistep_se(kind: str, integ: str, dt: float, X, U, H: int, params: dict | None = None):
    """Sum of squared endpoint errors over windows k = 0..rows-H-1, every window starting from zero lag ("reset")
    -> (se, n_windows, rmse)."""
    L = lib()
    X = np.ascontiguousarray(X, float)
    U = np.ascontiguousarray(U, float)
    rows, nx = X.shape
    ns = rows - H
    if ns <= 0:
        return 0.0, 0, float("nan")
    phys = pack_phys(O.default_params() if params is None else params)
    Ad, Bd, r, e = _model_consts(kind, dt)
    se = L.brov_oracle_multistep_se(MODEL_ID[kind], int(integ == "euler"), rows, int(H), float(dt), _ptr(phys),
                                    _ptr(Ad), _ptr(Bd), _ptr(r), _ptr(e), _ptr(X), _ptr(U))
    return float(se), ns, float(np.sqrt(se / (ns * nx)))
