"""PINc comparison model on the B200 engine: inference of the reference's residual network
x_{k+1} = x_k + f_theta([x_k, u_k, dt]) and its rollout / endpoint-RMSE evaluators (reference:
training/train_tank_brov2_rk4.py:553-840).  Training the network is outside the accelerated path; a trained
`PINcNet` (its `state_dict()`, e.g. the reference's models/pinc_best.pt) is loaded into a `PincModel`.

Same function names and argument order as the reference script: `thrusters_to_body_wrenches`, `dataset12_to_9`,
`batch12_to_9`, `state9_to_12`, `batch9_to_12`, `simulate_pinc`, `multistep_rmse_endpoint_pinc`.  Where the reference
takes a torch module and a torch device, these take a `PincModel` (anything with a `state_dict()` is converted) and
ignore the device argument: the compute runs in libbrov.so (brov_pinc_* of include/brov.h), one thread per window.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .engine import Engine, default_allocation, lag_discretize


# ----------------------------------------------------------------------------------------------------------------------
# 12-state <-> 9-state conversions (host helpers, :565-596)
# ----------------------------------------------------------------------------------------------------------------------
def dataset12_to_9(x12: np.ndarray) -> np.ndarray:
    """[x,y,z,phi,theta,psi,u,v,w,p,q,r] -> [x,y,z,cos psi,sin psi,u,v,w,r]."""
    x12 = np.asarray(x12, dtype=float)
    return np.array([x12[0], x12[1], x12[2], math.cos(x12[5]), math.sin(x12[5]), x12[6], x12[7], x12[8], x12[11]])


def batch12_to_9(X12: np.ndarray) -> np.ndarray:
    X12 = np.asarray(X12, dtype=float)
    return np.stack([X12[:, 0], X12[:, 1], X12[:, 2], np.cos(X12[:, 5]), np.sin(X12[:, 5]), X12[:, 6], X12[:, 7],
                     X12[:, 8], X12[:, 11]], axis=1)


def state9_to_12(x9: np.ndarray) -> np.ndarray:
    """9-state -> 12-state row for plotting / metrics; roll, pitch and their rates are not modelled (zeros)."""
    x9 = np.asarray(x9, dtype=float)
    out = np.zeros(12)
    out[0:3] = x9[0:3]
    out[5] = math.atan2(x9[4], x9[3])
    out[6:9] = x9[5:8]
    out[11] = x9[8]
    return out


def batch9_to_12(X9: np.ndarray) -> np.ndarray:
    X9 = np.asarray(X9, dtype=float)
    out = np.zeros((X9.shape[0], 12))
    out[:, 0:3] = X9[:, 0:3]
    out[:, 5] = np.arctan2(X9[:, 4], X9[:, 3])
    out[:, 6:9] = X9[:, 5:8]
    out[:, 11] = X9[:, 8]
    return out


def thrusters_to_body_wrenches(U8_row: np.ndarray, dt: float, old_model_obj) -> np.ndarray:
    """[X, Y, Z, Mz] of the 6-DOF thruster map; advances `old_model_obj`'s lag state by one step like the reference."""
    tau6 = old_model_obj.compute_thruster_forces(U8_row, dt)
    return np.array([tau6[0], tau6[1], tau6[2], tau6[5]], dtype=float)


def make_pinc_dataset(X12: np.ndarray, U8: np.ndarray, dt: float, old6):
    """(x9_k, u4_k, dt) -> x9_{k+1} training pairs of the PINc network: returns (z_in [N-1,14], y [N-1,9], U4 [N,4])
    like the reference (training/train_tank_brov2_rk4.py:676-696).  The thruster map runs along the whole series with
    `old6`'s lag state carried from row to row — one launch (Engine.thruster_wrench_series) instead of N Python calls —
    and `old6` is left with the lag state after the last row."""
    X12 = np.asarray(X12, dtype=float)
    U8 = np.asarray(U8, dtype=float)
    eng = old6.engine("f64")
    lag0 = _rov_lag(old6)
    tau, lag_end = eng.thruster_wrench_series(U8, lag0=lag0, dt=dt)
    if len(U8):
        old6._store_lag(lag_end.reshape(1, 24), dt)
    U4 = tau[:, [0, 1, 2, 5]].cpu().numpy()
    X9 = batch12_to_9(X12)
    z_in = np.hstack([X9[:-1], U4[:-1], np.full((len(X9) - 1, 1), dt, dtype=float)])
    return z_in, X9[1:], U4


# ----------------------------------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------------------------------
class PincModel:
    """Device-resident PINcNet weights (4 hidden layers of 64 units, as the reference's PINc_HIDDEN)."""

    def __init__(self, state_dict, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("bluerov2_dynamics_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        sd = state_dict.state_dict() if hasattr(state_dict, "state_dict") else state_dict

        def g(k):
            v = sd[k]
            v = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
            return np.ascontiguousarray(v, dtype=np.float32)

        lin, act, ln = (0, 3, 6, 9, 12), (1, 4, 7, 10), (2, 5, 8, 11)
        self._keep = []   # host arrays must outlive brov_pinc_create
        w = L.PincWeights()
        w.struct_size = C.sizeof(L.PincWeights)
        w.n_hidden_layers = 4
        shapes = [(64, 14), (64, 64), (64, 64), (64, 64), (9, 64)]
        for i, k in enumerate(lin):
            W, b = g(f"net.{k}.weight"), g(f"net.{k}.bias")
            if W.shape != shapes[i] or b.shape != (shapes[i][0],):
                raise ValueError(f"net.{k}: expected weight {shapes[i]}, got {W.shape} — the engine is built for the "
                                 "reference's 14-64-64-64-64-9 network")
            self._keep += [W, b]
            w.W[i], w.b[i] = W.ctypes.data, b.ctypes.data
        w.hidden = 64
        for i, (ka, kl) in enumerate(zip(act, ln)):
            w.beta[i] = float(g(f"net.{ka}.beta").reshape(-1)[0])
            lw, lb = g(f"net.{kl}.weight"), g(f"net.{kl}.bias")
            self._keep += [lw, lb]
            w.ln_w[i], w.ln_b[i] = lw.ctypes.data, lb.ctypes.data
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        L.check(L.lib.brov_pinc_create(self.device_index, C.byref(w), C.byref(h)))
        self._h = h
        self._dt = None
        self._thr = None

    @classmethod
    def from_checkpoint(cls, path, device: Optional[int] = None) -> "PincModel":
        return cls(torch.load(path, map_location="cpu", weights_only=True), device)   # a plain state_dict of tensors

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and L is not None and getattr(L, "lib", None) is not None:   # module globals may be gone at shutdown
            L.lib.brov_pinc_destroy(h)
            self._h = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _prepare(self, dt: float):
        if self._dt != dt:
            Ad, Bd = lag_discretize(dt)
            alloc = default_allocation()[0]
            L.check(L.lib.brov_pinc_set_thruster_map(self._h, float(dt), L.dptr(np.ascontiguousarray(Ad)),
                                                     L.dptr(np.ascontiguousarray(Bd)), L.dptr(np.ascontiguousarray(alloc))))
            self._dt = dt

    def _f64(self, a, cols):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        t = t.to(self.device, torch.float64).contiguous()
        if t.shape[-1] != cols:
            raise ValueError(f"expected {cols} columns, got shape {tuple(t.shape)}")
        return t

    def thruster_engine(self) -> Engine:
        if self._thr is None:
            self._thr = Engine("thruster8", "f64", device=self.device_index)
        return self._thr

    # ------------------------------------------------------------------ PINcNet.forward
    def forward(self, z) -> torch.Tensor:
        """z [B,14] = [x9, u4, dt] -> x9_next [B,9] (float32 CUDA tensor)."""
        t = z if torch.is_tensor(z) else torch.from_numpy(np.asarray(z, dtype=np.float32))
        t = t.to(self.device, torch.float32).contiguous()
        if t.ndim != 2 or t.shape[1] != 14:
            raise ValueError("z must have shape [B, 14]")
        out = torch.empty((t.shape[0], 9), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(L.lib.brov_pinc_forward(self._h, t.data_ptr(), out.data_ptr(), t.shape[0], self._stream()))
        return out

    __call__ = forward

    def eval(self):
        return self

    # ------------------------------------------------------------------ batched rollout
    def rollout(self, x0_12, U8, dt: float, lag0=None, stride: int = 1):
        """simulate_pinc for N windows: x0_12 [N,12]; U8 [T,N,8] or [T,8] (shared).
        -> (traj12 [T//stride, N, 12] float64 or None, x9_T [N,9] float32, lag [N,12] projected float64)."""
        self._prepare(dt)
        x0 = self._f64(np.atleast_2d(x0_12) if not torch.is_tensor(x0_12) else x0_12, 12)
        U = self._f64(U8, 8)
        n, T = x0.shape[0], U.shape[0]
        shared = U.ndim == 2
        if not shared and U.shape[1] != n:
            raise ValueError("U8 must be [T, N, 8] or [T, 8]")
        d = L.PincRolloutDesc()
        d.struct_size = C.sizeof(L.PincRolloutDesc)
        d.n, d.steps = n, T
        d.x0_dev, d.u_dev = x0.data_ptr(), U.data_ptr()
        d.u_stride_t, d.u_stride_n = (8, 0) if shared else (n * 8, 8)
        lag_in = None
        if lag0 is not None:
            lag_in = self._f64(torch.as_tensor(np.asarray(lag0, dtype=np.float64)).reshape(n, 24), 24)
        d.lag_in_dev = lag_in.data_ptr() if lag_in is not None else None
        lag_out = torch.empty((n, 12), device=self.device, dtype=torch.float64)
        d.lag_out_dev = lag_out.data_ptr()
        traj = torch.empty((T // stride, n, 12), device=self.device, dtype=torch.float64) if stride else None
        d.traj_dev = traj.data_ptr() if traj is not None and traj.numel() else None
        d.stride = max(int(stride), 1)
        x9 = torch.empty((n, 9), device=self.device, dtype=torch.float32)
        d.x9T_dev = x9.data_ptr()
        with torch.cuda.device(self.device):
            L.check(L.lib.brov_pinc_rollout(self._h, C.byref(d), self._stream()))
        return traj, x9, lag_out

    # ------------------------------------------------------------------ evaluator
    def multistep_se(self, X12, U8, horizons, dt: float, lag_mode: str = "reset", carry_lag0=None, n_windows=None,
                     window0: int = 0, row0: int = 0):
        """Sum of squared endpoint errors per horizon -> (se [MAX_H] float64 CUDA tensor, window counts)."""
        self._prepare(dt)
        X, U = self._f64(X12, 12), self._f64(U8, 8)
        rows = X.shape[0]
        if U.shape[0] != rows:
            raise ValueError("X and U must have the same number of rows")
        hs = [int(h) for h in horizons]
        if not 1 <= len(hs) <= L.MAX_H or any(h < 1 for h in hs) or sorted(set(hs)) != hs:
            raise ValueError(f"horizons must be 1..{L.MAX_H} strictly ascending positive integers")
        if lag_mode not in ("reset", "carry"):
            raise ValueError("lag_mode must be 'reset' or 'carry'")
        off = int(window0) - int(row0)
        nwin = max(rows - off - hs[0], 0) if n_windows is None else int(n_windows)
        se = torch.zeros(L.MAX_H, dtype=torch.float64, device=self.device)
        d = L.PincSeDesc()
        d.struct_size = C.sizeof(L.PincSeDesc)
        d.n_horizons = len(hs)
        for i, h in enumerate(hs):
            d.horizons[i] = h
        d.rows, d.n_windows = rows, nwin
        d.X_dev, d.U_dev, d.se_out_dev = X.data_ptr(), U.data_ptr(), se.data_ptr()
        d.carry_steps = self.thruster_engine().carry_steps(dt, "euler") if lag_mode == "carry" else 0
        d.window0, d.row0 = (int(window0), int(row0)) if lag_mode == "carry" else (0, 0)
        l0 = None
        if lag_mode == "carry" and carry_lag0 is not None:
            l0 = self._f64(np.asarray(carry_lag0, dtype=np.float64).reshape(1, 24), 24)
        d.carry_lag0_dev = l0.data_ptr() if l0 is not None else None
        with torch.cuda.device(self.device):
            L.check(L.lib.brov_pinc_multistep_se(self._h, C.byref(d), self._stream()))
        counts = [max(min(rows - off - h, nwin), 0) for h in hs]
        return se, counts


def _as_model(model) -> PincModel:
    return model if isinstance(model, PincModel) else PincModel(model)


def _rov_lag(rov):
    if rov is None:
        return None
    return np.concatenate([np.asarray(l._x, float).reshape(3) for l in rov.thruster_lags])


def _advance_rov_lag(model: PincModel, rov, U_hist: np.ndarray, dt: float, lag0) -> None:
    """Leave `rov`'s lag state where the reference's loop over compute_thruster_forces would: one lag step per row of
    U_hist (an Euler rollout of the thruster engine advances the lag exactly once per step)."""
    if rov is None or len(U_hist) == 0:
        return
    eng = model.thruster_engine()
    res = eng.rollout(np.zeros((1, 12)), np.asarray(U_hist, float), dt=dt, integrator="euler",
                      lag0=None if lag0 is None else np.asarray(lag0, float).reshape(1, 24), u_layout="shared")
    rov._store_lag(res.lag, dt)


def simulate_pinc(x0_12: np.ndarray, U_seq_8: np.ndarray, dt: float, model, old_model_for_map=None, device=None):
    """Rollout of the PINc model; returns the trajectory in the 12-state projection, [len(U_seq_8)+1, 12], row 0 = x0.
    The thruster map starts from `old_model_for_map`'s lag state and leaves it advanced (None: zero lag)."""
    model = _as_model(model)
    x0_12 = np.asarray(x0_12, dtype=float)
    U = np.asarray(U_seq_8, dtype=float).reshape(-1, 8)
    H = len(U)
    traj12 = np.zeros((H + 1, 12))
    traj12[0] = x0_12
    if H == 0:
        return traj12
    lag0 = _rov_lag(old_model_for_map)
    traj, _, _ = model.rollout(x0_12.reshape(1, 12), U, dt, lag0=None if lag0 is None else lag0.reshape(1, 24))
    traj12[1:] = traj[:, 0, :].cpu().numpy()
    _advance_rov_lag(model, old_model_for_map, U, dt, lag0)
    return traj12


def multistep_rmse_endpoint_pinc(X_test: np.ndarray, U_test: np.ndarray, H, dt: float, model, old_model_for_map=None,
                                 device=None, lag_mode: str = "carry"):
    """Endpoint RMSE of the PINc model in the 12-state projection over all sliding windows (NaN if T <= H).
    lag_mode="carry" (default) is the reference's literal behaviour: the ONE thruster-map object scores all windows in
    order, so its lag state runs on from window to window (and from `old_model_for_map`'s state at entry, which is
    left advanced on return).  lag_mode="reset" starts every window from zero lag and scores a list of horizons in
    one pass.  `H` may be a list."""
    model = _as_model(model)
    X = np.asarray(X_test, dtype=float)
    U = np.asarray(U_test, dtype=float)
    single = np.isscalar(H)
    hs = [int(H)] if single else [int(h) for h in H]
    T = len(X)
    out = {}
    if lag_mode == "reset":
        order = sorted(set(hs))
        for i in range(0, len(order), L.MAX_H):
            part = order[i:i + L.MAX_H]
            se, cnt = model.multistep_se(X, U, part, dt, "reset")
            se = se.cpu().numpy()
            for j, h in enumerate(part):
                out[h] = float(np.sqrt(se[j] / (cnt[j] * 12))) if cnt[j] > 0 else float("nan")
    else:
        for h in hs:   # in call order: the lag state runs on from one horizon's pass to the next, as in the reference
            lag0 = _rov_lag(old_model_for_map)
            ns = T - h
            if ns <= 0:
                out[h] = float("nan")
                continue
            se, cnt = model.multistep_se(X, U, [h], dt, "carry", carry_lag0=lag0)
            out[h] = float(np.sqrt(float(se[0].item()) / (cnt[0] * 12)))
            if old_model_for_map is not None:
                m = model.thruster_engine().carry_steps(dt, "euler")
                total = ns * h
                keep = min(total, m)
                s = np.arange(total - keep, total)
                w = s // h
                _advance_rov_lag(model, old_model_for_map, U[w + (s - w * h)], dt, lag0 if total <= m else None)
    return out[hs[0]] if single else [out[h] for h in hs]
