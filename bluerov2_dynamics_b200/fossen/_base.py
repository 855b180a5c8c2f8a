"""Shared host-side state of the three 6-DOF model mirrors.

The mirrors keep the reference's attribute surface (`rov.m`, `rov.Xu_dot`, `rov.Minv`, `rov.current_speed`, ...) as
plain Python attributes.  Every call packs the *live* attribute values into the engine's physical vector, so editing
an attribute after construction behaves as in the reference: Coriolis, damping and restoring terms follow the edit,
while `Minv` stays whatever array the object holds (the reference computes it once in __init__, SURVEY trap T5).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from ..engine import Engine

_ADDED = ("Xu_dot", "Yv_dot", "Zw_dot", "Kp_dot", "Mq_dot", "Nr_dot")
_LIN = ("Xu", "Yv", "Zw", "Kp", "Mq", "Nr")
_QUAD = ("Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs")


class FossenModelBase:
    _MODEL = "wrench12"

    def _init_constants(self, rho, current_speed):
        from ..engine import default_physical
        p = default_physical(rho)
        self.rho = rho
        self.g = 9.82
        self.m = float(p[L.PH_M])
        self.volume = 0.0134
        self.W = float(p[L.PH_W])
        self.B = float(p[L.PH_B])
        self.xg = self.yg = self.zg = 0.0
        self.xb, self.yb, self.zb = (float(v) for v in p[L.PH_XB:L.PH_XB + 3])
        self.Ix, self.Iy, self.Iz = (float(v) for v in p[L.PH_I:L.PH_I + 3])
        self.MRB = np.diag([self.m, self.m, self.m, self.Ix, self.Iy, self.Iz]).astype(float)
        for i, name in enumerate(_ADDED):
            setattr(self, name, float(p[L.PH_ADDED + i]))
        self.MA = np.diag([-getattr(self, n) for n in _ADDED]).astype(float)
        self.M = self.MRB + self.MA
        self.Minv = np.diag(p[L.PH_MINV:L.PH_MINV + 6]).astype(float)
        for i, name in enumerate(_LIN):
            setattr(self, name, float(p[L.PH_LIN + i]))
        for i, name in enumerate(_QUAD):
            setattr(self, name, float(p[L.PH_QUAD + i]))
        if current_speed is None:
            current_speed = np.zeros(3, dtype=float)
        self.current_speed = current_speed
        self._engines = {}

    def physical_vector(self) -> np.ndarray:
        """Live attributes -> engine physical vector (include/brov.h BROV_PH_*)."""
        p = np.zeros(L.NPHYS)
        p[L.PH_M], p[L.PH_W], p[L.PH_B] = self.m, self.W, self.B
        p[L.PH_XB:L.PH_XB + 3] = (self.xb, self.yb, self.zb)
        p[L.PH_I:L.PH_I + 3] = (self.Ix, self.Iy, self.Iz)
        p[L.PH_ADDED:L.PH_ADDED + 6] = [getattr(self, n) for n in _ADDED]
        p[L.PH_LIN:L.PH_LIN + 6] = [getattr(self, n) for n in _LIN]
        p[L.PH_QUAD:L.PH_QUAD + 6] = [getattr(self, n) for n in _QUAD]
        Minv = np.asarray(self.Minv, float)
        if Minv.shape != (6, 6) or np.count_nonzero(Minv - np.diag(np.diagonal(Minv))):
            raise NotImplementedError("the engine supports the reference's diagonal mass matrix only")
        p[L.PH_MINV:L.PH_MINV + 6] = np.diagonal(Minv)
        p[L.PH_CURRENT:L.PH_CURRENT + 3] = np.asarray(self.current_speed, float).reshape(3)
        return p

    def engine(self, dtype: str = "f64") -> Engine:
        """The CUDA engine behind this object (created lazily), with the live attribute values pushed."""
        eng = self._engines.get(dtype)
        if eng is None:
            eng = self._engines[dtype] = Engine(self._MODEL, dtype)
            eng._pushed = None
        p = self.physical_vector()
        if eng._pushed is None or not np.array_equal(eng._pushed, p):
            eng.set_physical(p)
            eng._pushed = p
        return eng

    def _dynamics_one(self, x, u, nx, nu, dt, lag=None):
        """One dynamics() call through the host-buffer entry point (brov_rhs_host); `lag` is a numpy [1,24] array that
        is advanced in place."""
        return self.engine("f64").rhs_host(x, u, lag=lag, dt=dt)[0]
