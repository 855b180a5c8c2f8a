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

    # --- the reference's private term helpers, for callers that poke at them (fossen/BlueROV2.py:280-355 and the twins
    # in BlueROV2_thrust.py:150-230, BlueROV2_wrench.py:228-319).  Host numpy on the live attributes: they are not on
    # the accelerated path (the kernels evaluate C(nu) nu, D(nu_r) nu_r and g(eta) in closed form), they keep the
    # attribute surface whole.
    @staticmethod
    def _skew(a):
        return np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])

    def _coriolis(self, nu):
        """C(nu) = C_RB + C_A for a diagonal mass matrix and CG at the origin (Fossen 2011, eqs. 3.60 / 6.43):
        [[0, -S(M11 v1)], [-S(M11 v1), -S(M22 v2)]] with M = M_RB + M_A."""
        nu = np.asarray(nu, float).reshape(6)
        m11 = np.array([self.m - self.Xu_dot, self.m - self.Yv_dot, self.m - self.Zw_dot])
        m22 = np.array([self.Ix - self.Kp_dot, self.Iy - self.Mq_dot, self.Iz - self.Nr_dot])
        C = np.zeros((6, 6))
        C[0:3, 3:6] = C[3:6, 0:3] = -self._skew(m11 * nu[0:3])
        C[3:6, 3:6] = -self._skew(m22 * nu[3:6])
        return C

    def _damping(self, nu_r):
        """D(nu_r) = diag(-L_i - Q_i |nu_r,i|): linear + quadratic damping on the relative velocity."""
        nu_r = np.asarray(nu_r, float).reshape(6)
        lin = np.array([getattr(self, n) for n in _LIN])
        quad = np.array([getattr(self, n) for n in _QUAD])
        return np.diag(-lin - quad * np.abs(nu_r))

    def _restoring(self, phi, theta, psi=0.0):
        """g(eta) for CG at the origin and centre of buoyancy (xb, yb, zb)."""
        sph, cph, sth, cth = np.sin(phi), np.cos(phi), np.sin(theta), np.cos(theta)
        wb, bx, by, bz = self.W - self.B, self.xb * self.B, self.yb * self.B, self.zb * self.B
        return np.array([wb * sth, -wb * cth * sph, -wb * cth * cph,
                         by * cth * cph - bz * cth * sph, -bz * sth - bx * cth * cph, bx * cth * sph + by * sth])

    def _dynamics_one(self, x, u, nx, nu, dt, lag=None):
        """One dynamics() call through the host-buffer entry point (brov_rhs_host); `lag` is a numpy [1,24] array that
        is advanced in place."""
        return self.engine("f64").rhs_host(x, u, lag=lag, dt=dt)[0]
