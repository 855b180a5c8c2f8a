"""`fossen.BlueROV2_thrust` mirror — despite the file name this is the WRENCH-input 12-state Euler-angle model
(reference: fossen/BlueROV2_thrust.py; SURVEY trap T1).  State [x y z phi theta psi u v w p q r], input
tau = [Fx Fy Fz Mx My Mz] in the body frame."""
import numpy as np

from ._base import FossenModelBase


def rotation_matrix(phi, theta, psi):
    """R_{b->n} = Rz(psi) Ry(theta) Rx(phi) (fossen/BlueROV2_thrust.py:20-39); host-side helper."""
    cx, sx, cy, sy, cz, sz = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta), np.cos(psi), np.sin(psi)
    Rx = np.array([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
    Ry = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    Rz = np.array([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
    return Rz @ Ry @ Rx


def euler_kinematics_matrix(phi, theta, eps=1e-7):
    """Body rates -> Euler-angle rates, with the reference's cos(theta) clamp (fossen/BlueROV2_thrust.py:42-61)."""
    s, c, ct = np.sin(phi), np.cos(phi), np.cos(theta)
    if abs(ct) < eps:
        ct = eps * np.sign(ct)
    t = np.sin(theta) / ct
    return np.array([[1.0, s * t, c * t], [0.0, c, -s], [0.0, s / ct, c / ct]], dtype=float)


class BlueROV2(FossenModelBase):
    """BlueROV2 heavy, direct wrench input.  `dynamics(x, tau_body, dt)` -> xdot (12,)."""
    _MODEL = "wrench12"

    def __init__(self, rho=1000.0, current_speed=None):
        self._init_constants(rho, current_speed)
        self.current_speed = np.asarray(self.current_speed, dtype=float).reshape(3,)

    def dynamics(self, x, tau_body, dt=0.02):
        x = np.asarray(x, dtype=float).reshape(12,)        # ValueError on bad shapes, as the reference
        tau_body = np.asarray(tau_body, dtype=float).reshape(6,)
        return self._dynamics_one(x, tau_body, 12, 6, dt)
