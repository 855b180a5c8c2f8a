"""`fossen.BlueROV2_wrench` mirror — wrench-input 13-state quaternion model and its helper functions
(reference: fossen/BlueROV2_wrench.py).  State [x y z qw qx qy qz u v w p q r]."""
import numpy as np

from ._base import FossenModelBase


def quat_normalize(q, eps=1e-12):
    """Unit quaternion; identity when the norm is below eps (fossen/BlueROV2_wrench.py:27-36)."""
    q = np.asarray(q, dtype=float).reshape(4,)
    n = float(np.sqrt(q @ q))
    return np.array([1.0, 0.0, 0.0, 0.0]) if n < eps else q / n


def quat_to_rotation_matrix(q):
    """Scalar-first quaternion -> R_{b->n} (fossen/BlueROV2_wrench.py:39-53)."""
    w, x, y, z = quat_normalize(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=float)


def quat_multiply(q1, q2):
    """Hamilton product (fossen/BlueROV2_wrench.py:56-68)."""
    a = np.asarray(q1, dtype=float).reshape(4,)
    b = np.asarray(q2, dtype=float).reshape(4,)
    s = a[0] * b[0] - a[1:] @ b[1:]
    v = a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])
    return np.array([s, v[0], v[1], v[2]], dtype=float)


def quat_derivative(q, omega_body):
    """q_dot = 0.5 q (x) [0, omega] (fossen/BlueROV2_wrench.py:71-80)."""
    w = np.asarray(omega_body, dtype=float).reshape(3,)
    return 0.5 * quat_multiply(q, np.array([0.0, w[0], w[1], w[2]]))


def euler_to_quat(phi, theta, psi):
    """Z-Y-X Euler angles -> quaternion (fossen/BlueROV2_wrench.py:86-106)."""
    hx, hy, hz = 0.5 * float(phi), 0.5 * float(theta), 0.5 * float(psi)
    qx = np.array([np.cos(hx), np.sin(hx), 0.0, 0.0])
    qy = np.array([np.cos(hy), 0.0, np.sin(hy), 0.0])
    qz = np.array([np.cos(hz), 0.0, 0.0, np.sin(hz)])
    return quat_normalize(quat_multiply(qz, quat_multiply(qy, qx)))


def quat_to_euler(q):
    """Quaternion -> (phi, theta, psi) (fossen/BlueROV2_wrench.py:109-132)."""
    w, x, y, z = quat_normalize(q)
    phi = np.arctan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    theta = np.arcsin(np.clip(2.0 * (w * y - z * x), -1.0, 1.0))
    psi = np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return phi, theta, psi


def quat_to_yaw(q):
    """fossen/BlueROV2_wrench.py:134-138."""
    return float(quat_to_euler(q)[2])


class BlueROV2(FossenModelBase):
    """BlueROV2 heavy, direct wrench input, quaternion attitude.  `dynamics(x, tau_body, dt)` -> xdot (13,)."""
    _MODEL = "quat13"

    def __init__(self, rho=1000.0, current_speed=None):
        self._init_constants(rho, current_speed)
        self.current_speed = np.asarray(self.current_speed, dtype=float).reshape(3,)

    def dynamics(self, x, tau_body, dt=0.02):
        x = np.asarray(x, dtype=float).reshape(13,)
        tau_body = np.asarray(tau_body, dtype=float).reshape(6,)
        return self._dynamics_one(x, tau_body, 13, 6, dt)
