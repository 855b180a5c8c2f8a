"""`fossen.BlueROV2` mirror — the 8-thruster model: T200 polynomial, 3rd-order ThrusterLag per thruster (hidden,
STATEFUL: every dynamics()/compute_thruster_forces() call advances it once, SURVEY trap T2), allocation through the
reference's thruster geometry (reference: fossen/BlueROV2.py)."""
import numpy as np
import torch

from ..engine import default_allocation, lag_discretize
from ._base import FossenModelBase
from .BlueROV2_thrust import euler_kinematics_matrix, rotation_matrix  # noqa: F401  (same helpers, re-exported)


class ThrusterLag:
    """Third-order unity-gain thrust lag K(s) = (6136 s + 108700)/(s^3 + 89 s^2 + 9258 s + 108700), discretised
    (zero-order hold) lazily for the dt it is stepped with (fossen/BlueROV2.py:464-510).  Inside a BlueROV2 object the
    eight `_x` vectors are the hidden state the CUDA kernels read and write back; `step()` on a free-standing instance
    is a host-side convenience."""
    _Ac = np.array([[-89.0, -72.33, -26.54], [128.0, 0.0, 0.0], [0.0, 32.0, 0.0]])
    _Bc = np.array([[8.0], [0.0], [0.0]])
    _Cc = np.array([[0.0, 5.992, 3.317]])
    _Dc = np.zeros((1, 1))

    def __init__(self):
        self._dt = None
        self._Ad = None
        self._Bd = None
        self._x = np.zeros(3)

    @staticmethod
    def _discretise(A, B, C, D, dt):
        if not (np.array_equal(A, ThrusterLag._Ac) and np.array_equal(np.reshape(B, (3, 1)), ThrusterLag._Bc)):
            raise NotImplementedError("only the reference lag model is built into the engine")
        Ad, Bd = lag_discretize(dt)
        return Ad, Bd.reshape(3, 1)

    def _prepare(self, dt):
        if self._dt != dt:
            self._Ad, self._Bd = self._discretise(self._Ac, self._Bc, self._Cc, self._Dc, dt)
            self._dt = dt

    def step(self, u, dt):
        self._prepare(dt)
        self._x = self._Ad @ self._x + self._Bd[:, 0] * u
        return float(self._Cc[0] @ self._x)


class Tether:
    """The lumped-mass tether of the reference (fossen/BlueROV2.py:517-663) is outside the accelerated path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("the tether model is not part of the B200 engine (default-off in the reference)")


class BlueROV2(FossenModelBase):
    """BlueROV2 heavy with 8 thrusters.  `dynamics(x, u_thrust, dt)` -> xdot (12,), u_thrust = normalised voltages in
    [-1, 1].  The constructor's `dt` is accepted and ignored, as in the reference (trap T4)."""
    _MODEL = "thruster8"

    def __init__(self, rho=1000.0, current_speed=np.array([0.0, 0.0, 0.0]), dt=0.01):
        self._init_constants(rho, current_speed)
        self.n_thrusters = 8
        self.thrusters_r = self._define_thruster_placements()
        self.thruster_lags = [ThrusterLag() for _ in range(self.n_thrusters)]
        self.use_tether = False
        self.tether = None
        self.tether_state = None
        self.anchor_pos = np.zeros(3)

    def _define_thruster_placements(self):
        _, r, d = default_allocation()
        return [{"r": r[i].copy(), "dir": d[i].copy()} for i in range(8)]

    def _old_thruster_force_from_input(self, V):
        V = float(V)
        z = V * V
        return V * (8.9 + z * (176.0 + z * (-404.1 + z * (389.9 - 140.3 * z))))

    def _thruster_force_from_input(self, V, i, dt):
        """Static T200 curve, then ONE step of thruster i's lag (stateful), fossen/BlueROV2.py:245-263."""
        return float(self.thruster_lags[i].step(self._old_thruster_force_from_input(V), dt))

    # --- hidden lag state <-> device ---------------------------------------------------------------------------
    def _lag_tensor(self, eng):
        flat = np.concatenate([np.asarray(l._x, float).reshape(3) for l in self.thruster_lags])
        return torch.from_numpy(flat.reshape(1, 24)).to(eng.device)

    def _store_lag(self, lag, dt):
        vals = lag.cpu().numpy().reshape(8, 3)
        for i, l in enumerate(self.thruster_lags):
            l._prepare(dt)
            l._x = vals[i].copy()

    def _lag_host(self):
        return np.concatenate([np.asarray(l._x, float).reshape(3) for l in self.thruster_lags]).reshape(1, 24)

    def _store_lag_host(self, lag, dt):
        for i, l in enumerate(self.thruster_lags):
            l._prepare(dt)
            l._x = lag[0, 3 * i:3 * i + 3].copy()

    def compute_thruster_forces(self, u_thrust, dt):
        lag = self._lag_host()
        tau = self.engine("f64").thruster_wrench_host(np.asarray(u_thrust, dtype=float).reshape(1, 8), lag=lag, dt=dt)[0]
        self._store_lag_host(lag, dt)
        return tau

    def dynamics(self, x, u_thrust, dt):
        lag = self._lag_host()
        xd = self._dynamics_one(np.asarray(x, dtype=float)[:12], np.asarray(u_thrust, dtype=float)[:8], 12, 8, dt, lag=lag)
        self._store_lag_host(lag, dt)
        return xd

    def _n_tether_states(self):
        return 0

    def dynamics_with_tether(self, x, u_thrust, dt):
        raise NotImplementedError("the tether model is not part of the B200 engine (default-off in the reference)")
