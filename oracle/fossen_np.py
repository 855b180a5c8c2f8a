"""CPU ORACLE (test infrastructure, NOT product code) — numpy float64 restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The product
path (bluerov2_dynamics_b200/) never does: it fails loudly when the CUDA library is missing.

What is restated (all paths relative to the reference checkout, ViktorNfa/bluerov2_dynamics):
  fossen/BlueROV2.py          8-thruster Fossen model, T200 polynomial, 3rd-order ThrusterLag (stateful)
  fossen/BlueROV2_thrust.py   wrench-input 12-state model
  fossen/BlueROV2_wrench.py   wrench-input 13-state quaternion model + quaternion helpers
  fossen/bluerov_torch.py     reduced 9-state RHS
  training/*.py               simulate_physics (RK4 / Euler), one_step_rmse_physics,
                              multistep_rmse_endpoint_physics

Parity pinning: every function here is checked in tests/test_oracle_golden.py against
tests/golden/reference_vectors.npz, which holds outputs of the UNMODIFIED reference executed by
tests/golden/make_golden.py.  One third-party boundary is restated from its published algorithm instead of
from reference source: scipy.signal.cont2discrete(method="zoh") -> scipy.linalg.expm (call site
fossen/BlueROV2.py:490-501; reference lock file pins scipy 1.15.3 / 1.17.0).  ZOH is
expm([[A, B], [0, 0]] * dt) -> Ad = top-left, Bd = top-right; `lag_zoh` below evaluates that exponential by
scaling-and-squaring of a Taylor series in extended precision and is pinned against the golden (Ad, Bd)
produced by scipy 1.18.1 at dt = 0.01 / 0.02 / 0.05.

Unlike the reference (one vehicle per call, 6x6 matrices allocated per call) every function is batched over
a leading vehicle axis N; N = 1 reproduces a reference call.  The stateful lag of the reference
(`ThrusterLag._x`, advanced once per dynamics() call, i.e. 4x per RK4 step — fossen/BlueROV2.py:258,503-510)
is an explicit array `lag[N, 8, 3]` passed in and returned.

Extension without a reference counterpart ("parity unpinned", SURVEY trap T11): the optional first-order
wrench lag  tau_dot = (tau_cmd - tau) / T_lag  of `rhs_wrench*_lag1`; it is defined here and only here.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------------------------

#: continuous-time thruster lag, fossen/BlueROV2.py:476-481
LAG_AC = np.array([[-89.0, -72.33, -26.54], [128.0, 0.0, 0.0], [0.0, 32.0, 0.0]])
LAG_BC = np.array([8.0, 0.0, 0.0])
LAG_CC = np.array([0.0, 5.992, 3.317])

#: T200 static thrust curve, odd powers 1,3,5,7,9 — fossen/BlueROV2.py:251-257
POLY = (8.9, 176.0, -404.1, 389.9, -140.3)


def default_params(rho: float = 1000.0, current=(0.0, 0.0, 0.0)) -> dict:
    """Physical constants of the three 6-DOF classes (fossen/BlueROV2.py:81-150 = BlueROV2_thrust.py:82-147
    = BlueROV2_wrench.py:160-225).  Values may later be replaced by arrays of shape [N] (Monte-Carlo)."""
    g = 9.82
    m = 13.5
    vol = 0.0134
    p = dict(
        m=m, W=m * g, B=rho * g * vol,
        xb=0.0, yb=0.0, zb=-0.01,
        Ix=0.26, Iy=0.23, Iz=0.37,
        Xu_dot=-6.36, Yv_dot=-7.12, Zw_dot=-18.68, Kp_dot=-0.189, Mq_dot=-0.135, Nr_dot=-0.222,
        Xu=-13.7, Yv=-0.0, Zw=-33.0, Kp=-0.0, Mq=-0.8, Nr=-0.0,
        Xu_abs=-141.0, Yv_abs=-217.0, Zw_abs=-190.0, Kp_abs=-1.19, Mq_abs=-0.47, Nr_abs=-1.5,
        current=np.asarray(current, float),
    )
    p["Minv"] = minv_diag(p)
    return p


def minv_diag(p: dict):
    """diag(inv(MRB + MA)); M is diagonal (fossen/BlueROV2.py:103-126).  Returned as [..., 6]."""
    d = [p["m"] - p["Xu_dot"], p["m"] - p["Yv_dot"], p["m"] - p["Zw_dot"],
         p["Ix"] - p["Kp_dot"], p["Iy"] - p["Mq_dot"], p["Iz"] - p["Nr_dot"]]
    return 1.0 / np.stack(np.broadcast_arrays(*[np.asarray(v, float) for v in d]), axis=-1)


def _rotz(a):
    s, c = np.sin(a), np.cos(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def thruster_geometry():
    """Lever arms r_i, directions e_i and the 6x8 allocation matrix (fossen/BlueROV2.py:159-232, 265-278).
    Placement angles are the paper's rounded radians (trap T6) — never symmetrised."""
    r_h = np.array([0.156, 0.111, 0.085])
    r_v = np.array([0.12, 0.218, 0.0])
    e_h = np.array([1.0 / np.sqrt(2), -1.0 / np.sqrt(2), 0.0])
    ang_r = [0.0, 5.05, 1.91, np.pi, 0.0, 4.15, 1.01, np.pi]
    ang_e = [0.0, np.pi / 2, 3 * np.pi / 2, np.pi]
    r = np.zeros((8, 3))
    e = np.zeros((8, 3))
    for i in range(4):
        r[i] = _rotz(ang_r[i]) @ r_h
        e[i] = _rotz(ang_e[i]) @ e_h
    for i in range(4, 8):
        r[i] = _rotz(ang_r[i]) @ r_v
        e[i] = (0.0, 0.0, -1.0)
    alloc = np.zeros((6, 8))
    alloc[0:3] = e.T
    alloc[3:6] = np.cross(r, e).T
    return r, e, alloc


def lag_zoh(dt: float):
    """(Ad, Bd) of the zero-order-hold discretisation of (LAG_AC, LAG_BC) — what
    scipy.signal.cont2discrete(..., method='zoh') returns at fossen/BlueROV2.py:494-495.
    exp(M) with M = [[A, B], [0, 0]] dt by scaling-and-squaring: M / 2^s, 30-term Taylor, square s times,
    in np.longdouble (x87 80-bit on this platform) so the float64 result is correctly rounded to ~1 ulp."""
    M = np.zeros((4, 4), dtype=np.longdouble)
    M[:3, :3] = LAG_AC
    M[:3, 3] = LAG_BC
    M *= np.longdouble(dt)
    nrm = float(np.abs(M).sum(axis=1).max())
    s = max(0, int(np.ceil(np.log2(max(nrm, 1e-300)))) + 4)
    Ms = M / np.longdouble(2.0 ** s)
    E = np.eye(4, dtype=np.longdouble)
    term = np.eye(4, dtype=np.longdouble)
    for k in range(1, 30):
        term = term @ Ms / np.longdouble(k)
        E = E + term
    for _ in range(s):
        E = E @ E
    E = E.astype(np.float64)
    return E[:3, :3].copy(), E[:3, 3].copy()


# --------------------------------------------------------------------------------------------
# building blocks (batched)
# --------------------------------------------------------------------------------------------

def thrust_poly(V):
    """Static T200 curve F(V) (fossen/BlueROV2.py:251-257).  Uses ** like the reference (libm pow)."""
    V = np.asarray(V, float)
    return POLY[4] * V ** 9 + POLY[3] * V ** 7 + POLY[2] * V ** 5 + POLY[1] * V ** 3 + POLY[0] * V


def lag_step(lag, F, Ad, Bd):
    """One ThrusterLag.step for every thruster (fossen/BlueROV2.py:503-510): x <- Ad x + Bd F; y = Cc x.
    lag [N,8,3], F [N,8] -> (lag_new [N,8,3], y [N,8])."""
    new = lag @ Ad.T + F[..., None] * Bd
    return new, new @ LAG_CC


def _trig(phi, th, psi):
    return np.sin(phi), np.cos(phi), np.sin(th), np.cos(th), np.sin(psi), np.cos(psi)


def rot_b2n(phi, th, psi):
    """R_{b->n} = Rz(psi) Ry(theta) Rx(phi)  (fossen/BlueROV2.py:23-41) as [N,3,3]."""
    sphi, cphi, sth, cth, spsi, cpsi = _trig(phi, th, psi)
    R = np.empty(np.shape(phi) + (3, 3))
    R[..., 0, 0] = cpsi * cth
    R[..., 0, 1] = -spsi * cphi + cpsi * sth * sphi
    R[..., 0, 2] = spsi * sphi + cpsi * cphi * sth
    R[..., 1, 0] = spsi * cth
    R[..., 1, 1] = cpsi * cphi + sphi * sth * spsi
    R[..., 1, 2] = -cpsi * sphi + sth * spsi * cphi
    R[..., 2, 0] = -sth
    R[..., 2, 1] = cth * sphi
    R[..., 2, 2] = cth * cphi
    return R


def euler_rates(phi, th, w, eps=1e-7):
    """J2(phi, theta) @ [p,q,r] with the reference's cos(theta) clamp (fossen/BlueROV2.py:43-62, trap T7:
    eps*sign(c) and sign(0) = 0, so an exactly-zero cosine still divides by zero)."""
    sphi, cphi, sth, cth = np.sin(phi), np.cos(phi), np.sin(th), np.cos(th)
    cth = np.where(np.abs(cth) < eps, eps * np.sign(cth), cth)
    with np.errstate(divide="ignore", invalid="ignore"):
        tth = sth / cth
        p, q, r = w[..., 0], w[..., 1], w[..., 2]
        out = np.stack([p + sphi * tth * q + cphi * tth * r,
                        cphi * q - sphi * r,
                        sphi / cth * q + cphi / cth * r], axis=-1)
    return out


def coriolis_times_nu(nu, p):
    """(C_RB(nu) + C_A(nu)) @ nu, term by term as the 6x6 of fossen/BlueROV2.py:280-325 (the paper's [3,4]
    and [4,3] entries corrected to +-Iz r as the reference does)."""
    u, v, w, pp, q, r = (nu[..., i] for i in range(6))
    m = p["m"]
    a1u = m * u - p["Xu_dot"] * u
    a2v = m * v - p["Yv_dot"] * v
    a3w = m * w - p["Zw_dot"] * w
    b1p = p["Ix"] * pp - p["Kp_dot"] * pp
    b2q = p["Iy"] * q - p["Mq_dot"] * q
    b3r = p["Iz"] * r - p["Nr_dot"] * r
    return np.stack([
        a3w * q - a2v * r,
        -a3w * pp + a1u * r,
        a2v * pp - a1u * q,
        a3w * v - a2v * w + b3r * q - b2q * r,
        -a3w * u + a1u * w - b3r * pp + b1p * r,
        a2v * u - a1u * v + b2q * pp - b1p * q,
    ], axis=-1)


def damping_times_nu(nur, p):
    """diag(-L_i - Q_i |nu_r,i|) nu_r  (fossen/BlueROV2.py:327-338)."""
    lin = np.stack(np.broadcast_arrays(*[np.asarray(p[k], float) for k in ("Xu", "Yv", "Zw", "Kp", "Mq", "Nr")]), -1)
    quad = np.stack(np.broadcast_arrays(
        *[np.asarray(p[k], float) for k in ("Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs")]), -1)
    return (-lin - quad * np.abs(nur)) * nur


def restoring(sth, cth_sphi, cth_cphi, p):
    """g(eta) (fossen/BlueROV2.py:340-355; BlueROV2_wrench.py:293-319 feeds the same three terms from R)."""
    wmb = p["W"] - p["B"]
    xbB, ybB, zbB = p["xb"] * p["B"], p["yb"] * p["B"], p["zb"] * p["B"]
    return np.stack(np.broadcast_arrays(
        wmb * sth, -wmb * cth_sphi, -wmb * cth_cphi,
        ybB * cth_cphi - zbB * cth_sphi,
        -zbB * sth - xbB * cth_cphi,
        xbB * cth_sphi + ybB * sth), axis=-1)


def _nu_dot(nu, R, tau, sth, cs, cc, p):
    cur = np.asarray(p["current"], float)
    vcb = np.einsum("...ji,...j->...i", R, np.broadcast_to(cur, nu[..., :3].shape))  # R^T v_c
    nur = nu.copy()
    nur[..., :3] -= vcb
    rhs = tau - coriolis_times_nu(nu, p) - damping_times_nu(nur, p) - restoring(sth, cs, cc, p)
    return p["Minv"] * rhs


# --------------------------------------------------------------------------------------------
# state derivatives
# --------------------------------------------------------------------------------------------

def rhs_wrench12(x, tau, p):
    """BlueROV2_thrust.BlueROV2.dynamics (fossen/BlueROV2_thrust.py:235-282).  x [N,12], tau [N,6]."""
    x = np.asarray(x, float)
    phi, th, psi = x[..., 3], x[..., 4], x[..., 5]
    nu = x[..., 6:12]
    R = rot_b2n(phi, th, psi)
    sth, cth = np.sin(th), np.cos(th)
    nud = _nu_dot(nu, R, np.asarray(tau, float), sth, cth * np.sin(phi), cth * np.cos(phi), p)
    pd = np.einsum("...ij,...j->...i", R, nu[..., :3])
    return np.concatenate([pd, euler_rates(phi, th, nu[..., 3:6]), nud], axis=-1)


def thruster_wrench(u, lag, Ad, Bd, alloc):
    """compute_thruster_forces (fossen/BlueROV2.py:265-278): polynomial -> ONE lag step -> allocation.
    Returns (tau [N,6], lag_new)."""
    lag_new, Fdyn = lag_step(lag, thrust_poly(u), Ad, Bd)
    return Fdyn @ alloc.T, lag_new


def rhs_thruster(x, u, lag, Ad, Bd, alloc, p):
    """BlueROV2.dynamics (fossen/BlueROV2.py:357-400) — advances the lag state once (trap T2)."""
    tau, lag_new = thruster_wrench(np.asarray(u, float), lag, Ad, Bd, alloc)
    return rhs_wrench12(x, tau, p), lag_new


def quat_normalize(q, eps=1e-12):
    """fossen/BlueROV2_wrench.py:27-36 (identity fallback below eps)."""
    q = np.asarray(q, float)
    n = np.sqrt(np.sum(q * q, axis=-1, keepdims=True))
    ident = np.zeros_like(q)
    ident[..., 0] = 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(n < eps, ident, q / n)


def quat_to_R(q):
    """fossen/BlueROV2_wrench.py:39-53 (normalises first)."""
    q = quat_normalize(q)
    qw, qx, qy, qz = (q[..., i] for i in range(4))
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1.0 - 2.0 * (qy * qy + qz * qz)
    R[..., 0, 1] = 2.0 * (qx * qy - qz * qw)
    R[..., 0, 2] = 2.0 * (qx * qz + qy * qw)
    R[..., 1, 0] = 2.0 * (qx * qy + qz * qw)
    R[..., 1, 1] = 1.0 - 2.0 * (qx * qx + qz * qz)
    R[..., 1, 2] = 2.0 * (qy * qz - qx * qw)
    R[..., 2, 0] = 2.0 * (qx * qz - qy * qw)
    R[..., 2, 1] = 2.0 * (qy * qz + qx * qw)
    R[..., 2, 2] = 1.0 - 2.0 * (qx * qx + qy * qy)
    return R


def quat_multiply(a, b):
    """Hamilton product, scalar first (fossen/BlueROV2_wrench.py:56-68)."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=-1)


def quat_derivative(q, omega):
    """0.5 q (x) [0, omega]  (fossen/BlueROV2_wrench.py:71-80)."""
    omega = np.asarray(omega, float)
    oq = np.concatenate([np.zeros(omega.shape[:-1] + (1,)), omega], axis=-1)
    return 0.5 * quat_multiply(q, oq)


def euler_to_quat(phi, th, psi):
    """fossen/BlueROV2_wrench.py:86-106."""
    c1, s1 = np.cos(phi * 0.5), np.sin(phi * 0.5)
    c2, s2 = np.cos(th * 0.5), np.sin(th * 0.5)
    c3, s3 = np.cos(psi * 0.5), np.sin(psi * 0.5)
    return quat_normalize(np.stack([c3 * c2 * c1 + s3 * s2 * s1, c3 * c2 * s1 - s3 * s2 * c1,
                                    c3 * s2 * c1 + s3 * c2 * s1, s3 * c2 * c1 - c3 * s2 * s1], axis=-1))


def quat_to_euler(q):
    """fossen/BlueROV2_wrench.py:109-132."""
    q = quat_normalize(q)
    qw, qx, qy, qz = (q[..., i] for i in range(4))
    phi = np.arctan2(2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx * qx + qy * qy))
    th = np.arcsin(np.clip(2.0 * (qw * qy - qz * qx), -1.0, 1.0))
    psi = np.arctan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz))
    return np.stack([phi, th, psi], axis=-1)


def rhs_quat13(x, tau, p):
    """BlueROV2_wrench.BlueROV2.dynamics (fossen/BlueROV2_wrench.py:322-367)."""
    x = np.asarray(x, float)
    q = quat_normalize(x[..., 3:7])
    nu = x[..., 7:13]
    R = quat_to_R(q)
    nud = _nu_dot(nu, R, np.asarray(tau, float), -R[..., 2, 0], R[..., 2, 1], R[..., 2, 2], p)
    pd = np.einsum("...ij,...j->...i", R, nu[..., :3])
    return np.concatenate([pd, quat_derivative(q, nu[..., 3:6]), nud], axis=-1)


# reduced model constants, fossen/parameters.py:3-33
RED = dict(m=11.4, g=9.82, F_bouy=1026 * 0.0115 * 9.82, X_ud=-2.6, Y_vd=-18.5, Z_wd=-13.3, N_rd=-0.28,
           I_zz=0.245, X_u=-0.09, Y_v=-0.26, Z_w=-0.19, N_r=-4.64, X_uc=-34.96, Y_vc=-103.25, Z_wc=-74.23,
           N_rc=-0.43)


def rhs_reduced9(x, u, dtype=np.float64):
    """bluerov_compute (fossen/bluerov_torch.py:20-67): x [B,9] = [x,y,z,cos psi,sin psi,u,v,w,r], u [B,4]."""
    x = np.atleast_2d(np.asarray(x, dtype))
    u = np.atleast_2d(np.asarray(u, dtype))
    c = {k: dtype(v) for k, v in RED.items()}
    cps, sps = x[:, 3], x[:, 4]
    uu, v, w, r = x[:, 5], x[:, 6], x[:, 7], x[:, 8]
    X, Y, Z, Mz = u[:, 0], u[:, 1], u[:, 2], u[:, 3]
    one = dtype(1)
    return np.stack([
        cps * uu - sps * v, sps * uu + cps * v, w, -sps * r, cps * r,
        one / (c["m"] - c["X_ud"]) * (X + (c["m"] - c["Y_vd"]) * v * r + (c["X_u"] + c["X_uc"] * np.abs(uu)) * uu),
        one / (c["m"] - c["Y_vd"]) * (Y - (c["m"] - c["X_ud"]) * uu * r + (c["Y_v"] + c["Y_vc"] * np.abs(v)) * v),
        one / (c["m"] - c["Z_wd"]) * (Z + (c["Z_w"] + c["Z_wc"] * np.abs(w)) * w + c["m"] * c["g"] - c["F_bouy"]),
        one / (c["I_zz"] - c["N_rd"]) * (Mz - (c["X_ud"] - c["Y_vd"]) * uu * v + (c["N_r"] + c["N_rc"] * np.abs(r)) * r),
    ], axis=1).astype(dtype)


def ssa(a):
    """fossen/bluerov_torch.py:8-18."""
    a = np.asarray(a)
    return a - 2 * np.pi * np.floor((a + np.pi) / (2 * np.pi))


# --------------------------------------------------------------------------------------------
# integrators, rollout, evaluators
# --------------------------------------------------------------------------------------------

class Model:
    """One of 'thruster8' | 'wrench12' | 'quat13' with its constants; `f(x, u, lag)` -> (xdot, lag_new).

    lag1_T: optional first-order wrench lag time constant(s) (EXTENSION, parity unpinned): the state is
    augmented with the 6 filtered wrench components, tau_dot = (u - tau) / lag1_T."""

    def __init__(self, kind, dt, params=None, lag1_T=None):
        self.kind = kind
        self.dt = float(dt)
        self.p = default_params() if params is None else params
        self.lag1_T = lag1_T
        self.nx = 13 if kind == "quat13" else 12
        self.nu = 8 if kind == "thruster8" else 6
        if kind == "thruster8":
            self.Ad, self.Bd = lag_zoh(self.dt)
            self.alloc = thruster_geometry()[2]
        if lag1_T is not None:
            assert kind != "thruster8"
            self.nx += 6

    def f(self, x, u, lag):
        if self.kind == "thruster8":
            return rhs_thruster(x, u, lag, self.Ad, self.Bd, self.alloc, self.p)
        base = 13 if self.kind == "quat13" else 12
        fn = rhs_quat13 if self.kind == "quat13" else rhs_wrench12
        if self.lag1_T is None:
            return fn(x, u, self.p), lag
        tau = x[..., base:base + 6]
        Tl = np.asarray(self.lag1_T, float)
        Tl = Tl[..., None] if Tl.ndim else Tl
        return np.concatenate([fn(x[..., :base], tau, self.p), (u - tau) / Tl], axis=-1), lag

    def post(self, x):
        """Per-step quaternion re-normalisation of training/train_tank_brov2_wrench_quat.py:262-263."""
        if self.kind == "quat13":
            x = x.copy()
            x[..., 3:7] = quat_normalize(x[..., 3:7])
        return x

    def zero_lag(self, n):
        return np.zeros((n, 8, 3))


def step(model: Model, integ: str, x, u, lag):
    """One integrator step.  RK4 as training/train_tank_brov2_rk4.py:385-394 (input held over the step, the
    lag state advanced by each of the four dynamics() calls); Euler as
    training/train_tank_brov2_full_comparison.py:462-465."""
    dt = model.dt
    if integ == "euler":
        k, lag = model.f(x, u, lag)
        return model.post(x + dt * k), lag
    k1, lag = model.f(x, u, lag)
    k2, lag = model.f(x + 0.5 * dt * k1, u, lag)
    k3, lag = model.f(x + 0.5 * dt * k2, u, lag)
    k4, lag = model.f(x + dt * k3, u, lag)
    return model.post(x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)), lag


def rollout(model: Model, integ: str, x0, U, lag0=None, stride=0):
    """simulate_physics for N vehicles at once.  x0 [N,nx]; U [T,N,nu] (per vehicle) or [T,nu] (shared).
    Returns (snapshots [S,N,nx] after steps stride, 2*stride, ... ; x_T [N,nx]; lag_T [N,8,3])."""
    x = np.array(x0, float, ndmin=2)
    N = x.shape[0]
    lag = model.zero_lag(N) if lag0 is None else np.array(lag0, float).reshape(N, 8, 3)
    U = np.asarray(U, float)
    snaps = []
    for k in range(U.shape[0]):
        uk = U[k] if U.ndim == 3 else np.broadcast_to(U[k], (N, U.shape[1]))
        x, lag = step(model, integ, x, uk, lag)
        if stride and (k + 1) % stride == 0:
            snaps.append(x.copy())
    return (np.array(snaps) if snaps else np.zeros((0, N, model.nx))), x, lag


def multistep_se(model: Model, integ: str, X, U, horizons, lag_mode="reset"):
    """multistep_rmse_endpoint_physics (training/train_tank_brov2_rk4.py:399-417 and twins): sliding windows
    k = 0..T-H-1, H-step open-loop rollout from X[k] under U[k:k+H], squared endpoint error vs X[k+H].
    Returns {H: (sum_sq_err, n_windows, rmse)}; rmse = sqrt(se / (n_windows * n_states)), NaN if T <= H.

    lag_mode = 'reset': every window starts from zero lag state (window-parallel semantics of the engine).
    lag_mode = 'carry': the reference's literal behaviour — ONE model object for all windows, so window k
                        starts from the lag state left by window k-1 (trap T3); sequential by nature."""
    X = np.asarray(X, float)
    U = np.asarray(U, float)
    T, nx = X.shape
    out = {}
    for H in horizons:
        ns = T - H
        if ns <= 0:
            out[H] = (0.0, 0, float("nan"))
            continue
        if lag_mode == "reset":
            x = X[:ns].copy()
            lag = model.zero_lag(ns)
            for j in range(H):
                x, lag = step(model, integ, x, U[j:j + ns], lag)
            err = x - X[H:H + ns]
            se = float(np.sum(err * err))
        else:
            lag = model.zero_lag(1)
            se = 0.0
            for k in range(ns):
                x = X[k:k + 1].copy()
                for j in range(H):
                    x, lag = step(model, integ, x, U[k + j:k + j + 1], lag)
                e = x[0] - X[k + H]
                se += float(e @ e)
        out[H] = (se, ns, float(np.sqrt(se / (ns * nx))))
    return out


def one_step_rmse(model: Model, X, U, lag_mode="reset"):
    """one_step_rmse_physics (training/train_tank_brov2_koopmanEDMDc.py:237-247, wrench twins): teacher-forced
    single Euler step from every row, rmse over rows 1..T-1."""
    X = np.asarray(X, float)
    U = np.asarray(U, float)
    T = X.shape[0]
    if lag_mode == "reset":
        pred, _ = step(model, "euler", X[:-1].copy(), U[:-1], model.zero_lag(T - 1))
    else:
        lag = model.zero_lag(1)
        rows = []
        for k in range(T - 1):
            xn, lag = step(model, "euler", X[k:k + 1].copy(), U[k:k + 1], lag)
            rows.append(xn[0])
        pred = np.array(rows)
    return float(np.sqrt(np.mean((X[1:] - pred) ** 2)))


def smooth_inputs(rng, T, nu, n=None, scale=1.0, alpha=0.98, sigma=0.02):
    """Reference "random but smooth" command generator u_k = clip(alpha u_{k-1} + sigma N(0,1), -1, 1)
    (training/train_sim_brov2_koopmanEDMDc.py:161-164), batched: returns [T, n, nu] (or [T, nu] if n is None)."""
    shape = (nu,) if n is None else (n, nu)
    U = np.zeros((T,) + shape)
    u = np.zeros(shape)
    for k in range(T):
        u = np.clip(alpha * u + sigma * rng.standard_normal(shape), -1.0, 1.0)
        U[k] = u
    return U * np.asarray(scale)
