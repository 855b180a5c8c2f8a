#!/usr/bin/env python3
"""Generate the frozen golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Everything written here is an *output of the reference's own code* (numpy float64 / torch) on seeded
synthetic inputs; the inputs are stored next to the outputs so that the oracle (oracle/) and the
CUDA engine consume exactly the same arrays.  Nothing from this repository's package is imported.

Reference entry points exercised (paths relative to /root/reference):
  fossen/BlueROV2.py:79,245-278,357-400,464-510   8-thruster model, ThrusterLag, dynamics
  fossen/BlueROV2_thrust.py:235-282               wrench 12-state dynamics
  fossen/BlueROV2_wrench.py:27-138,322-367        quaternion model + helpers
  fossen/bluerov_torch.py:8-67                    reduced 9-state RHS, ssa
  training/train_tank_brov2_rk4.py:375-417        RK4 simulate_physics / multistep_rmse_endpoint_physics
  training/train_tank_brov2_full_comparison.py:453-487   Euler twins (8-thruster)
  training/train_tank_brov2_koopmanEDMDc.py:237-247      one_step_rmse_physics (8-thruster)
  training/train_tank_brov2_wrench_comp.py:208-250       Euler twins (wrench 12)
  training/train_tank_brov2_wrench_quat.py:249-297       Euler twins (quaternion 13)
  training/train_sim_brov2_koopmanEDMDc.py:153-197       data generator (restated: the module body
                                                         runs 240 000 steps at import time)
"""
import importlib.util
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

sys.dont_write_bytecode = True
# make sure the reference's `fossen` is the one that gets imported, never this repo's mirror
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(OUT, "..", ".."))]
for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
           "matplotlib.lines", "matplotlib.cm", "matplotlib.colors"):
    sys.modules.setdefault(_m, MagicMock())

import torch  # noqa: E402
from fossen.BlueROV2 import BlueROV2 as RefThruster, ThrusterLag  # noqa: E402
from fossen.BlueROV2_thrust import BlueROV2 as RefWrench12  # noqa: E402
import fossen.BlueROV2_wrench as refq  # noqa: E402
from fossen.bluerov_torch import bluerov_compute, ssa  # noqa: E402

assert RefThruster.__module__ == "fossen.BlueROV2" and sys.modules["fossen"].__path__[0].startswith(REF)


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_rk4 = _load("ref_rk4", "training/train_tank_brov2_rk4.py")
ref_cmp = _load("ref_cmp", "training/train_tank_brov2_full_comparison.py")
ref_koop = _load("ref_koop", "training/train_tank_brov2_koopmanEDMDc.py")
ref_wc = _load("ref_wc", "training/train_tank_brov2_wrench_comp.py")
ref_wq = _load("ref_wq", "training/train_tank_brov2_wrench_quat.py")


# ----------------------------------------------------------------------------------------------
# synthetic input generators (seeded; the arrays are stored, so the generators need not be portable)
# ----------------------------------------------------------------------------------------------
def smooth_inputs(rng, T, nu, scale=1.0, alpha=0.98, sigma=0.02, u0=None):
    """u_k = clip(alpha*u_{k-1} + sigma*N(0,1), -1, 1) * scale  (train_sim_brov2_koopmanEDMDc.py:161-164)."""
    U = np.zeros((T, nu))
    u = np.zeros(nu) if u0 is None else np.array(u0, float)
    for k in range(T):
        u = np.clip(alpha * u + sigma * rng.standard_normal(nu), -1.0, 1.0)
        U[k] = u
    return U * np.asarray(scale)


def random_state12(rng, n):
    x = np.zeros((n, 12))
    x[:, 0:2] = rng.uniform(-2, 2, (n, 2))
    x[:, 2] = rng.uniform(0, 3, n)
    x[:, 3:5] = rng.uniform(-0.2, 0.2, (n, 2))
    x[:, 5] = rng.uniform(-np.pi, np.pi, n)
    x[:, 6:9] = rng.uniform(-0.5, 0.5, (n, 3))
    x[:, 9:12] = rng.uniform(-0.3, 0.3, (n, 3))
    return x


def state12_to_13(x12):
    out = np.zeros((x12.shape[0], 13))
    out[:, 0:3] = x12[:, 0:3]
    for i in range(x12.shape[0]):
        out[i, 3:7] = refq.euler_to_quat(*x12[i, 3:6])
    out[:, 7:13] = x12[:, 6:12]
    return out


WRENCH_SCALE = np.array([40.0, 40.0, 40.0, 5.0, 5.0, 5.0])


def rk4_generic(f, x, u, dt):
    k1 = f(x, u, dt)
    k2 = f(x + 0.5 * dt * k1, u, dt)
    k3 = f(x + 0.5 * dt * k2, u, dt)
    k4 = f(x + dt * k3, u, dt)
    return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def lag_states(rov):
    return np.stack([np.array(l._x, float).reshape(3) for l in rov.thruster_lags])  # [8,3]


def set_lag_states(rov, L):
    for i, l in enumerate(rov.thruster_lags):
        l._x = np.array(L[i], float)


G = {}

# ----------------------------------------------------------------------------------------------
# 1. constants
# ----------------------------------------------------------------------------------------------
rov = RefThruster()
G["const_Minv_diag"] = np.diag(rov.Minv).copy()
G["const_M_diag"] = np.diag(rov.M).copy()
G["const_W"] = np.array(rov.W)
G["const_B"] = np.array(rov.B)
Tal = np.zeros((6, 8))
for i, th in enumerate(rov.thrusters_r):
    Tal[0:3, i] = th["dir"]
    Tal[3:6, i] = np.cross(th["r"], th["dir"])
G["const_alloc"] = Tal
G["const_thr_r"] = np.stack([t["r"] for t in rov.thrusters_r])
G["const_thr_dir"] = np.stack([t["dir"] for t in rov.thrusters_r])
for dt in (0.01, 0.02, 0.05):
    Ad, Bd = ThrusterLag._discretise(ThrusterLag._Ac, ThrusterLag._Bc, ThrusterLag._Cc, ThrusterLag._Dc, dt)
    G[f"const_lag_Ad_{dt}"] = Ad
    G[f"const_lag_Bd_{dt}"] = Bd[:, 0]
Vs = np.linspace(-1, 1, 41)
G["const_poly_V"] = Vs
G["const_poly_F"] = np.array([rov._old_thruster_force_from_input(np.float64(v)) for v in Vs])

# ----------------------------------------------------------------------------------------------
# 2. single RHS evaluations (known-answer), all four models
# ----------------------------------------------------------------------------------------------
rng = np.random.default_rng(100)
n = 64
X = random_state12(rng, n)
X[0] = 0.0
X[0, 2] = 5.0
U8 = rng.uniform(-1, 1, (n, 8))
U8[0] = [0.1, 0.1, 0.1, 0.0, 0.5, 0.5, 0.5, 0.5]
L0 = rng.uniform(-0.5, 0.5, (n, 8, 3))
L0[: n // 2] = 0.0  # first half: fresh lag
cur = np.array([0.3, -0.2, 0.1])
for tag, current in (("", np.zeros(3)), ("_cur", cur)):
    XD = np.zeros((n, 12))
    L1 = np.zeros((n, 8, 3))
    for i in range(n):
        r = RefThruster(current_speed=current.copy())
        set_lag_states(r, L0[i])
        XD[i] = r.dynamics(X[i].copy(), U8[i].copy(), 0.02)
        L1[i] = lag_states(r)
    G[f"rhs_thr{tag}_xdot"] = XD
    G[f"rhs_thr{tag}_lag1"] = L1
G["rhs_thr_x"] = X
G["rhs_thr_u"] = U8
G["rhs_thr_lag0"] = L0
G["rhs_current"] = cur

TAU = rng.uniform(-1, 1, (n, 6)) * WRENCH_SCALE
TAU[0] = [1, 0.5, -0.2, 0.01, 0.02, 0.03]
G["rhs_w_tau"] = TAU
for tag, current in (("", None), ("_cur", cur)):
    r12 = RefWrench12(current_speed=current)
    G[f"rhs_w12{tag}_xdot"] = np.stack([r12.dynamics(X[i], TAU[i]) for i in range(n)])
X13 = state12_to_13(X)
X13[n // 2:, 3:7] *= rng.uniform(0.5, 1.5, (n - n // 2, 1))  # un-normalised quaternions exercise trap T9
G["rhs_q13_x"] = X13
for tag, current in (("", None), ("_cur", cur)):
    r13 = refq.BlueROV2(current_speed=current)
    G[f"rhs_q13{tag}_xdot"] = np.stack([r13.dynamics(X13[i], TAU[i]) for i in range(n)])

# quaternion helpers
Q = rng.standard_normal((16, 4))
Q2 = rng.standard_normal((16, 4))
W3 = rng.standard_normal((16, 3))
G["quat_q"] = Q
G["quat_q2"] = Q2
G["quat_w"] = W3
G["quat_normalize"] = np.stack([refq.quat_normalize(q) for q in Q])
G["quat_normalize_tiny"] = refq.quat_normalize(np.array([1e-13, 0, 0, 0]))
G["quat_to_R"] = np.stack([refq.quat_to_rotation_matrix(q) for q in Q])
G["quat_multiply"] = np.stack([refq.quat_multiply(a, b) for a, b in zip(Q, Q2)])
G["quat_derivative"] = np.stack([refq.quat_derivative(a, w) for a, w in zip(Q, W3)])
G["quat_to_euler"] = np.stack([np.array(refq.quat_to_euler(q)) for q in Q])
G["quat_to_yaw"] = np.array([refq.quat_to_yaw(q) for q in Q])
E = rng.uniform(-1.2, 1.2, (16, 3))
G["quat_euler_in"] = E
G["quat_euler_to_quat"] = np.stack([refq.euler_to_quat(*e) for e in E])

# reduced 9-state torch model
x9 = rng.standard_normal((256, 9))
c = rng.uniform(-np.pi, np.pi, 256)
x9[:, 3] = np.cos(c)
x9[:, 4] = np.sin(c)
u4 = rng.uniform(-30, 30, (256, 4))
x9[0] = [0, 0, 0, 1, 0, 0, 0, 0, 0]
u4[0] = [1, 1, 1, 1]
G["red9_x"] = x9
G["red9_u"] = u4
G["red9_xdot_f64"] = bluerov_compute(0.0, torch.tensor(x9), torch.tensor(u4)).numpy()
G["red9_xdot_f32"] = bluerov_compute(0.0, torch.tensor(x9, dtype=torch.float32),
                                     torch.tensor(u4, dtype=torch.float32)).numpy()
ang = rng.uniform(-20, 20, 64)
G["ssa_in"] = ang
G["ssa_out"] = ssa(torch.tensor(ang)).numpy()

# ----------------------------------------------------------------------------------------------
# 3. config-1 KAT: 8-thruster, RK4, dt=0.02, 1000 steps (fossen/test_ode.py inputs; RK4 loop of
#    training/train_tank_brov2_rk4.py:385-394), constant input and smooth time-varying input
# ----------------------------------------------------------------------------------------------
x0 = np.zeros(12)
x0[2] = 5.0
uc = np.array([0.1, 0.1, 0.1, 0.0, 0.5, 0.5, 0.5, 0.5])
dt = 0.02
r = RefThruster()
traj = ref_rk4.simulate_physics(x0, np.tile(uc, (1000, 1)), dt, r)
G["cfg1_x0"] = x0
G["cfg1_u_const"] = uc
G["cfg1_const_traj_s10"] = traj[::10].copy()  # rows 0,10,...,1000
G["cfg1_const_lagT"] = lag_states(r)
rng = np.random.default_rng(0)
Uv = smooth_inputs(rng, 1000, 8)
r = RefThruster()
traj = ref_rk4.simulate_physics(x0, Uv, dt, r)
G["cfg1_U_var"] = Uv
G["cfg1_var_traj_s10"] = traj[::10].copy()
G["cfg1_var_lagT"] = lag_states(r)
# Euler twin (full_comparison simulate_physics), dt = 0.01 as fossen/test_euler.py
r = RefThruster(dt=0.01)
traj = ref_cmp.simulate_physics(x0, np.tile(uc, (500, 1)), 0.01, r)
G["cfg1_euler_dt001_traj_s10"] = traj[::10].copy()
G["cfg1_euler_dt001_lagT"] = lag_states(r)

# ----------------------------------------------------------------------------------------------
# 4. small ensembles, all models, RK4 + Euler, smooth inputs, random x0 (cfg2/cfg3 in miniature)
# ----------------------------------------------------------------------------------------------
rng = np.random.default_rng(1)
nv, T = 6, 400
X0 = random_state12(rng, nv)
X0[:, 6:] = 0.0
Uens = np.stack([smooth_inputs(rng, T, 8, sigma=0.05) for _ in range(nv)])  # [nv,T,8]
G["ens_x0"] = X0
G["ens_U8"] = Uens
out_rk4 = np.zeros((nv, T // 20 + 1, 12))
out_eul = np.zeros_like(out_rk4)
lag_rk4 = np.zeros((nv, 8, 3))
lag_eul = np.zeros((nv, 8, 3))
for i in range(nv):
    r = RefThruster()
    out_rk4[i] = ref_rk4.simulate_physics(X0[i], Uens[i], dt, r)[::20]
    lag_rk4[i] = lag_states(r)
    r = RefThruster()
    out_eul[i] = ref_cmp.simulate_physics(X0[i], Uens[i], dt, r)[::20]
    lag_eul[i] = lag_states(r)
G["ens_thr_rk4_s20"] = out_rk4
G["ens_thr_rk4_lagT"] = lag_rk4
G["ens_thr_euler_s20"] = out_eul
G["ens_thr_euler_lagT"] = lag_eul

Wens = np.stack([smooth_inputs(rng, T, 6, scale=WRENCH_SCALE, sigma=0.05) for _ in range(nv)])
G["ens_W6"] = Wens
o_rk4 = np.zeros((nv, T // 20 + 1, 12))
o_eul = np.zeros_like(o_rk4)
for i in range(nv):
    r = RefWrench12()
    x = X0[i].copy()
    tr = [x.copy()]
    for k in range(T):
        x = rk4_generic(r.dynamics, x, Wens[i, k], dt)
        tr.append(x.copy())
    o_rk4[i] = np.array(tr)[::20]
    o_eul[i] = ref_wc.simulate_physics(X0[i], Wens[i], dt, r)[::20]
G["ens_w12_rk4_s20"] = o_rk4
G["ens_w12_euler_s20"] = o_eul

X0q = state12_to_13(X0)
G["ens_x0_q13"] = X0q
q_rk4 = np.zeros((nv, T // 20 + 1, 13))
q_eul = np.zeros_like(q_rk4)
for i in range(nv):
    r = refq.BlueROV2()
    x = X0q[i].copy()
    tr = [x.copy()]
    for k in range(T):
        # RK4 is not in the reference for the quaternion model; it is composed from the reference's
        # dynamics() and the per-step re-normalisation of train_tank_brov2_wrench_quat.py:262-263.
        x = rk4_generic(r.dynamics, x, Wens[i, k], dt)
        x[3:7] = refq.quat_normalize(x[3:7])
        tr.append(x.copy())
    q_rk4[i] = np.array(tr)[::20]
    q_eul[i] = ref_wq.simulate_physics(X0q[i], Wens[i], dt, r)[::20]
G["ens_q13_rk4_s20"] = q_rk4
G["ens_q13_euler_s20"] = q_eul

# wrench KAT of SURVEY 8(c)
r = RefWrench12()
x = x0.copy()
tau_c = np.array([1, 0.5, -0.2, 0.01, 0.02, 0.03])
for k in range(1000):
    x = rk4_generic(r.dynamics, x, tau_c, dt)
G["kat_w12_rk4_tau"] = tau_c
G["kat_w12_rk4_xend"] = x

# ----------------------------------------------------------------------------------------------
# 5. Monte-Carlo per-vehicle parameters (cfg4): wrench 12 + quaternion 13, RK4, perturbed
#    added-mass and damping attributes; M/Minv REBUILT after editing (trap T5)
# ----------------------------------------------------------------------------------------------
rng = np.random.default_rng(3)
nm, Tm = 6, 200
ADD = ["Xu_dot", "Yv_dot", "Zw_dot", "Kp_dot", "Mq_dot", "Nr_dot"]
DMP = ["Xu", "Yv", "Zw", "Kp", "Mq", "Nr", "Xu_abs", "Yv_abs", "Zw_abs", "Kp_abs", "Mq_abs", "Nr_abs"]
scales = rng.uniform(0.7, 1.3, (nm, 18))
Xm = random_state12(rng, nm)
Wm = np.stack([smooth_inputs(rng, Tm, 6, scale=WRENCH_SCALE, sigma=0.05) for _ in range(nm)])
G["mc_scales"] = scales
G["mc_x0"] = Xm
G["mc_W6"] = Wm
G["mc_attr_names"] = np.array(ADD + DMP)


def perturbed(cls, s):
    r = cls()
    for j, nme in enumerate(ADD + DMP):
        setattr(r, nme, getattr(r, nme) * s[j])
    r.MA = np.diag([-r.Xu_dot, -r.Yv_dot, -r.Zw_dot, -r.Kp_dot, -r.Mq_dot, -r.Nr_dot])
    r.M = r.MRB + r.MA
    r.Minv = np.linalg.inv(r.M)
    return r


mc12 = np.zeros((nm, 12))
mc13 = np.zeros((nm, 13))
Xmq = state12_to_13(Xm)
G["mc_x0_q13"] = Xmq
for i in range(nm):
    r = perturbed(RefWrench12, scales[i])
    x = Xm[i].copy()
    for k in range(Tm):
        x = rk4_generic(r.dynamics, x, Wm[i, k], dt)
    mc12[i] = x
    r = perturbed(refq.BlueROV2, scales[i])
    x = Xmq[i].copy()
    for k in range(Tm):
        x = rk4_generic(r.dynamics, x, Wm[i, k], dt)
        x[3:7] = refq.quat_normalize(x[3:7])
    mc13[i] = x
G["mc_w12_rk4_xT"] = mc12
G["mc_q13_rk4_xT"] = mc13

# ----------------------------------------------------------------------------------------------
# 6. multi-step endpoint RMSE + one-step RMSE evaluators on a synthetic tank-shaped series (cfg5 in
#    miniature).  For the 8-thruster model BOTH semantics are frozen:
#      *_carry : the unmodified reference function (one rov, lag state leaks across windows, trap T3)
#      *_reset : fresh rov per window (what a window-parallel evaluator computes with lag0 = 0)
# ----------------------------------------------------------------------------------------------
rng = np.random.default_rng(4)
Ts = 140
Us = smooth_inputs(rng, Ts, 8, sigma=0.06)
r = RefThruster()
Xs = ref_cmp.simulate_physics(np.array([0.5, -0.3, 1.0, 0.02, -0.01, 0.4, 0, 0, 0, 0, 0, 0.0]), Us, dt, r)[:Ts]
noise = np.concatenate([5e-4 * rng.standard_normal((Ts, 3)), 1e-3 * rng.standard_normal((Ts, 3)),
                        5e-4 * rng.standard_normal((Ts, 3)), 1e-3 * rng.standard_normal((Ts, 3))], axis=1)
Xs = Xs + noise
G["rmse_X12"] = Xs
G["rmse_U8"] = Us
HS = (1, 10, 100)
G["rmse_H"] = np.array(HS)
G["rmse_thr_rk4_carry"] = np.array([ref_rk4.multistep_rmse_endpoint_physics(Xs, Us, H, dt) for H in HS])
G["rmse_thr_euler_carry"] = np.array([ref_cmp.multistep_rmse_endpoint_physics(Xs, Us, H, dt) for H in HS])
G["rmse_thr_onestep_carry"] = np.array(ref_koop.one_step_rmse_physics(Xs, Us, dt))


def rmse_reset(sim, mk, X, U, H):
    se = 0.0
    ns = len(X) - H
    for k in range(ns):
        xe = sim(X[k], U[k:k + H], dt, mk())[-1]
        e = xe - X[k + H]
        se += float(e @ e)
    return np.sqrt(se / (ns * X.shape[1]))


G["rmse_thr_rk4_reset"] = np.array([rmse_reset(ref_rk4.simulate_physics, RefThruster, Xs, Us, H) for H in HS])
G["rmse_thr_euler_reset"] = np.array([rmse_reset(ref_cmp.simulate_physics, RefThruster, Xs, Us, H) for H in HS])
# one-step with reset = every row starts from zero lag
pred = np.stack([Xs[k] + dt * RefThruster().dynamics(Xs[k], Us[k], dt) for k in range(Ts - 1)])
G["rmse_thr_onestep_reset"] = np.array(np.sqrt(np.mean((Xs[1:] - pred) ** 2)))

Ws = smooth_inputs(rng, Ts, 6, scale=WRENCH_SCALE, sigma=0.06)
G["rmse_W6"] = Ws
G["rmse_w12_euler"] = np.array([ref_wc.multistep_rmse_endpoint_physics(Xs, Ws, H, dt) for H in HS])
G["rmse_w12_onestep"] = np.array(ref_wc.one_step_rmse_physics(Xs, Ws, dt))
Xsq = state12_to_13(Xs)
G["rmse_X13"] = Xsq
G["rmse_q13_euler"] = np.array([ref_wq.multistep_rmse_endpoint_physics(Xsq, Ws, H, dt) for H in HS])
G["rmse_q13_onestep"] = np.array(ref_wq.one_step_rmse_physics(Xsq, Ws, dt))
G["rmse_nan_short"] = np.array(ref_wc.multistep_rmse_endpoint_physics(Xs[:5], Ws[:5], 10, dt))

# ----------------------------------------------------------------------------------------------
# 7. simulation data generator (train_sim_brov2_koopmanEDMDc.py:153-197), first 1500 of 240 000 steps.
#    Restated verbatim in structure (legacy np.random.seed(42) stream) around the reference class.
# ----------------------------------------------------------------------------------------------
np.random.seed(42)
dts = 0.05
Ng = 1500
r = RefThruster(dt=dts)
xs = np.zeros(12)
up = np.zeros(8)
st_true = np.zeros((Ng, 12))
st_noisy = np.zeros((Ng, 12))
ins = np.zeros((Ng, 8))
for k in range(Ng):
    u = np.clip(0.98 * up + 0.02 * np.random.randn(8), -1.0, 1.0)
    xs = xs + dts * r.dynamics(xs, u, dts)
    st_true[k] = xs
    ns_ = xs.copy()
    ns_[0:3] += 0.0005 * np.random.randn(3)
    ns_[3:6] += 0.001 * np.random.randn(3)
    ns_[6:9] += 0.0005 * np.random.randn(3)
    ns_[9:12] += 0.001 * np.random.randn(3)
    st_noisy[k] = ns_
    ins[k] = u
    up = u
G["simgen_inputs"] = ins
G["simgen_states_true_s10"] = st_true[::10].copy()
G["simgen_states_noisy_s10"] = st_noisy[::10].copy()

np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **G)
meta = os.path.join(OUT, "reference_vectors.meta.txt")
with open(meta, "w") as f:
    f.write(f"generated by tests/golden/make_golden.py from {REF}\n")
    f.write(f"numpy {np.__version__}; scipy {__import__('scipy').__version__}; torch {torch.__version__}\n")
    for k in sorted(G):
        f.write(f"{k} {tuple(np.shape(G[k]))}\n")
print("wrote", len(G), "arrays;", os.path.getsize(os.path.join(OUT, "reference_vectors.npz")), "bytes")
print("cfg1 const x_end:", np.array2string(G["cfg1_const_traj_s10"][-1], precision=17))
