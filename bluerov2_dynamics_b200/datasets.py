"""Data formats either side of the hot path: the CSV wire format the reference's training scripts read, and its
simulated-dataset generator.

`load_dataset` mirrors the three variants of the reference (training/train_tank_brov2_rk4.py:84-112 — 12 states +
thruster inputs u1..u8; train_tank_brov2_wrench_comp.py:172-204 — 12 states + wrench Fx..Mz;
train_tank_brov2_wrench_quat.py:180-243 — 13 quaternion states + wrench, with the legacy Euler -> quaternion
conversion): sort by t, drop duplicate stamps, drop rows with non-finite states, zero-fill missing input columns,
dt = median(diff t).  Host-side ETL (pandas), as in the reference.

`generate_sim_dataset` reproduces training/train_sim_brov2_koopmanEDMDc.py:153-197 — smooth random thruster commands,
explicit-Euler rollout of the 8-thruster model, Gaussian sensor noise — with the rollout on the GPU engine.  The
random stream is numpy's legacy generator in the reference's draw order (8 input draws, then 3 + 3 + 3 + 3 noise draws
per step), so seed 42 yields the reference's arrays.
"""
from __future__ import annotations

from pathlib import Path
from typing import Tuple

import numpy as np

STATE_COLS_EULER = ["x", "y", "z", "phi", "theta", "psi", "u", "v", "w", "p", "q", "r"]
STATE_COLS_QUAT = ["x", "y", "z", "qw", "qx", "qy", "qz", "u", "v", "w", "p", "q", "r"]
INPUT_COLS_THRUSTERS = [f"u{i}" for i in range(1, 9)]
INPUT_COLS_WRENCH = ["Fx", "Fy", "Fz", "Mx", "My", "Mz"]


def load_dataset(csv_path, inputs: str = "thrusters", quaternion: bool = False, verbose: bool = True):
    """CSV -> (X [N, 12 or 13], U [N, 8 or 6], dt).  inputs = "thrusters" (u1..u8) | "wrench" (Fx..Mz)."""
    import pandas as pd
    if inputs not in ("thrusters", "wrench"):
        raise ValueError("inputs must be 'thrusters' or 'wrench'")
    if verbose:
        print(f"[i] Loading: {csv_path}")
    df = pd.read_csv(Path(csv_path))
    state_cols = STATE_COLS_QUAT if quaternion else STATE_COLS_EULER
    input_cols = INPUT_COLS_THRUSTERS if inputs == "thrusters" else INPUT_COLS_WRENCH
    has_quat = all(c in df.columns for c in ("qw", "qx", "qy", "qz"))
    if quaternion and not has_quat and all(c in df.columns for c in ("phi", "theta", "psi")):
        if verbose:
            print("[warn] Euler angles detected in dataset; converting to quaternions...")
        half = 0.5 * df[["phi", "theta", "psi"]].to_numpy(dtype=float)
        (c1, c2, c3), (s1, s2, s3) = np.cos(half).T, np.sin(half).T
        q = np.stack([c3 * c2 * c1 + s3 * s2 * s1, c3 * c2 * s1 - s3 * s2 * c1,
                      c3 * s2 * c1 + s3 * c2 * s1, s3 * c2 * c1 - c3 * s2 * s1], axis=1)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
        for j, name in enumerate(("qw", "qx", "qy", "qz")):
            df[name] = q[:, j]
    for c in state_cols:
        if c not in df.columns:
            raise ValueError(f"Missing state column: {c}")
    for c in input_cols:
        if c not in df.columns:
            df[c] = 0.0
    if "t" not in df.columns:
        raise ValueError("CSV must contain a 't' time column.")
    df = df.sort_values("t").drop_duplicates(subset="t")
    df = df.replace([np.inf, -np.inf], np.nan).dropna(subset=state_cols)
    X = np.array(df[state_cols].to_numpy(dtype=float))   # own copy: pandas may hand out a read-only view
    if quaternion:
        X[:, 3:7] /= np.maximum(np.linalg.norm(X[:, 3:7], axis=1, keepdims=True), 1e-12)
    U = df[input_cols].to_numpy(dtype=float)
    t = df["t"].to_numpy(dtype=float)
    dt = np.median(np.diff(t)) if len(t) > 1 else 0.05
    if verbose:
        print(f"[i] Samples: {len(df)} | median dt ≈ {dt:.5f}s (~{1.0 / max(dt, 1e-9):.2f} Hz)")
    return np.ascontiguousarray(X), np.ascontiguousarray(U), float(dt)


def save_dataset(csv_path, X: np.ndarray, U: np.ndarray, dt: float, t0: float = 0.0) -> None:
    """Write a series in the same wire format (column set chosen from the array widths)."""
    import pandas as pd
    X, U = np.asarray(X, float), np.asarray(U, float)
    cols = STATE_COLS_QUAT if X.shape[1] == 13 else STATE_COLS_EULER
    ucols = INPUT_COLS_THRUSTERS if U.shape[1] == 8 else INPUT_COLS_WRENCH
    df = pd.DataFrame(np.hstack([t0 + dt * np.arange(len(X))[:, None], X, U]), columns=["t", *cols, *ucols])
    df.to_csv(Path(csv_path), index=False)


def smooth_random_inputs(n_steps: int, seed: int = 42, n_noise: int = 12) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's command generator u_k = clip(0.98 u_{k-1} + 0.02 N(0,1), -1, 1) and the unit noise draws that
    follow it in the same legacy random stream: (inputs [n_steps, 8], noise [n_steps, n_noise])."""
    rs = np.random.RandomState(seed)
    draws = rs.randn(n_steps, 8 + n_noise)   # one stream, the reference's per-step draw order
    U = np.zeros((n_steps, 8))
    u = np.zeros(8)
    for k in range(n_steps):
        u = np.clip(0.98 * u + 0.02 * draws[k, :8], -1.0, 1.0)
        U[k] = u
    return U, draws[:, 8:]


def generate_sim_dataset(n_steps: int, dt: float = 0.05, seed: int = 42, x0=None, rov=None,
                         noise_std=(0.0005, 0.001, 0.0005, 0.001)):
    """-> (states_true [N,12], states [N,12] noisy, inputs [N,8]); row k is the state AFTER step k, as the reference
    stores it.  noise_std = (position, Euler angles, linear velocity, angular velocity)."""
    from .fossen.BlueROV2 import BlueROV2
    rov = BlueROV2(dt=dt) if rov is None else rov
    U, nz = smooth_random_inputs(n_steps, seed)
    x0 = np.zeros(12) if x0 is None else np.asarray(x0, float)
    eng = rov.engine("f64")
    lag0 = rov._lag_tensor(eng)
    res = eng.rollout(x0.reshape(1, 12), U, dt=dt, integrator="euler", lag0=lag0, stride=1, u_layout="shared")
    rov._store_lag(res.lag, dt)
    states_true = res.traj[:, 0, :].cpu().numpy()
    scale = np.repeat(np.asarray(noise_std, float), 3)
    return states_true, states_true + nz * scale, U
