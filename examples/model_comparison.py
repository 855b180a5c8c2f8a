#!/usr/bin/env python3
"""The reference's model-comparison table (training/train_tank_brov2_full_comparison.py:870-1009: Koopman EDMDc, the
Fossen model, the double integrator and PINc scored with the same endpoint-RMSE evaluator at H = 1/10/100, plus the
timing table) on the B200 engine.

    python examples/model_comparison.py [dataset.csv] [--pinc checkpoint.pt] [--rows N]

Without a CSV a 50 Hz "tank-like" series is simulated with the engine's copy of the reference's data generator
(`generate_sim_dataset`) and written next to this script in the reference's wire format first, so that the run goes
through `load_dataset` like the reference's.  PINc is scored when a trained `PINcNet.state_dict()` is given (training
it is outside the accelerated path); without one the PINc column is skipped.
"""
import argparse
import os
import sys
from time import perf_counter

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_dynamics_b200.datasets import generate_sim_dataset, load_dataset, save_dataset  # noqa: E402
from bluerov2_dynamics_b200.evaluators import (estimate_di_gains, multistep_rmse_endpoint_di,  # noqa: E402
                                               multistep_rmse_endpoint_physics)
from bluerov2_dynamics_b200.Koopman.koopmanEDMDc import KoopmanEDMDc  # noqa: E402

TRAIN_SPLIT, N_RBFS, GAMMA, RIDGE = 0.8, 500, 3.0, 1e-1   # the reference script's settings
HORIZONS = (1, 10, 100)


def timed(fn):
    import torch
    torch.cuda.synchronize()
    t0 = perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="?")
    ap.add_argument("--pinc", help="torch checkpoint with a PINcNet state_dict (e.g. the reference's models/pinc_best.pt)")
    ap.add_argument("--rows", type=int, default=45_823, help="length of the simulated series when no CSV is given")
    args = ap.parse_args()

    csv = args.csv
    if csv is None:
        csv = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simulated_dataset_50Hz.csv")
        _, states, inputs = generate_sim_dataset(args.rows, dt=0.02, seed=42)
        save_dataset(csv, states, inputs, 0.02)
        print(f"[i] simulated {args.rows} rows -> {csv}")
    X, U, dt = load_dataset(csv)
    n_train = int(TRAIN_SPLIT * len(X))
    X_train, U_train, X_test, U_test = X[:n_train], U[:n_train], X[n_train:], U[n_train:]
    print(f"[i] Train: {len(X_train)} | Test: {len(X_test)}")

    modelK = KoopmanEDMDc(state_dim=X.shape[1], input_dim=U.shape[1], n_rbfs=N_RBFS, gamma=GAMMA, ridge=RIDGE)
    _, t_fit_koop = timed(lambda: modelK.fit(X_train, U_train))
    (K_lin, K_ang), t_fit_di = timed(lambda: estimate_di_gains(X_train, U_train, dt))

    pinc_model = None
    if args.pinc:
        from bluerov2_dynamics_b200 import pinc as P
        pinc_model = P.PincModel.from_checkpoint(args.pinc)

    rows, times = {}, {}
    rows["Koopman"], times["Koopman"] = zip(*[timed(lambda h=h: modelK.multistep_rmse(X_test, U_test, H=h)) for h in HORIZONS])
    rows["Fossen (BlueROV2)"], times["Fossen (BlueROV2)"] = zip(*[
        timed(lambda h=h: multistep_rmse_endpoint_physics(X_test, U_test, h, dt, integrator="euler")) for h in HORIZONS])
    rows["Double Integrator"], times["Double Integrator"] = zip(*[
        timed(lambda h=h: multistep_rmse_endpoint_di(X_test, U_test, h, dt, K_lin, K_ang, integrator="euler"))
        for h in HORIZONS])
    if pinc_model is not None:
        from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2
        rov_old = BlueROV2(dt=dt)
        rows["PINc (ResDNN)"], times["PINc (ResDNN)"] = zip(*[
            timed(lambda h=h: P.multistep_rmse_endpoint_pinc(X_test, U_test, h, dt, pinc_model, rov_old, None))
            for h in HORIZONS])

    print("\n[metrics] Endpoint RMSE (full 12D state) with identical evaluator:")
    print("  Model                 | 1-step RMSE | 10-step RMSE | 100-step RMSE")
    print("  ----------------------|------------:|-------------:|--------------:")
    for name, r in rows.items():
        print(f"  {name:<21} | {r[0]:11.6f} | {r[1]:12.6f} | {r[2]:13.6f}")
    print("\n[timings] Computation time (seconds):")
    print("  Phase \\ Model         | " + " | ".join(f"{n[:10]:>10}" for n in rows))
    print("  Train/Fit             | " + " | ".join(f"{ {'Koopman': t_fit_koop, 'Double Integrator': t_fit_di}.get(n, 0.0):10.4f}" for n in rows))
    for i, h in enumerate(HORIZONS):
        print(f"  Metrics H={h:<11} | " + " | ".join(f"{times[n][i]:10.4f}" for n in rows))
    return rows


if __name__ == "__main__":
    main()
