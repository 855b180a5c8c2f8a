import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_dynamics_b200.fossen.BlueROV2 import BlueROV2
from bluerov2_dynamics_b200.fossen.BlueROV2_thrust import BlueROV2 as W
from bluerov2_dynamics_b200.evaluators import simulate_physics
rov = BlueROV2(dt=0.02); x = np.zeros(12); x[2] = 5; u = np.full(8, 0.2)
for _ in range(20): rov.dynamics(x, u, 0.02)
t = time.perf_counter()
for _ in range(500): rov.dynamics(x, u, 0.02)
print("thruster dynamics() per call: %.1f us" % ((time.perf_counter() - t) / 500 * 1e6))
w = W(); tau = np.ones(6)
for _ in range(20): w.dynamics(x, tau)
t = time.perf_counter()
for _ in range(500): w.dynamics(x, tau)
print("wrench dynamics() per call: %.1f us" % ((time.perf_counter() - t) / 500 * 1e6))
U = np.tile(u, (1000, 1))
simulate_physics(x, U, 0.02, rov)
t = time.perf_counter(); simulate_physics(x, U, 0.02, rov); print("simulate_physics 1000 RK4 steps, one vehicle: %.2f ms" % ((time.perf_counter() - t) * 1e3))
