#!/usr/bin/env python3
"""Times compute_accelerations and compute_energies (device time) at a few sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import torch
from galaxify import galaxies, simulation

for n in (1024, 16384, 262144, 1 << 20):
    pos, vel, mass = galaxies.generate_plummer(n_bodies=n, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=1)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=1.0, softening=0.01, dt=1e-3,
                                       calc_energy=False)
    for name, fn in (("accel", sim.compute_accelerations), ("energies", sim.compute_energies)):
        fn(); torch.cuda.synchronize()
        reps = 2 if n >= 262144 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"N={n:8d} {name:9s} {ms:10.3f} ms   {n * n / (ms * 1e-3):.3e} pairs(N^2)/s")
