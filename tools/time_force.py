#!/usr/bin/env python3
"""Quick timing of compute_accelerations at a few sizes (CUDA events, best of 3). Not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from galaxify import simulation  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1 << 20]
for n in sizes:
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((n, 3)).astype(np.float32) * 3
    mass = (rng.uniform(0.5, 1.5, n) / n).astype(np.float32)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=np.zeros((n, 3), np.float32), masses=mass, g_const=1.0,
                                       softening=0.05, dt=1e-4, calc_energy=False)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.compute_accelerations()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"n {n:8d}  {best:9.3f} ms  {n * n / best * 1e-9:8.3f} T int/s  {20 * n * n / best * 1e3 / 74.45e12 * 100:5.1f}% of 74.45 TF", flush=True)
