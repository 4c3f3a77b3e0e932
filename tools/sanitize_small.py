#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): tiled force path (both shapes, split-j,
exact-diagonal variant), integrator epilogues, energies, persistent/batched kernels with clusters, host-buffer API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from galaxify import batched, galaxies, host, simulation

kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)
sim_kw = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)
for n, eps in ((3000, 0.05), (5000, 0.0), (40000, 0.05), (700, 0.05), (300, 0.0)):
    pos, vel, mass = galaxies.generate_disk(n_bodies=n, seed=n, **kw)
    for cls in (simulation.LeapFrogSimulator, simulation.EulerSimulator):
        sim = cls(positions=pos, velocities=vel, masses=mass, calc_energy=True, **dict(sim_kw, softening=eps))
        states = sim.run(3)
        sim.step()
        sim.compute_energies()
        assert torch.isfinite(sim.accelerations).all(), (n, eps)
ics = [galaxies.generate_spiral(n_bodies=600, seed=s, **kw) for s in range(3)]
b = batched.BatchedLeapFrogSimulator(positions=np.stack([i[0] for i in ics]), velocities=np.stack([i[1] for i in ics]),
                                     masses=np.stack([i[2] for i in ics]), calc_energy=True, **sim_kw)
b.run(4, record_every=2)
ics = [galaxies.generate_spiral(n_bodies=1500, seed=s, **kw) for s in range(2)]
b = batched.BatchedEulerSimulator(positions=np.stack([i[0] for i in ics]), velocities=np.stack([i[1] for i in ics]),
                                  masses=np.stack([i[2] for i in ics]), **sim_kw)
b.run(2)
pos, vel, mass = galaxies.generate_disk(n_bodies=2500, seed=1, **kw)
acc, _, _ = host.accelerations_host(pos, mass, g_const=4.5e-6, softening=0.05)
p32, v32, m32 = (np.ascontiguousarray(a, dtype=np.float32) for a in (pos, vel, mass))
host.integrate_host("leapfrog", p32, v32, acc, m32, steps=3, record_every=1, calc_energy=True, **sim_kw)
torch.cuda.synchronize()
print("sanitize_small ok")
