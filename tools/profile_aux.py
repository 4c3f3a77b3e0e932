#!/usr/bin/env python3
"""One energy evaluation at N=262,144 and one batched run (512 systems x 512 bodies x 100 steps): the two kernels
besides force_kernel, for an ncu capture (-k regex:potential_kernel|batched_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from galaxify import batched, galaxies, simulation

pos, vel, mass = galaxies.generate_plummer(n_bodies=262144, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=1)
sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=1.0, softening=0.01, dt=1e-3,
                                   calc_energy=False)
print("energies", sim.compute_energies())
kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)
base = [galaxies.generate_spiral(n_bodies=512, seed=s, **kw) for s in range(4)]
b = batched.BatchedLeapFrogSimulator(positions=np.stack([base[s % 4][0] for s in range(512)]),
                                     velocities=np.stack([base[s % 4][1] for s in range(512)]),
                                     masses=np.stack([base[s % 4][2] for s in range(512)]), g_const=4.5e-6,
                                     softening=0.05, dt=1e-4)
b._integrate(100, 1, None)
torch.cuda.synchronize()
print("batched ok", float(b.positions.abs().max()))
