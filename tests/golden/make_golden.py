#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (bikuta6/nbody-deep-sim) on CPU.

Run in the build container only (the reference is mounted read-only at /root/reference and does not exist on the
GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors for this path, so these files ARE the parity
pin: initial conditions from galaxify.galaxies, then galaxify.simulation.{LeapFrog,Euler}Simulator(device="cpu").
Each file stores the float64 initial conditions, the construction-time accelerations, selected recorded states
(FP32 positions / velocities / accelerations), the reference's own energies, and the library versions used.
"""

import json
import os
import sys

REFERENCE_SRC = os.environ.get("NBODY_REFERENCE_SRC", "/root/reference/src")
sys.dont_write_bytecode = True
sys.path.insert(0, REFERENCE_SRC)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from galaxify import galaxies, simulation  # noqa: E402  (the reference's package)

HERE = os.path.dirname(os.path.abspath(__file__))
S01 = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)  # s01:44-65
S01_SIM = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)

# name, generator, n, seed, integrator, sim kwargs, steps, kept (0-based) steps
CASES = []
for kind in ("disk", "spiral"):
    for n, steps, keep in ((3, 1000, (0, 9, 99, 999)), (25, 1000, (0, 9, 99, 999)), (500, 1000, (0, 9, 99, 999)),
                           (1024, 1000, (0, 9, 99, 999))):
        CASES.append((f"{kind}_n{n}_leapfrog", kind, n, 42, "leapfrog", S01_SIM, steps, keep))
    CASES.append((f"{kind}_n500_euler", kind, 500, 42, "euler", S01_SIM, 100, (0, 9, 99)))
# BaseSimulator's own defaults (G=1, softening=0.1, dt=0.01): genuinely dynamical, short horizon
CASES.append(("disk_n25_defaults_leapfrog", "disk", 25, 7, "leapfrog", dict(g_const=1.0, softening=0.1, dt=0.01), 20,
              (0, 4, 19)))
CASES.append(("spiral_n500_defaults_euler", "spiral", 500, 7, "euler", dict(g_const=1.0, softening=0.1, dt=0.01), 20,
              (0, 4, 19)))
# softening = 0: finite in the reference thanks to fill_diagonal_(0) (simulation.py:85)
CASES.append(("spiral_n25_eps0_leapfrog", "spiral", 25, 3, "leapfrog", dict(g_const=4.5e-6, softening=0.0, dt=1e-4), 10,
              (0, 9)))
# single body: zero acceleration, free drift
CASES.append(("disk_n1_leapfrog", "disk", 1, 1, "leapfrog", S01_SIM, 5, (0, 4)))


def main():
    torch.manual_seed(0)
    meta = dict(torch=torch.__version__, numpy=np.__version__, threads=torch.get_num_threads(),
                reference="bikuta6/nbody-deep-sim src/galaxify (unmodified, device='cpu')")
    for name, kind, n, seed, integ, sim_kw, steps, keep in CASES:
        ic_kw = dict(S01, g_const=sim_kw["g_const"])
        gen = galaxies.generate_disk if kind == "disk" else galaxies.generate_spiral
        pos, vel, mass = gen(n_bodies=n, seed=seed, **ic_kw)
        cls = simulation.LeapFrogSimulator if integ == "leapfrog" else simulation.EulerSimulator
        sim = cls(positions=pos, velocities=vel, masses=mass, calc_energy=True, device="cpu", **sim_kw)
        acc0 = sim.accelerations.clone().numpy()
        u0, k0 = sim.compute_energies()
        states = sim.run(steps)
        out = dict(
            ic_pos=pos, ic_vel=vel, ic_mass=mass, acc0=acc0, u0=np.float64(u0), k0=np.float64(k0),
            steps=np.int64(steps), keep=np.array(keep, dtype=np.int64),
            pos=np.stack([states[s].positions.numpy() for s in keep]),
            vel=np.stack([states[s].velocities.numpy() for s in keep]),
            acc=np.stack([states[s].accelerations.numpy() for s in keep]),
            u=np.array([states[s].u_energy for s in range(steps)], dtype=np.float64),
            k=np.array([states[s].k_energy for s in range(steps)], dtype=np.float64),
            meta=np.array(json.dumps(dict(meta, kind=kind, n=n, seed=seed, integrator=integ, ic=ic_kw, sim=sim_kw))),
        )
        assert [st.step for st in states] == list(range(steps))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
