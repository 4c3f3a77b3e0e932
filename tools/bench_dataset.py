#!/usr/bin/env python3
"""Dataset-generation sizes of the reference's experiments (gnn_experiment.py:28-48: spiral, n in 3..500, 1,000 leapfrog
steps, energies on) against the only numbers the reference publishes for this path: the mean leapfrog step time read
off figures/stepwise_time.png. Prints JSON; also times the CSV sink."""

import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from galaxify import dataset, galaxies, simulation  # noqa: E402
from oracle import galaxify_oracle as oracle  # noqa: E402

PUBLISHED_MS = {3: 0.06, 25: 0.10, 50: 0.20, 100: 0.62, 250: 1.32, 500: 2.95}  # BASELINE.md §1, CPU, read by eye
KW = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)


def main():
    steps = 1000
    rows = []
    for n in (3, 25, 50, 100, 250, 500, 1024):
        pos, vel, mass = galaxies.generate_spiral(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3,
                                                  g_const=4.5e-6, black_hole_mass=0.01, seed=1)
        sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=True, **KW)
        sim.run(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        states = sim.run(steps)
        wall = time.perf_counter() - t0
        dev_ms = sum(s.step_time for s in states) * 1e3 / steps
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            with dataset.CsvSink(os.path.join(d, "x.csv")) as sink:
                n_rows = sink.write_scene(0, "spiral", mass, states)
            csv_s = time.perf_counter() - t0
        # the reference's operators on this host's CPU (oracle port), 20 steps
        st = oracle.State(pos, vel, mass, **KW)
        t0 = time.perf_counter()
        for _ in range(20):
            st.leapfrog_step()
            st.energies()
        cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
        rows.append(dict(n=n, device_ms_per_step_incl_energies=dev_ms, wall_ms_per_step_run=wall / steps * 1e3,
                         published_reference_cpu_ms_per_step=PUBLISHED_MS.get(n),
                         oracle_port_cpu_ms_per_step_incl_energies=cpu_ms, csv_rows=n_rows,
                         csv_rows_per_second=n_rows / csv_s))
    print(json.dumps(dict(steps=steps, cpu_threads=torch.get_num_threads(), rows=rows), indent=1))


if __name__ == "__main__":
    main()
