#!/usr/bin/env python3
"""BASELINE.json configs[2]: batched dataset generation, independent N=512 systems, 1,000 leapfrog steps.

Per GPU the full config is 4,096 / 8 = 512 systems (sharded by system index, no communication). Reports
interactions/s (systems x n^2 x steps / device time) with and without trajectory recording, as JSON.
"""

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from galaxify import batched, galaxies  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", type=int, default=512)
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--record-every", type=int, nargs="+", default=[0, 1, 10])
    args = ap.parse_args()
    kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)
    base = [galaxies.generate_spiral(n_bodies=args.n, seed=s, **kw) for s in range(8)]
    pos = np.stack([base[s % 8][0] for s in range(args.systems)])
    vel = np.stack([base[s % 8][1] for s in range(args.systems)])
    mass = np.stack([base[s % 8][2] for s in range(args.systems)])
    out = dict(config="BASELINE.json configs[2] (one GPU's share)", systems=args.systems, n=args.n, steps=args.steps,
               runs=[])
    inter = args.systems * args.n * args.n * args.steps
    for rec in args.record_every:
        sim = batched.BatchedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6,
                                               softening=0.05, dt=1e-4)
        traj = None
        if rec:
            traj = torch.empty((args.steps // rec, 3, args.systems, args.n, 3), dtype=torch.float32, device="cuda")
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sim._integrate(args.steps, max(rec, 1), traj)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        out["runs"].append(dict(record_every=rec, seconds=best, interactions_per_second=inter / best,
                                system_steps_per_second=args.systems * args.steps / best,
                                fp32_frac_of_74_45_tflops=20 * inter / best / 74.45e12,
                                trajectory_bytes=0 if traj is None else traj.numel() * 4))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
