#!/usr/bin/env python3
"""BASELINE.json configs[2]: batched dataset generation, independent N=512 systems, 1,000 leapfrog steps.

Single process: `--systems` systems on one GPU (the full config is 4,096 / 8 = 512 per GPU). Under torchrun
(`--total-systems 4096`, one rank per GPU) the systems are sharded by index with `batched.shard_systems`, there is
no communication in the data path, and the reported rate is the aggregate over ranks (max time over ranks).
Reports interactions/s (systems x n^2 x steps / device time) with and without trajectory recording, as JSON.
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
    ap.add_argument("--bodies", dest="n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--record-every", type=int, nargs="+", default=[0, 1, 10])
    ap.add_argument("--total-systems", type=int, default=0, help="under torchrun: systems of the whole job")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        total = args.total_systems or args.systems * world
        mine = batched.shard_systems(total, rank, world)
        args.systems = mine.stop - mine.start
    kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)
    base = [galaxies.generate_spiral(n_bodies=args.n, seed=s, **kw) for s in range(8)]
    pos = np.stack([base[s % 8][0] for s in range(args.systems)])
    vel = np.stack([base[s % 8][1] for s in range(args.systems)])
    mass = np.stack([base[s % 8][2] for s in range(args.systems)])
    out = dict(config="BASELINE.json configs[2] (one GPU's share)", systems=args.systems, n=args.n, steps=args.steps,
               runs=[])
    inter = args.systems * args.n * args.n * args.steps
    if world > 1:
        out["config"] = f"BASELINE.json configs[2], {total} systems sharded by index over {world} GPUs"
        t = torch.tensor([float(inter)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        inter = t.item()
        out["systems"] = total
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
            secs = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.barrier()
                dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            best = min(best, secs.item())
        out["runs"].append(dict(record_every=rec, seconds=best, interactions_per_second=inter / best,
                                system_steps_per_second=out["systems"] * args.steps / best,
                                fp32_frac_of_74_45_tflops=20 * inter / best / 74.45e12 / world,
                                trajectory_bytes=0 if traj is None else traj.numel() * 4))
    if rank == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
