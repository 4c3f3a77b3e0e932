#!/usr/bin/env python3
"""Diagnostic (not part of the product): where a sharded pair-path step spends its time, stage by stage, under torchrun.
   python -m torch.distributed.run --nproc-per-node P tools/diag_sharded_pair.py [n] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from galaxify import _native, galaxies, sharded  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
overlap = None if len(sys.argv) <= 3 else (True if int(sys.argv[3]) else None)
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)
pos, vel, mass = galaxies.generate_disk(n_bodies=n, seed=5, **kw)
sim = sharded.ShardedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6, softening=0.05, dt=1e-4,
                                       overlap=overlap)
integ = _native.INTEGRATOR_LEAPFROG
names = ["gather", "pair0", "pair1", "reduce", "finish"]
acc = {k: 0.0 for k in names}
tot = 0.0


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


sim._advance(3)
torch.cuda.synchronize()
dist.barrier()
cur = 0
sim._prepare(integ, sim._bodies[cur])
for s in range(steps):
    b, bn = sim._bodies[cur], sim._bodies[cur ^ 1]
    t0 = ev()
    sim._gather(b).wait()
    t1 = ev()
    sim._pair_force(0, b)
    t2 = ev()
    if sim._pair_split:
        sim._pair_force(1, b)
    t3 = ev()
    sim._pair_reduce()
    t4 = ev()
    sim._pair_finish(integ, b, bn, 1)
    t5 = ev()
    torch.cuda.synchronize()
    for k, (a, c) in zip(names, ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))):
        acc[k] += a.elapsed_time(c)
    tot += t0.elapsed_time(t5)
    cur ^= 1
    dist.barrier()
# free-running loops (no host sync between steps), with and without the bench's L2 flush
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
free = {}
for name, cb in (("free", None), ("free+flush", lambda s_, b_: flush.zero_())):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a = ev()
    sim._advance(20, on_state=cb)
    c = ev()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(c) / 20], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    free[name] = float(t.item())
line = f"rank {rank} n {n} split {sim._pair_split}: " + "  ".join(f"{k} {acc[k] / steps:8.3f} ms" for k in names)
print(line + f"   total {tot / steps:8.3f} ms   " + "  ".join(f"{k} {v:.3f} ms/step" for k, v in free.items()), flush=True)
dist.barrier()
dist.destroy_process_group()
