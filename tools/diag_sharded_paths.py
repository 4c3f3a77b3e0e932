#!/usr/bin/env python3
"""Times the sharded stepping paths under torchrun: K fused steps in one _advance() call vs K step() calls, with and
without a host synchronisation per step, with and without the L2 flush. Prints ms per step (max over ranks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
from galaxify import galaxies, sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
rank = dist.get_rank()
pos, vel, mass = galaxies.generate_plummer(n_bodies=n, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t[0].item() / K, t[1].item() / K


for overlap in (False, True):
    sim = sharded.ShardedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=1.0, softening=0.01,
                                           dt=1e-3, overlap=overlap)
    sim._advance(4); [sim.step() for _ in range(3)]
    variants = {
        "advance(K)": lambda: sim._advance(K),
        "advance(K)+flush": lambda: sim._advance(K, on_state=lambda s, b: flush.zero_()),
        "K x step()": lambda: [sim.step() for _ in range(K)],
        "K x (flush, step())": lambda: [(flush.zero_(), sim.step()) for _ in range(K)],
        "K x (step(), sync)": lambda: [(sim.step(), torch.cuda.synchronize()) for _ in range(K)],
    }
    for name, fn in variants.items():
        fn()
        dev, wall = timed(fn)
        if rank == 0:
            print(f"n={n} overlap={overlap} {name:22s} device {dev:8.3f} ms/step   wall {wall:8.3f} ms/step", flush=True)
dist.destroy_process_group()
