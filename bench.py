#!/usr/bin/env python3
"""Headline benchmark: pairwise interactions/s of one leapfrog step at N = 1,048,576 (BASELINE.json configs[4]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-bodies N]

A "step" is one pass of the hot path over the whole system: all-pairs softened acceleration with the leapfrog
kick/drift fused into the kernel epilogue (N^2 pairwise terms, self pairs included as the reference evaluates them,
SURVEY.md §8d). Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for what every key means.

  value     device-resident throughput: state already in HBM, K steps, CUDA events, max over ranks.
  e2e       the same step through the host-buffer C-ABI call (nbody_integrate_host_f32): pinned host arrays in,
            host arrays out, copies inside the timed region, wall clock.
  roofline  the force kernel against the FP32 FMA peak measured live by the library's FFMA2 probe (there is no FP32
            entry in MEASURED_PEAKS.json), 20 FLOPs per interaction (BASELINE.json).
  cpu_baseline  the CPU oracle port of the reference (same torch operators) on a bounded row sample of the same system.

With --gpus N > 1 (launched by torchrun, one rank per GPU) the i-bodies are sharded and positions all-gathered over
NCCL each step: total work is fixed, so "scaling" is "strong".
"""

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pairwise_interactions_per_second"
UNIT = "interactions/s"
FLOPS_PER_INTERACTION = 20.0  # BASELINE.json north_star
S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)  # s01-dataset-generation.py:44-50 defaults
L2_FLUSH_BYTES = 256 << 20
# dram__bytes_read.sum + dram__bytes_write.sum of one force_kernel launch at N = 1,048,576, from the round-1
# `ncu --set full` capture of this command (profiles/r1_ncu_force_kernel_n1m.txt): 40.89 MB + 33.46 MB.
NCU_TRAFFIC_BYTES_N1M = 74_353_408


def make_system(n):
    from galaxify import galaxies

    return galaxies.generate_disk(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=S01["g_const"],
                                  black_hole_mass=0.01, seed=5)


def workload(n, gpus):
    name = "config5: single disk galaxy (generate_disk, Hernquist-weighted masses + central black hole), "
    name += f"N={n:,}, leapfrog, s01 parameters"
    return {"workload": name, "n_bodies": n, "integrator": "leapfrog", "softening": S01["softening"], "dt": S01["dt"],
            "g_const": S01["g_const"], "interactions_per_step": n * n,
            "sharding": "none" if gpus == 1 else f"i-bodies over {gpus} ranks, NCCL all-gather of positions per step",
            "l2": f"flushed between steps by zeroing a {L2_FLUSH_BYTES >> 20} MiB buffer (inside the timed region)"}


# --------------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.1):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.period = float(os.environ.get("NBODY_BENCH_CLOCK_PERIOD", period))
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            try:  # CUDA_VISIBLE_DEVICES can renumber devices: match NVML to the CUDA device by UUID
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.reasons.update(name for bit, name in self.REASONS.items() if mask & bit and bit != 0x1)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------- CPU arm

def cpu_sample(pos, mass, rows):
    """One bounded sample of the workload on the host: the oracle port of simulation.py:71-89 for `rows` i-bodies
    against all N j-bodies, all torch CPU threads. Returns seconds."""
    from oracle import galaxify_oracle as oracle

    t0 = time.perf_counter()
    oracle.accelerations(pos, mass, S01["g_const"], S01["softening"], rows=slice(0, rows), chunk=32)
    return time.perf_counter() - t0


def cpu_baseline(pos, mass, rows):
    n = len(mass)
    cpu_sample(pos, mass, 32)  # warm torch's thread pool
    if rows <= 0:  # size the sample for ~12 s of CPU work on this host, from a 64-row probe
        probe = cpu_sample(pos, mass, 64)
        rows = int(min(8192, max(256, 64 * 12.0 / probe))) // 32 * 32
    secs = cpu_sample(pos, mass, rows)
    return {"value": rows * n / secs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle/galaxify_oracle.accelerations (torch CPU restatement of simulation.py:71-89) for the "
                      f"first {rows} i-bodies x all {n:,} j-bodies of the same system, {secs:.1f} s; the unmodified "
                      f"reference cannot run this N (its (N,N,3) temporaries need 13 TB)",
            "host_cpus": os.cpu_count()}


def run_reference_arm(args, rank, json_out):
    if rank != 0:
        return
    n = args.n_bodies
    pos, vel, mass = make_system(n)
    rows = args.cpu_rows_per_step
    for _ in range(args.warmup):
        cpu_sample(pos, mass, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(pos, mass, rows)
    secs = time.perf_counter() - t0
    value = rows * n * args.steps / secs
    base = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"each step = oracle port of simulation.py:71-89 for {rows} i-bodies x all {n:,} j-bodies "
                      f"(the full step is {n // rows}x that); unmodified reference cannot allocate this N",
            "host_cpus": os.cpu_count()}
    print(file=json_out, flush=True, *[json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
                      "ms_per_full_step_extrapolated": secs / args.steps * 1e3 * (n / rows),
                      "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(n, args.gpus),
                      "cpu_baseline": base,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})])


# --------------------------------------------------------------------------------------------------- GPU arm

def fp32_peak_tflops(device_index):
    import ctypes

    from galaxify import _native

    out = {}
    for packed, key in ((1, "ffma2"), (0, "ffma")):
        tf = ctypes.c_double()
        _native.call("nbody_probe_fp32_peak", device_index, packed, ctypes.byref(tf))
        out[key] = tf.value
    return out


def run_single(args, dev):
    from galaxify import _native, host, simulation

    n = args.n_bodies
    pos, vel, mass = make_system(n)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    kernel_ms = []

    def step(record):
        flush.zero_()
        ms = np.zeros(1, dtype=np.float32)
        sim._integrate(1, 1, None, None, ms)  # prep + fused force/leapfrog kernel; ms = the force kernel alone
        if record:
            kernel_ms.append(float(ms[0]))

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    with ClockSampler(dev) as clocks:
        e0.record()
        for _ in range(args.steps):
            step(True)
        e1.record()
        torch.cuda.synchronize()
    launches = _native.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)

    # end to end: host buffers through the C ABI, copies inside the timed region
    hp = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().numpy()
          for a in (sim.positions.cpu().numpy(), sim.velocities.cpu().numpy(), sim.accelerations.cpu().numpy(),
                    sim.masses.cpu().numpy())]
    r = None
    for _ in range(max(1, args.warmup)):
        r = host.integrate_host("leapfrog", *hp, steps=1, device=dev, **S01)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = host.integrate_host("leapfrog", *hp, steps=1, device=dev, **S01)
    e2e_secs = time.perf_counter() - t0
    assert np.isfinite(hp[0]).all()
    return dict(n=n, total_ms=total_ms, kernel_ms=kernel_ms, launches=launches, clocks=clocks.summary(),
                e2e_value=n * n * args.steps / e2e_secs, h2d=r["h2d_bytes"], d2h=r["d2h_bytes"], pos=pos, mass=mass,
                interactions_per_kernel=n * n)


def run_sharded(args, dev, rank, world):
    import torch.distributed as dist

    from galaxify import _native, sharded

    n = args.n_bodies
    pos, vel, mass = make_system(n)
    sim = sharded.ShardedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, **S01)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    # warm-up = the timed code path (both ping-pong body arrays go through a collective at least once)
    sim._advance(args.warmup, on_state=lambda s, bodies: flush.zero_())
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    with ClockSampler(dev) as clocks:
        e0.record()
        # K consecutive steps of one run: each step's epilogue opens the next, L2 flushed between steps
        sim._advance(args.steps, on_state=lambda s, bodies: flush.zero_())
        e1.record()
        torch.cuda.synchronize()
    launches = _native.launch_count() - launches0
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    # end to end: this rank's slice of the state comes from pinned host memory and goes back every step
    state = [sim.positions, sim.velocities, sim.accelerations, sim.masses]
    pinned = [x.cpu().pin_memory() for x in state]
    out = [torch.empty_like(p) for p in pinned[:3]]
    out = [o.pin_memory() for o in out]

    def e2e_step():
        for d, h in zip(state, pinned):
            d.copy_(h, non_blocking=True)
        sim.step()
        for h, d in zip(out, state[:3]):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([sum(p.numel() * 4 for p in pinned)], device="cuda", dtype=torch.float64)
    d2h = torch.tensor([sum(o.numel() * 4 for o in out)], device="cuda", dtype=torch.float64)
    dist.all_reduce(h2d)
    dist.all_reduce(d2h)
    return dict(n=n, total_ms=total_ms, kernel_ms=[], launches=launches, clocks=clocks.summary(),
                e2e_value=n * n * args.steps / float(t.item()), h2d=int(h2d.item()), d2h=int(d2h.item()), pos=pos,
                mass=mass, interactions_per_kernel=None, launches_per_step=sim.launches_per_step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n-bodies", type=int, default=1 << 20)
    ap.add_argument("--cpu-rows", type=int, default=0,
                    help="i-bodies of the cpu_baseline sample (0 = sized for ~12 s on this host)")
    ap.add_argument("--cpu-rows-per-step", type=int, default=256, help="i-bodies per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args, rank, json_out)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        res = run_sharded(args, local_rank, rank, world)
    else:
        res = run_single(args, local_rank)

    if rank == 0:
        n = res["n"]
        secs = res["total_ms"] * 1e-3
        value = n * n * args.steps / secs
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["total_ms"] / args.steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload(n, world), "steps_per_second": args.steps / secs,
                "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"]},
                "gpu_launches": res["launches"], "clocks": res["clocks"]}
        peaks = fp32_peak_tflops(local_rank)
        nominal = 148 * 128 * 2 * 1.965e9 / 1e12
        if res["kernel_ms"]:
            k_ms = statistics.mean(res["kernel_ms"])
            achieved = FLOPS_PER_INTERACTION * res["interactions_per_kernel"] / (k_ms * 1e-3) / 1e12
            kernel = "force_kernel<2,16,1,1024,unroll 32,fold 32> (1 launch per step)"
        else:  # sharded: several force launches per step; report the whole-step rate per GPU
            k_ms = res["total_ms"] / args.steps
            achieved = FLOPS_PER_INTERACTION * n * n / world / (k_ms * 1e-3) / 1e12
            kernel = f"force launches of one rank ({res['launches_per_step']} per step), per-GPU share of the step"
        line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peaks["ffma2"], "unit": "TFLOP/s",
                            "frac": achieved / peaks["ffma2"],
                            "traffic": NCU_TRAFFIC_BYTES_N1M if (world == 1 and n == 1 << 20) else None,
                            "kernel": kernel,
                            "kernel_ms": k_ms,
                            "peak_source": "measured live: nbody_probe_fp32_peak (register-resident FFMA2 chains, "
                                           "best of 6); MEASURED_PEAKS.json has no FP32 entry",
                            "peak_ffma_scalar": peaks["ffma"], "nominal_peak": nominal,
                            "frac_of_nominal": achieved / nominal,
                            "flops_per_interaction": FLOPS_PER_INTERACTION,
                            "issue_bound_frac": 20.0 / 24.0}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(res["pos"], res["mass"], args.cpu_rows)
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
