#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's configurations. Default: the headline, configs[4].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config5|config4|config3]

  config5 (default)  single disk galaxy, N = 1,048,576, leapfrog: pairwise interactions/s of one step. With --gpus N > 1
                     (torchrun, one rank per GPU) the bodies are sharded and every unordered pair is evaluated once
                     over all ranks (all-gather of positions + reduce-scatter of forces over NCCL): "scaling": "strong".
  config4            two-galaxy merger (galaxies.merge of two generate_disk galaxies of 131,072 bodies), N = 262,144,
                     same step, 1/2/4/8 GPUs.
  config3            batched dataset generation: 4,096 independent 512-body systems x 1,000 leapfrog steps, sharded by
                     system index over the ranks, no communication. A "step" of the bench is one such 1,000-step run.

A "step" (config5/4) is one pass of the hot path over the whole system: all-pairs softened acceleration with the
leapfrog kick/drift fused into the kernel epilogue (N^2 pairwise terms, self pairs included as the reference
evaluates them). Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for what every key means.

  value         device-resident throughput: state already in HBM, K steps, CUDA events, max over ranks.
  e2e           the same work through the public host-facing call with pinned HOST buffers, copies inside the timed
                region (1 GPU: the C-ABI nbody_integrate_host_f32; N GPUs: ShardedLeapFrogSimulator.step with the
                rank's slice copied in and out every step; config3: BatchedLeapFrogSimulator with the trajectory
                copied back at the stated stride).
  roofline      the dominant kernel against the FP32 FMA peak measured live by the library's FFMA2 probe (there is no
                FP32 entry in MEASURED_PEAKS.json), 20 FLOPs per interaction (BASELINE.json).
  parity_check  after the timed region: sampled rows of the final accelerations against the FP64 C oracle at the
                final positions (<= 1e-5 per particle, or the run exits non-zero). The oracle is the checker only.
  cpu_baseline  the CPU oracle port of the reference (same torch operators) on a bounded sample of the same workload.
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
GAL = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=S01["g_const"], black_hole_mass=0.01)
L2_FLUSH_BYTES = 256 << 20
PARITY_RTOL = 1e-5
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the `ncu --set full` capture of
# this command kept under profiles/ (null where no capture of that workload exists).
NCU_TRAFFIC_BYTES = {"config5": None, "config4": None, "config3": None}
_traffic_file = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
if os.path.exists(_traffic_file):
    NCU_TRAFFIC_BYTES.update(json.load(open(_traffic_file)))

CONFIG3 = dict(systems=4096, bodies=512, inner_steps=1000, record_every=100)


def make_system(workload, n):
    from galaxify import galaxies

    if workload == "config4":
        a = galaxies.generate_disk(n_bodies=n // 2, seed=1, **GAL)
        b = galaxies.generate_disk(n_bodies=n - n // 2, seed=2, offset=(12.0, 3.0, 1.0), initial_vel=(-2e-4, 0.0, 0.0),
                                   angle=(0.4, 0.0, 0.3), **GAL)
        return galaxies.merge(a, b)
    return galaxies.generate_disk(n_bodies=n, seed=5, **GAL)


def make_batch(systems, bodies, first=0):
    """`systems` spiral galaxies of `bodies` bodies (seeds first, first+1, ...: 64 distinct ICs, cycled)."""
    from galaxify import galaxies

    base = [galaxies.generate_spiral(n_bodies=bodies, seed=s, **GAL) for s in range(64)]
    idx = [(first + s) % 64 for s in range(systems)]
    return tuple(np.stack([base[i][k] for i in idx]) for k in range(3))


def workload_config(workload, n, gpus):
    if workload == "config3":
        c = CONFIG3
        return {"workload": f"config3: batched dataset generation, {c['systems']:,} independent spiral-galaxy systems of "
                            f"{c['bodies']} bodies, {c['inner_steps']:,} leapfrog steps per bench step, s01 parameters",
                "systems": c["systems"], "n_bodies": c["bodies"], "inner_steps": c["inner_steps"], "integrator": "leapfrog",
                "softening": S01["softening"], "dt": S01["dt"], "g_const": S01["g_const"],
                "interactions_per_step": c["systems"] * c["bodies"] ** 2 * c["inner_steps"],
                "sharding": "none" if gpus == 1 else f"systems split by index over {gpus} ranks, no communication",
                "l2": "working set lives in shared memory / registers for the whole launch; nothing to flush"}
    if workload == "config4":
        name = f"config4: two-galaxy merger (galaxies.merge of two generate_disk galaxies), N={n:,}, leapfrog, s01 parameters"
    else:
        name = "config5: single disk galaxy (generate_disk, Hernquist-weighted masses + central black hole), "
        name += f"N={n:,}, leapfrog, s01 parameters"
    sharding = "none"
    if gpus > 1:
        sharding = (f"bodies over {gpus} ranks; per step NCCL all-gather of positions (16 B/body) and, on the pair path, "
                    f"reduce-scatter of FP64 force sums (24 B/body)")
    return {"workload": name, "n_bodies": n, "integrator": "leapfrog", "softening": S01["softening"], "dt": S01["dt"],
            "g_const": S01["g_const"], "interactions_per_step": n * n, "sharding": sharding,
            "l2": f"flushed between steps by zeroing a {L2_FLUSH_BYTES >> 20} MiB buffer (inside the timed region)",
            "clock_warmup": f"untimed steps are repeated until the GPU has been loaded for {CLOCK_WARMUP_SECONDS} s"}


# --------------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML from a background thread.

    start() is called BEFORE the warm-up steps, so that the expensive first calls of the NVML queries (tens of ms, during
    which this process's kernel launches stall) are paid outside the timed region; `with sampler:` then marks the timed
    region, and only samples taken inside it are reported. One sample per 0.2 s, on rank 0 only: on an 8-GPU box every
    poll costs the polled process milliseconds of launch latency (measured: 8 ranks polling at 10 Hz added 1.1 ms to every
    step, rank 0 alone at 10 Hz still 0.3 ms)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.2, enabled=True):
        self.samples, self.max_mhz, self._stop = [], None, threading.Event()  # samples: (time, mhz, reason mask)
        self.period = float(os.environ.get("NBODY_BENCH_CLOCK_PERIOD", period))
        self.t0 = self.t1 = None
        self.nv, self.t = None, None
        if not enabled:
            self.err = "sampling is done by rank 0 only"
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES can renumber devices: match NVML to the CUDA device by UUID
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.err = repr(e)

    def start(self):
        if self.nv and self.t is None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        return self

    def _loop(self):
        while not self._stop.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self.start()
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()
        self._stop.set()
        if self.t is not None:
            self.t.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        inside = [x for x in self.samples if self.t0 <= x[0] <= self.t1]
        note = None
        if not inside:  # region shorter than the sampling period: the last sample before it (GPU under warm-up load)
            inside = [x for x in self.samples if x[0] <= self.t0][-1:]
            note = "no sample fell inside the timed region; the last one before it (under warm-up load) is reported"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no samples"}
        reasons = set()
        for _, _, mask in inside:
            reasons.update(name for bit, name in self.REASONS.items() if mask & bit and bit != 0x1)
        out = {"sm_mhz": statistics.median(x[1] for x in inside), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
               "samples": len(inside), "period_s": self.period}
        if note:
            out["note"] = note
        return out


# --------------------------------------------------------------------------------------------------- CPU arms

def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and takes the whole host."""
    n = os.cpu_count() or 1
    try:
        torch.set_num_threads(n)
    except Exception:
        pass
    return torch.get_num_threads()


def cpu_sample(pos, mass, rows):
    """One bounded sample of config5/config4 on the host: the oracle port of simulation.py:71-89 for `rows` i-bodies
    against all N j-bodies, all torch CPU threads. Returns seconds."""
    from oracle import galaxify_oracle as oracle

    t0 = time.perf_counter()
    oracle.accelerations(pos, mass, S01["g_const"], S01["softening"], rows=slice(0, rows), chunk=32)
    return time.perf_counter() - t0


def cpu_sample_batched(pos, vel, mass, systems, steps):
    """One bounded sample of config3 on the host: `steps` leapfrog steps of the first `systems` systems, one after
    the other as s01-dataset-generation.py:130-214 runs its scenes, through the oracle port. Returns seconds."""
    from oracle import galaxify_oracle as oracle

    t0 = time.perf_counter()
    for s in range(systems):
        oracle.run(pos[s], vel[s], mass[s], integrator="leapfrog", steps=steps, keep=(), **S01)
    return time.perf_counter() - t0


def cpu_baseline(workload, data, rows):
    cores = use_all_host_threads()
    if workload == "config3":
        pos, vel, mass = data
        n = pos.shape[1]
        cpu_sample_batched(pos, vel, mass, 1, 2)
        probe = cpu_sample_batched(pos, vel, mass, 1, 20) / 21  # seconds per force evaluation of one system
        budget = 12.0 / probe  # force evaluations in ~12 s
        steps = int(min(CONFIG3["inner_steps"], max(20, budget / 2)))
        systems = int(max(1, min(len(pos), budget / (steps + 1))))
        secs = cpu_sample_batched(pos, vel, mass, systems, steps)
        # force evaluations: one per step plus the one in the constructor (simulation.py:69)
        return {"value": systems * (steps + 1) * n * n / secs, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"oracle/galaxify_oracle.run (torch CPU restatement of simulation.py:71-89,153-170), "
                          f"{systems} of the {CONFIG3['systems']:,} systems x {steps} of the {CONFIG3['inner_steps']:,} "
                          f"steps, one system at a time as s01 runs its scenes, {secs:.1f} s",
                "host_cpus": os.cpu_count()}
    pos, mass = data
    n = len(mass)
    cpu_sample(pos, mass, 32)  # warm torch's thread pool
    if rows <= 0:  # size the sample for ~12 s of CPU work on this host, from a 64-row probe
        probe = cpu_sample(pos, mass, 64)
        rows = int(min(8192, max(256, 64 * 12.0 / probe))) // 32 * 32
    secs = cpu_sample(pos, mass, rows)
    return {"value": rows * n / secs, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/galaxify_oracle.accelerations (torch CPU restatement of simulation.py:71-89) for the "
                      f"first {rows} i-bodies x all {n:,} j-bodies of the same system, {secs:.1f} s; the unmodified "
                      f"reference cannot run this N (its (N,N,3) temporaries need {12 * n * n / 1e12:.1f} TB)",
            "host_cpus": os.cpu_count()}


def reference_config1_block(dev_index):
    """BASELINE.json configs[0] / BASELINE.md 3: N = 1,024, leapfrog, s01 parameters, on the host CPU through the oracle
    port (operator for operator the reference's loop; the unmodified reference is not on the GPU box and bench.py may
    not read /root/reference), with and without energies, on a bounded number of steps; and the same run through this
    engine's public API on the GPU for the end-to-end ratio at that size."""
    from galaxify import galaxies, simulation
    from oracle import galaxify_oracle as oracle

    n, steps_cpu, steps_gpu = 1024, 100, 1000
    pos, vel, mass = galaxies.generate_disk(n_bodies=n, seed=42, **GAL)
    out = {"n_bodies": n, "cpu_steps_timed": steps_cpu, "cpu_kind": "port", "cores": torch.get_num_threads()}
    oracle.run(pos, vel, mass, integrator="leapfrog", steps=3, keep=(), **S01)
    for key, energy in (("cpu_ms_per_step", False), ("cpu_ms_per_step_with_energy", True)):
        t0 = time.perf_counter()
        oracle.run(pos, vel, mass, integrator="leapfrog", steps=steps_cpu, calc_energy=energy, **S01)
        out[key] = (time.perf_counter() - t0) / steps_cpu * 1e3
    for key, energy in (("gpu_ms_per_step", False), ("gpu_ms_per_step_with_energy", True)):
        sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=energy, **S01)
        sim.run(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        states = sim.run(steps_gpu)  # wall clock: kernels + trajectory D2H + SimulationState list, as s01 consumes it
        out[key] = (time.perf_counter() - t0) / steps_gpu * 1e3
        assert len(states) == steps_gpu
    out["note"] = ("reference-style run(): every state recorded and returned on the host; CPU = oracle port, "
                   f"{steps_cpu} steps; GPU = this engine's LeapFrogSimulator.run({steps_gpu}), wall clock")
    return out


def reference_gpu_block(sizes=(1024, 4096, 16384)):
    """BASELINE.md 3 'reference GPU' line: the reference's own operators (simulation.py:80-88, restated in
    oracle/galaxify_oracle.accelerations) run unfused by torch on this GPU, synchronised, next to this engine."""
    from galaxify import galaxies, simulation
    from oracle import galaxify_oracle as oracle

    rows = []
    for n in sizes:
        pos, vel, mass = galaxies.generate_disk(n_bodies=n, seed=7, **GAL)
        p = torch.tensor(pos, dtype=torch.float32, device="cuda")
        m = torch.tensor(mass, dtype=torch.float32, device="cuda")
        oracle.accelerations(p, m, S01["g_const"], S01["softening"], chunk=n, device="cuda")
        torch.cuda.synchronize()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.accelerations(p, m, S01["g_const"], S01["softening"], chunk=n, device="cuda")
        torch.cuda.synchronize()
        ref_ms = (time.perf_counter() - t0) / reps * 1e3
        del p, m
        torch.cuda.empty_cache()
        sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.compute_accelerations()
        e0.record()
        for _ in range(20):
            sim.compute_accelerations()
        e1.record()
        torch.cuda.synchronize()
        ours_ms = e0.elapsed_time(e1) / 20
        rows.append({"n_bodies": n, "reference_ops_cuda_ms": ref_ms,
                     "reference_ops_cuda_interactions_per_s": n * n / ref_ms * 1e3, "ours_ms": ours_ms,
                     "ours_interactions_per_s": n * n / ours_ms * 1e3})
    return {"sizes": rows, "note": "one compute_accelerations; reference operators = torch ATen on the same B200, "
                                   "unchunked (N,N,3) temporaries, wall clock with synchronize; ours = CUDA events"}


def run_reference_arm(args, rank, json_out):
    if rank != 0:
        return
    cores = use_all_host_threads()
    wl = args.workload
    if wl == "config3":
        c = CONFIG3
        pos, vel, mass = make_batch(4, c["bodies"])
        n, systems, inner = c["bodies"], args.cpu_systems_per_step, args.cpu_inner_steps
        for _ in range(args.warmup):
            cpu_sample_batched(pos, vel, mass, 1, 2)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_sample_batched(pos, vel, mass, systems, inner)
        secs = time.perf_counter() - t0
        value = systems * (inner + 1) * n * n * args.steps / secs
        sample = (f"each step = oracle port of simulation.py:71-89,153-170 for {systems} systems x {inner} leapfrog steps "
                  f"(the full step is {c['systems']:,} systems x {c['inner_steps']:,} steps)")
        full_factor = c["systems"] * c["inner_steps"] / (systems * (inner + 1))
        cfg = workload_config(wl, n, args.gpus)
    else:
        n = args.n_bodies
        pos, vel, mass = make_system(wl, n)
        rows = args.cpu_rows_per_step
        for _ in range(args.warmup):
            cpu_sample(pos, mass, rows)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_sample(pos, mass, rows)
        secs = time.perf_counter() - t0
        value = rows * n * args.steps / secs
        sample = (f"each step = oracle port of simulation.py:71-89 for {rows} i-bodies x all {n:,} j-bodies "
                  f"(the full step is {n // rows}x that); unmodified reference cannot allocate this N")
        full_factor = n / rows
        cfg = workload_config(wl, n, args.gpus)
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "host_cpus": os.cpu_count()}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
                      "ms_per_full_step_extrapolated": secs / args.steps * 1e3 * full_factor,
                      "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
          file=json_out, flush=True)


# --------------------------------------------------------------------------------------------------- GPU arm

CLOCK_WARMUP_SECONDS = 1.0


def clock_warmup_steps(probe_seconds, probe_steps):
    """Extra untimed steps so that the GPU has been under load for about CLOCK_WARMUP_SECONDS when the timed region
    starts. The W warm-up steps of the contract are too short for that when a step takes a few milliseconds (config4 on
    8 GPUs: 3 steps = 10 ms after seconds of host-side initial-condition generation with the GPU idle), and the first
    ~50 ms after idle run below the boost clock: the same 20 steps measured 4.8 ms/step right after 3 warm-up steps and
    3.1 ms/step once warm (profiles/r2_diag_sharded_pair.log). `probe_seconds` is the wall time of `probe_steps` steps
    run AFTER the W warm-up steps (whose own wall time includes one-off costs such as NCCL communicator set-up)."""
    per_step = max(probe_seconds / max(probe_steps, 1), 1e-5)
    return int(min(5000, max(0, CLOCK_WARMUP_SECONDS / per_step)))


CLOCK_PROBE_STEPS = 3


def fp32_peak_tflops(device_index):
    import ctypes

    from galaxify import _native

    out = {}
    for packed, key in ((1, "ffma2"), (0, "ffma")):
        tf = ctypes.c_double()
        _native.call("nbody_probe_fp32_peak", device_index, packed, ctypes.byref(tf))
        out[key] = tf.value
    return out


def parity_check(pos, mass, acc, n_rows=256):
    """Sampled rows of `acc` against the FP64 C oracle evaluated at `pos` (test infrastructure used as the checker)."""
    from oracle import c_oracle

    n = len(mass)
    rows = np.unique(np.concatenate([np.linspace(0, n - 1, n_rows - 2).astype(np.int64), [0, n - 1]]))
    want = np.stack([c_oracle.accelerations_f64(pos, mass, S01["g_const"], S01["softening"], int(i), int(i) + 1)[0]
                     for i in rows])
    got = np.asarray(acc, dtype=np.float64)[rows]
    err = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    return {"max_rel": float(err.max()), "median_rel": float(np.median(err)), "rows": int(len(rows)), "rtol": PARITY_RTOL,
            "finite": bool(np.isfinite(got).all()), "ok": bool(np.isfinite(got).all() and err.max() <= PARITY_RTOL),
            "against": "oracle/nbody_oracle.c accelerations (FP64) at the final positions"}


def run_single(args, dev):
    from galaxify import _native, host, simulation

    n = args.n_bodies
    pos, vel, mass = make_system(args.workload, n)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    kernel_ms = []

    def step(record):
        flush.zero_()
        ms = np.zeros(1, dtype=np.float32)
        sim._integrate(1, 1, None, None, ms)  # prep + force (+ finish) launches; ms = the step's force launches alone
        if record:
            kernel_ms.append(float(ms[0]))

    clocks = ClockSampler(dev).start()
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(CLOCK_PROBE_STEPS):
        step(False)
    torch.cuda.synchronize()
    for _ in range(clock_warmup_steps(time.perf_counter() - t0, CLOCK_PROBE_STEPS)):
        step(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    with clocks:
        e0.record()
        for _ in range(args.steps):
            step(True)
        e1.record()
        torch.cuda.synchronize()
    launches = _native.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    parity = parity_check(sim.positions.cpu().numpy(), mass, sim.accelerations.cpu().numpy())

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
    pair = n >= _native.lib().nbody_pair_min_bodies()
    kernel = ("pair_kernel<2,12,1> (384 threads x 4 i-bodies, every unordered pair once) + finish_kernel, 2 launches per step"
              if pair else "force_kernel (directed, 1 launch per step)")
    return dict(n=n, total_ms=total_ms, kernel_ms=kernel_ms, launches=launches, clocks=clocks.summary(),
                e2e_value=n * n * args.steps / e2e_secs, h2d=r["h2d_bytes"], d2h=r["d2h_bytes"], cpu_data=(pos, mass),
                interactions_per_step=n * n, kernel=kernel, parity=parity)


def run_sharded(args, dev, rank, world):
    import torch.distributed as dist

    from galaxify import _native, sharded

    n = args.n_bodies
    pos, vel, mass = make_system(args.workload, n)
    sim = sharded.ShardedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, **S01)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    # warm-up = the timed code path (both ping-pong body arrays go through a collective at least once)
    clocks = ClockSampler(dev, enabled=(rank == 0)).start()
    sim._advance(args.warmup, on_state=lambda s, bodies: flush.zero_())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim._advance(CLOCK_PROBE_STEPS, on_state=lambda s, bodies: flush.zero_())
    torch.cuda.synchronize()
    extra = torch.tensor([clock_warmup_steps(time.perf_counter() - t0, CLOCK_PROBE_STEPS)], device="cuda")
    dist.all_reduce(extra, op=dist.ReduceOp.MAX)  # every rank must run the same number of collective steps
    sim._advance(int(extra.item()), on_state=lambda s, bodies: flush.zero_())
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    with clocks:
        e0.record()
        # K consecutive steps of one run: each step's epilogue opens the next, L2 flushed between steps
        sim._advance(args.steps, on_state=lambda s, bodies: flush.zero_())
        e1.record()
        torch.cuda.synchronize()
    launches = _native.launch_count() - launches0
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    # parity of what was just timed: the gathered final state against the FP64 oracle (rank 0 evaluates the oracle)
    gp, gv, ga = sim.gather_state()
    parity = parity_check(gp.numpy(), mass, ga.numpy()) if rank == 0 else None

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
    path = "pair path: own-slot triangle || all-gather, cross rectangles, FP64 reduce-scatter, finish" if sim.pair else \
        "directed path: own slice || all-gather, rest"
    return dict(n=n, total_ms=total_ms, kernel_ms=[], launches=launches, clocks=clocks.summary(),
                e2e_value=n * n * args.steps / float(t.item()), h2d=int(h2d.item()), d2h=int(d2h.item()),
                cpu_data=(pos, mass), interactions_per_step=n * n, parity=parity,
                kernel=f"force launches of one rank ({sim.launches_per_step} per step; {path}), per-GPU share of the step")


def run_batched(args, dev, rank, world):
    """config3: this rank's share of the systems, all inner steps in ONE persistent launch per bench step."""
    from galaxify import _native, batched

    c = CONFIG3
    mine = batched.shard_systems(c["systems"], rank, world)
    n_sys = mine.stop - mine.start
    pos, vel, mass = make_batch(n_sys, c["bodies"], first=mine.start)
    n, inner, rec = c["bodies"], c["inner_steps"], c["record_every"]
    sim = batched.BatchedLeapFrogSimulator(positions=pos, velocities=vel, masses=mass, **S01)
    kernel_ms = []

    def step(record):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sim._integrate(inner, 1, None)
        b.record()
        if record:
            kernel_ms.append((a, b))

    clocks = ClockSampler(dev, enabled=(rank == 0)).start()
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(False)
    torch.cuda.synchronize()
    for _ in range(clock_warmup_steps(time.perf_counter() - t0, 1)):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    with clocks:
        e0.record()
        for _ in range(args.steps):
            step(True)
        e1.record()
        torch.cuda.synchronize()
    launches = _native.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    kernel_ms = [a.elapsed_time(b) for a, b in kernel_ms]

    # parity of one system of this rank against the CPU oracle port over a short horizon
    parity = None
    if rank == 0:
        from oracle import galaxify_oracle as oracle

        chk = batched.BatchedLeapFrogSimulator(positions=pos[:2], velocities=vel[:2], masses=mass[:2], **S01)
        st = chk.run(20)[-1]
        ref, _ = oracle.run(pos[1], vel[1], mass[1], integrator="leapfrog", steps=20, keep=(19,), **S01)
        got, want = st.accelerations[1].numpy().astype(np.float64), ref[19]["acc"].astype(np.float64)
        err = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        dx = np.abs(st.positions[1].numpy() - ref[19]["pos"]).max() / np.abs(ref[19]["pos"]).max()
        parity = {"max_rel": float(err.max()), "median_rel": float(np.median(err)), "rows": int(len(err)), "rtol": PARITY_RTOL,
                  "pos_max_rel": float(dx), "finite": bool(np.isfinite(got).all()),
                  "ok": bool(err.max() <= PARITY_RTOL and dx <= 1e-6),
                  "against": "oracle/galaxify_oracle.run (CPU port of the reference), system 1 after 20 steps"}

    # end to end: initial conditions from pinned host arrays, the recorded trajectory back to pinned host memory
    hp = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory() for a in (pos, vel, mass)]
    slots = inner // rec
    host_traj = torch.empty((slots, 3, n_sys, n, 3), dtype=torch.float32, pin_memory=True)
    dev_traj = torch.empty((slots, 3, n_sys, n, 3), dtype=torch.float32, device="cuda")

    def e2e_step():
        sim.positions.copy_(hp[0], non_blocking=True)
        sim.velocities.copy_(hp[1], non_blocking=True)
        sim.masses.copy_(hp[2], non_blocking=True)
        sim.accelerations = sim.compute_accelerations()  # what the constructor does (simulation.py:69)
        sim._integrate(inner, rec, dev_traj)
        host_traj.copy_(dev_traj, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    secs = time.perf_counter() - t0
    h2d, d2h = sum(h.numel() * 4 for h in hp), host_traj.numel() * 4
    if world > 1:
        t = torch.tensor([total_ms, secs, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        total_ms, secs, h2d, d2h = float(tmax[0]), float(tmax[1]), int(t[2]), int(t[3])
    inter = c["systems"] * n * n * inner
    return dict(n=n, total_ms=total_ms, kernel_ms=kernel_ms, launches=launches, clocks=clocks.summary(),
                e2e_value=inter * args.steps / secs, h2d=h2d, d2h=d2h, cpu_data=(pos[:32], vel[:32], mass[:32]),
                interactions_per_step=inter, parity=parity, kernel_interactions=n_sys * n * n * inner,
                kernel=f"batched_kernel (one cluster per system, {inner} steps per launch), {n_sys} systems on this GPU; "
                       f"e2e records every {rec}th state ({slots} slots) and copies it to the host")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["config5", "config4", "config3"], default="config5")
    ap.add_argument("--n-bodies", type=int, default=0, help="override N of config5/config4")
    ap.add_argument("--cpu-rows", type=int, default=0,
                    help="i-bodies of the cpu_baseline sample (0 = sized for ~12 s on this host)")
    ap.add_argument("--cpu-rows-per-step", type=int, default=256, help="i-bodies per step of --impl reference")
    ap.add_argument("--cpu-systems-per-step", type=int, default=2, help="config3: systems per step of --impl reference")
    ap.add_argument("--cpu-inner-steps", type=int, default=50, help="config3: leapfrog steps per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-baselines", action="store_true",
                    help="skip the config1 CPU/GPU block and the reference-operators-on-GPU block (BASELINE.md 3), "
                         "which add ~15 s to a single-GPU run")
    args = ap.parse_args()
    args.n_bodies_overridden = args.n_bodies > 0
    if args.n_bodies <= 0:
        args.n_bodies = 262144 if args.workload == "config4" else 1 << 20
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
    if args.workload == "config3":
        res = run_batched(args, local_rank, rank, world)
    elif world > 1:
        res = run_sharded(args, local_rank, rank, world)
    else:
        res = run_single(args, local_rank)

    rc = 0
    if rank == 0:
        n = res["n"]
        secs = res["total_ms"] * 1e-3
        per_step = res["interactions_per_step"]
        value = per_step * args.steps / secs
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["total_ms"] / args.steps, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.workload, n, world), "steps_per_second": args.steps / secs,
                "e2e": {"value": res["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"]},
                "gpu_launches": res["launches"], "clocks": res["clocks"], "parity_check": res["parity"]}
        peaks = fp32_peak_tflops(local_rank)
        nominal = 148 * 128 * 2 * 1.965e9 / 1e12
        if res["kernel_ms"]:
            k_ms = statistics.mean(res["kernel_ms"])
            achieved = FLOPS_PER_INTERACTION * res.get("kernel_interactions", per_step) / (k_ms * 1e-3) / 1e12
        else:  # sharded: several force launches per step; report the whole-step rate per GPU
            k_ms = res["total_ms"] / args.steps
            achieved = FLOPS_PER_INTERACTION * per_step / world / (k_ms * 1e-3) / 1e12
        line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peaks["ffma2"], "unit": "TFLOP/s",
                            "frac": achieved / peaks["ffma2"],
                            "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if (world == 1 and not args.n_bodies_overridden) else None,
                            "kernel": res["kernel"], "kernel_ms": k_ms,
                            "peak_source": "measured live: nbody_probe_fp32_peak (register-resident FFMA2 chains, "
                                           "best of 6); MEASURED_PEAKS.json has no FP32 entry",
                            "peak_ffma_scalar": peaks["ffma"], "nominal_peak": nominal,
                            "frac_of_nominal": achieved / nominal,
                            "flops_per_interaction": FLOPS_PER_INTERACTION,
                            "note": "20 FLOPs per DIRECTED interaction (N^2 per step); the pair kernel evaluates each "
                                    "unordered pair once (16 packed-FP32 instructions per 4 directed interactions), so "
                                    "its issue-bound ceiling is above 1.0 of this convention"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, res["cpu_data"], args.cpu_rows)
        if world == 1 and not args.no_extra_baselines and not args.no_cpu_baseline:
            line["reference_config1"] = reference_config1_block(local_rank)
            line["reference_gpu"] = reference_gpu_block()
        print(json.dumps(line), file=json_out, flush=True)
        if res["parity"] is not None and not res["parity"]["ok"]:
            print(f"bench.py: parity check FAILED: {res['parity']}", file=sys.stderr)
            rc = 3
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
