"""bench.py's driver contract, checked on CPU through the reference arm (the only arm that runs without a GPU)."""

import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--n-bodies", "8192",
                  "--cpu-rows-per-step", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "pairwise_interactions_per_second"
    assert d["unit"] == "interactions/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["n_bodies"] == 8192 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1")
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_reference_arm_config4_and_config3_lines():
    """The other workloads print the same contract line through the CPU arm (bounded samples)."""
    r = run_bench("--impl", "reference", "--workload", "config4", "--steps", "1", "--warmup", "0", "--n-bodies", "4096",
                  "--cpu-rows-per-step", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["config"]["workload"].startswith("config4") and d["config"]["n_bodies"] == 4096 and d["value"] > 0
    r = run_bench("--impl", "reference", "--workload", "config3", "--steps", "1", "--warmup", "0",
                  "--cpu-systems-per-step", "1", "--cpu-inner-steps", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["config"]["workload"].startswith("config3") and d["config"]["systems"] == 4096
    assert d["config"]["interactions_per_step"] == 4096 * 512 * 512 * 1000 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["metric"] == "pairwise_interactions_per_second"


def test_reference_arm_uses_every_host_thread_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use the whole host."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--n-bodies", "4096", "--cpu-rows-per-step", "64"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == os.cpu_count()


def _load_bench():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_clock_warmup_and_sampler_logic():
    """The timing hygiene helpers: extra warm-up steps from a post-warm-up probe, and a sampler that reports only the
    samples taken inside the timed region (or the last one before it for very short regions)."""
    import time

    b = _load_bench()
    assert b.clock_warmup_steps(0.003 * 3, 3) == int(b.CLOCK_WARMUP_SECONDS / 0.003)   # 3 ms steps: ~333 more
    assert b.clock_warmup_steps(0.35 * 3, 3) == int(b.CLOCK_WARMUP_SECONDS / 0.35)      # 350 ms steps: 2 more
    assert b.clock_warmup_steps(1e-9, 3) == 5000                                         # capped

    class FakeNvml:
        NVML_CLOCK_SM = 0

        def nvmlDeviceGetClockInfo(self, h, k):
            return 1965

        def nvmlDeviceGetCurrentClocksEventReasons(self, h):
            return 0x4 | 0x1  # sw_power_cap (kept and noted) + gpu_idle (ignored)

    disabled = b.ClockSampler(0, enabled=False).start()
    with disabled:
        pass
    assert disabled.summary()["sm_mhz"] is None  # ranks other than 0 do not poll

    s = b.ClockSampler(0, period=0.01, enabled=False)
    s.nv, s.h, s.max_mhz = FakeNvml(), None, 1965
    s.start()
    time.sleep(0.03)
    with s:
        time.sleep(0.06)
    out = s.summary()
    assert out["sm_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"] and 2 <= out["samples"] <= 8 and "note" not in out

    s = b.ClockSampler(0, period=0.5, enabled=False)
    s.nv, s.h, s.max_mhz = FakeNvml(), None, 1965
    s.start()
    time.sleep(0.03)
    with s:
        time.sleep(0.01)
    out = s.summary()
    assert out["samples"] == 1 and "before it" in out["note"]
