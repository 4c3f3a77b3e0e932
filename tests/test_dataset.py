"""Dataset CLI and CSV sink: same flags, scene grid, header and row order as the reference's
src/s01-dataset-generation.py. The golden CSV was written by the unmodified reference script on CPU
(`--n-bodies 3 25 --integrator leapfrog --sim-type disk spiral --steps 3 --seed 7 --device cpu`)."""

import csv
import io
import os
from dataclasses import dataclass

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN_DIR
from galaxify import dataset

GOLDEN_CSV = os.path.join(GOLDEN_DIR, "s01_reference_n3_n25_disk_spiral_seed7_steps3.csv")
ARGV = ["--n-bodies", "3", "25", "--integrator", "leapfrog", "--sim-type", "disk", "spiral", "--steps", "3", "--seed", "7"]


@dataclass
class FakeState:
    step: int
    step_time: float
    positions: np.ndarray
    velocities: np.ndarray
    accelerations: np.ndarray
    u_energy: float = None
    k_energy: float = None


def test_header_and_flags_match_reference_script():
    with open(GOLDEN_CSV, newline="") as f:
        header = next(csv.reader(f))
    assert header == dataset.FIELDNAMES
    args = dataset.build_parser().parse_args(ARGV + ["--output", "x.csv"])
    assert (args.dt, args.softening, args.g, args.total_mass, args.radial_scale, args.height_scale,
            args.black_hole_mass, args.n_arms, args.arm_strength, args.device) == (
        1e-4, 0.05, 4.5e-6, 1.0, 3.0, 0.3, 0.01, 2, 0.3, None)
    with pytest.raises(SystemExit):
        dataset.build_parser().parse_args(["--n-bodies", "3", "--output", "x.csv"])  # --integrator is required


def test_scene_grid_order_matches_reference_csv():
    args = dataset.build_parser().parse_args(ARGV + ["--output", "x.csv"])
    grid = dataset.scene_grid(args)
    assert [(c["n_bodies"], c["sim_type"]) for c in grid] == [(3, "disk"), (3, "spiral"), (25, "disk"), (25, "spiral")]
    ref = pd.read_csv(GOLDEN_CSV, float_precision="round_trip")
    per_scene = ref.groupby("scene").agg(n=("mass", lambda s: len(s) // 3), kind=("scene_type", "first"))
    assert [(int(r.n), r.kind) for r in per_scene.itertuples()] == [(c["n_bodies"], c["sim_type"]) for c in grid]
    # and the initial conditions behind those rows are the ones this package generates
    for scene_id, combo in enumerate(grid):
        with np.errstate(divide="ignore", invalid="ignore"):
            pos, vel, mass = dataset.build_scene(combo)
        rows = ref[(ref.scene == scene_id) & (ref.step == 0)]
        np.testing.assert_allclose(rows["mass"].to_numpy(), mass, rtol=1e-15)


def test_sink_writes_what_dictwriter_would(tmp_path):
    rng = np.random.default_rng(0)
    n, steps = 7, 4
    masses = rng.random(n)
    states = [FakeState(step=s, step_time=rng.random() * 1e-3,
                        positions=rng.standard_normal((n, 3)).astype(np.float32) * 10.0 ** rng.integers(-20, 5),
                        velocities=rng.standard_normal((n, 3)).astype(np.float32),
                        accelerations=rng.standard_normal((n, 3)).astype(np.float32) * 1e-7,
                        u_energy=-rng.random(), k_energy=rng.random()) for s in range(steps)]
    path = tmp_path / "ours.csv"
    with dataset.CsvSink(str(path)) as sink:
        assert sink.write_scene(0, "disk", masses, states) == n * steps
        assert sink.write_scene(1, "spiral", masses, states[:2]) == n * 2
    # the reference's writer loop (s01:218-241), restated here as the expectation
    buf = io.StringIO()
    w = csv.DictWriter(buf, fieldnames=dataset.FIELDNAMES)
    w.writeheader()
    for scene_id, kind, sts in ((0, "disk", states), (1, "spiral", states[:2])):
        for st in sts:
            for i in range(n):
                w.writerow(dict(scene=scene_id, scene_type=kind, step=st.step, step_time=st.step_time, mass=masses[i],
                                x=st.positions[i, 0], y=st.positions[i, 1], z=st.positions[i, 2],
                                vx=st.velocities[i, 0], vy=st.velocities[i, 1], vz=st.velocities[i, 2],
                                ax=st.accelerations[i, 0], ay=st.accelerations[i, 1], az=st.accelerations[i, 2],
                                u=st.u_energy, k=st.k_energy))
    want = pd.read_csv(io.StringIO(buf.getvalue()), float_precision="round_trip")
    got = pd.read_csv(path, float_precision="round_trip")
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    for col in dataset.FIELDNAMES:
        if col == "scene_type":
            assert (got[col] == want[col]).all()
        elif col in ("scene", "step"):
            np.testing.assert_array_equal(got[col].to_numpy(), want[col].to_numpy())
        elif col in ("step_time", "mass", "u", "k"):
            np.testing.assert_array_equal(got[col].to_numpy(), want[col].to_numpy())  # float64 round trip
        else:  # float32 columns: both texts round-trip to the same float32
            np.testing.assert_array_equal(got[col].to_numpy().astype(np.float32), want[col].to_numpy().astype(np.float32))


def test_sink_keeps_row_order_when_rendering_blocks_in_parallel(tmp_path):
    n, steps = 400, 300  # 120,000 rows: rendered as several blocks on a thread pool
    rng = np.random.default_rng(1)
    states = [FakeState(step=s, step_time=1e-5, positions=rng.standard_normal((n, 3)).astype(np.float32),
                        velocities=np.full((n, 3), s, np.float32), accelerations=np.zeros((n, 3), np.float32),
                        u_energy=-1.0, k_energy=float(s)) for s in range(steps)]
    path = tmp_path / "big.csv"
    with dataset.CsvSink(str(path)) as sink:
        assert sink.write_scene(3, "disk", np.arange(n, dtype=np.float64), states) == n * steps
    got = pd.read_csv(path, float_precision="round_trip")
    assert len(got) == n * steps
    np.testing.assert_array_equal(got["step"].to_numpy(), np.repeat(np.arange(steps), n))
    np.testing.assert_array_equal(got["mass"].to_numpy(), np.tile(np.arange(n, dtype=np.float64), steps))
    np.testing.assert_array_equal(got["vx"].to_numpy(), np.repeat(np.arange(steps), n).astype(np.float64))
    want_x = np.concatenate([st.positions[:, 0] for st in states])
    np.testing.assert_array_equal(got["x"].to_numpy().astype(np.float32), want_x)
    assert open(path, "rb").read().count(b"\r\n") == n * steps + 1


def test_sink_without_energies(tmp_path):
    st = FakeState(0, 1e-3, np.zeros((2, 3), np.float32), np.zeros((2, 3), np.float32), np.zeros((2, 3), np.float32))
    path = tmp_path / "e.csv"
    with dataset.CsvSink(str(path)) as sink:
        sink.write_scene(0, "disk", np.ones(2), [st])
    got = pd.read_csv(path, float_precision="round_trip")
    assert got["u"].isna().all() and got["k"].isna().all() and len(got) == 2


@pytest.mark.gpu
def test_cli_reproduces_reference_csv(tmp_path):
    out = tmp_path / "ours.csv"
    assert dataset.main(ARGV + ["--output", str(out), "--device", "cuda"]) == 0
    got, want = pd.read_csv(out, float_precision="round_trip"), pd.read_csv(GOLDEN_CSV, float_precision="round_trip")
    assert list(got.columns) == list(want.columns) and len(got) == len(want)
    for col in ("scene", "scene_type", "step"):
        assert (got[col] == want[col]).all()
    np.testing.assert_allclose(got["mass"].to_numpy(), want["mass"].to_numpy(), rtol=1e-15)
    for cols, tol in ((("x", "y", "z"), 1e-6), (("vx", "vy", "vz"), 1e-6), (("ax", "ay", "az"), 1e-5)):
        g, w = got[list(cols)].to_numpy(), want[list(cols)].to_numpy()
        for scene in range(4):
            m = (want["scene"] == scene).to_numpy()
            assert np.abs(g[m] - w[m]).max() <= tol * np.abs(w[m]).max(), (scene, cols)
    for col in ("u", "k"):
        np.testing.assert_allclose(got[col].to_numpy(), want[col].to_numpy(), rtol=1e-5)
    assert (got["step_time"] > 0).all()
