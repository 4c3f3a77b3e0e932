"""Dataset generation: the scene grid and CSV sink of the reference's CLI, with a vectorised writer.

Mirrors (reference, read-only) src/s01-dataset-generation.py: same command-line flags (:12-91), same Cartesian
product of the list-valued flags in the same order (:93-103), same CSV header and row order — scene, then step, then
particle (:107-127, :218-241) — so datautils.py:23-34 groups the file identically. What changes is the sink: the
reference formats one dict per particle per step through csv.DictWriter; here each scene's recorded states become
columnar arrays written in one call (pyarrow's CSV writer, float columns at shortest round-trip precision, which is
also what the reference's str(np.float32) gives), and the simulation itself runs inside one C call.

    python -m galaxify.dataset --integrator leapfrog --n-bodies 3 25 50 --steps 1000 --sim-type spiral --seed 1 \
        --output out.csv
"""

from __future__ import annotations

import argparse
import itertools
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

FIELDNAMES = ["scene", "scene_type", "step", "step_time", "mass", "x", "y", "z", "vx", "vy", "vz", "ax", "ay", "az",
              "u", "k"]


def build_parser() -> argparse.ArgumentParser:
    """The flags of s01-dataset-generation.py:12-91, same names, types, choices and defaults."""
    p = argparse.ArgumentParser(description="Generación de dataset de simulaciones de galaxias")
    p.add_argument("--n-bodies", type=int, nargs="+", required=True)
    p.add_argument("--integrator", type=str, default="leapfrog", choices=["leapfrog", "euler"], required=True)
    p.add_argument("--output", type=str, required=True)
    p.add_argument("--sim-type", type=str, nargs="+", choices=["disk", "spiral"], default=["disk"])
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--dt", type=float, default=0.0001)
    p.add_argument("--softening", type=float, default=0.05)
    p.add_argument("--g", type=float, default=4.5e-6)
    p.add_argument("--total-mass", type=float, default=1.0)
    p.add_argument("--radial-scale", type=float, default=3.0)
    p.add_argument("--height-scale", type=float, default=0.3)
    p.add_argument("--black-hole-mass", type=float, default=0.01)
    p.add_argument("--n-arms", type=int, default=2)
    p.add_argument("--pitch-angle", type=float, default=-np.pi / 6)
    p.add_argument("--arm-strength", type=float, default=0.3)
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--device", type=str, choices=["cuda", "cpu"], default=None)
    return p


def scene_grid(args: argparse.Namespace) -> list[dict]:
    """Every flag except output/device is a product axis, in argparse order (s01:93-103)."""
    params = {}
    for key, value in vars(args).items():
        if key in ("output", "device"):
            continue
        params[key] = value if isinstance(value, list) else [value]
    keys = list(params)
    return [dict(zip(keys, combo)) for combo in itertools.product(*(params[k] for k in keys))]


def build_scene(combo: dict):
    """Initial conditions of one scene (s01:159-185)."""
    from . import galaxies

    common = dict(n_bodies=combo["n_bodies"], total_mass=combo["total_mass"], radial_scale=combo["radial_scale"],
                  height_scale=combo["height_scale"], g_const=combo["g"], black_hole_mass=combo["black_hole_mass"],
                  seed=combo["seed"])
    if combo["sim_type"] == "disk":
        return galaxies.generate_disk(**common)
    if combo["sim_type"] == "spiral":
        return galaxies.generate_spiral(n_arms=combo["n_arms"], pitch_angle=combo["pitch_angle"],
                                        arm_strength=combo["arm_strength"], **common)
    raise ValueError(f"Tipo de simulación desconocido: {combo['sim_type']}")


def scene_columns(scene_id: int, scene_type: str, masses, states) -> dict:
    """Columnar form of the rows s01:218-241 writes for one scene: state-major, particle-minor."""
    n = len(masses)
    s = len(states)
    cols = {
        "scene": np.full(s * n, scene_id, dtype=np.int64),
        "scene_type": np.full(s * n, scene_type, dtype=object),
        "step": np.repeat(np.array([st.step for st in states], dtype=np.int64), n),
        "step_time": np.repeat(np.array([st.step_time for st in states], dtype=np.float64), n),
        "mass": np.tile(np.asarray(masses, dtype=np.float64), s),
    }
    for names, attr in ((("x", "y", "z"), "positions"), (("vx", "vy", "vz"), "velocities"),
                        (("ax", "ay", "az"), "accelerations")):
        block = np.concatenate([np.asarray(getattr(st, attr).cpu().numpy() if hasattr(getattr(st, attr), "cpu")
                                           else getattr(st, attr), dtype=np.float32) for st in states], axis=0)
        for c, name in enumerate(names):
            cols[name] = np.ascontiguousarray(block[:, c])
    for name, attr in (("u", "u_energy"), ("k", "k_energy")):
        vals = [getattr(st, attr) for st in states]
        if any(v is None for v in vals):
            cols[name] = np.full(s * n, "", dtype=object)  # csv.DictWriter writes None as an empty field
        else:
            cols[name] = np.repeat(np.array(vals, dtype=np.float64), n)
    return cols


class CsvSink:
    """Appends scenes to one CSV file with the reference's header (s01:107-127)."""

    def __init__(self, path: str):
        import pyarrow as pa
        import pyarrow.csv as pacsv

        self._pa, self._pacsv = pa, pacsv
        self._threads = max(1, min(8, os.cpu_count() or 1))
        self._f = open(path, "wb")
        self._f.write((",".join(FIELDNAMES) + "\r\n").encode())  # csv.DictWriter's default line terminator

    def write_scene(self, scene_id: int, scene_type: str, masses, states) -> int:
        if not states:
            return 0
        cols = scene_columns(scene_id, scene_type, masses, states)
        arrays = []
        for name in FIELDNAMES:
            c = cols[name]
            arrays.append(self._pa.array(c.tolist() if c.dtype == object else c))
        table = self._pa.Table.from_arrays(arrays, names=FIELDNAMES)
        opts = self._pacsv.WriteOptions(include_header=False, quoting_style="none")

        def render(bounds):
            lo, hi = bounds
            buf = self._pa.BufferOutputStream()
            self._pacsv.write_csv(table.slice(lo, hi - lo), buf, write_options=opts)
            return buf.getvalue().to_pybytes().replace(b"\n", b"\r\n")  # csv.DictWriter terminates lines with CRLF

        # number formatting dominates: render row blocks on a few threads (pyarrow releases the GIL), keep their order
        rows = table.num_rows
        blocks = max(1, min(self._threads, rows // 50_000))
        bounds = [(rows * b // blocks, rows * (b + 1) // blocks) for b in range(blocks)]
        if blocks == 1:
            self._f.write(render(bounds[0]))
        else:
            with ThreadPoolExecutor(blocks) as pool:
                for chunk in pool.map(render, bounds):
                    self._f.write(chunk)
        return rows

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def generate(args: argparse.Namespace, log=print) -> int:
    """Runs every scene of the grid and writes args.output. Returns the number of rows written."""
    from . import simulation

    combos = scene_grid(args)
    log(f"Generando {len(combos)} escenarios...")
    log(f"Creando dataset {args.output}")
    rows = 0
    with CsvSink(args.output) as sink:
        for scene_id, combo in enumerate(combos):
            pos, vel, masses = build_scene(combo)
            cls = simulation.EulerSimulator if args.integrator == "euler" else simulation.LeapFrogSimulator
            sim = cls(positions=pos, velocities=vel, masses=masses, g_const=combo["g"], softening=combo["softening"],
                      dt=combo["dt"], calc_energy=True, device=args.device)
            states = sim.run(combo["steps"])
            rows += sink.write_scene(scene_id, combo["sim_type"], masses, states)
            log(f"Escenario {scene_id + 1}/{len(combos)}: n={combo['n_bodies']} {combo['sim_type']} ✅")
    return rows


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    generate(args)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
