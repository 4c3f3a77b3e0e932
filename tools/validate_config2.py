#!/usr/bin/env python3
"""BASELINE.json configs[1]: single disk galaxy N=16,384, 10,000 leapfrog steps on one B200 — accuracy and drift.

The reference's CPU path would need ~14 h for this, so the comparison run is the oracle restatement of
the reference executed with device="cuda" (the same ATen operators the reference's own `device="cuda"` path runs,
simulation.py:46-51, 80-88). Writes one JSON document (default profiles/r1_config2_validation.json):
per-particle acceleration error at steps {0, 1, 10, 100, 1000, 10000}, trajectory deviation, energy drift under the
reference's energy definition (simulation.py:91-115) for both engines, momentum drift for both.
"""

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from galaxify import galaxies, simulation  # noqa: E402
from oracle import galaxify_oracle as oracle  # noqa: E402


def rel_rows(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_config2_validation.json"))
    args = ap.parse_args()
    kw = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)
    pos, vel, mass = galaxies.generate_disk(n_bodies=args.n, total_mass=1.0, radial_scale=3.0, height_scale=0.3,
                                            g_const=4.5e-6, black_hole_mass=0.01, seed=42)
    marks = sorted({0, 1, 10, 100, 1000, args.steps} & set(range(args.steps + 1)))  # 1-based step counts; 0 = initial
    m64 = mass.astype(np.float64)[:, None]

    # ---- ours: one call, every step recorded on the device, energies every step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=True, **kw)
    acc0 = sim.accelerations.cpu().numpy()
    u0, k0 = sim.compute_energies()
    states = sim.run(args.steps)
    ours_secs = time.perf_counter() - t0
    ours_gpu_secs = float(sum(s.step_time for s in states))

    # ---- oracle on the GPU: the reference's operators, step by step
    t0 = time.perf_counter()
    st = oracle.State(pos, vel, mass, device="cuda", chunk=4096, **kw)
    ref = {0: dict(acc=st.acc.cpu().numpy())}
    ref_u, ref_k = [], []
    e = st.energies()
    ref_e0 = e
    for s in range(1, args.steps + 1):
        st.leapfrog_step()
        if s in marks:
            ref[s] = dict(pos=st.pos.cpu().numpy(), vel=st.vel.cpu().numpy(), acc=st.acc.cpu().numpy())
        if s in marks or s % 500 == 0:
            u, k = st.energies()
            ref_u.append((s, u))
            ref_k.append((s, k))
    torch.cuda.synchronize()
    ref_secs = time.perf_counter() - t0

    out = dict(config="BASELINE.json configs[1]", n=args.n, steps=args.steps, params=kw,
               ours_wall_seconds=ours_secs, ours_device_seconds_steps_and_energies=ours_gpu_secs,
               reference_ops_on_gpu_wall_seconds=ref_secs, marks={})
    out["marks"]["0"] = dict(acc_rel_max=float(rel_rows(acc0, ref[0]["acc"]).max()),
                             acc_rel_median=float(np.median(rel_rows(acc0, ref[0]["acc"]))))
    for s in marks:
        if s == 0:
            continue
        mine, want = states[s - 1], ref[s]
        err = rel_rows(mine.accelerations.numpy(), want["acc"])
        out["marks"][str(s)] = dict(
            acc_rel_max=float(err.max()), acc_rel_median=float(np.median(err)),
            pos_dev_over_max=float(np.abs(mine.positions.numpy() - want["pos"]).max() / np.abs(want["pos"]).max()),
            vel_dev_over_max=float(np.abs(mine.velocities.numpy() - want["vel"]).max() / np.abs(want["vel"]).max()))
    e_ours = np.array([s.u_energy + s.k_energy for s in states])
    e_ref = {s: u + dict(ref_k)[s] for s, u in ref_u}
    last = args.steps
    out["energy"] = dict(
        ours_initial=u0 + k0, reference_initial=ref_e0[0] + ref_e0[1],
        ours_drift_first_to_last=float(abs(e_ours[-1] - e_ours[0]) / abs(e_ours[0])),
        reference_drift_first_to_last=float(abs(e_ref[last] - e_ref[1]) / abs(e_ref[1])) if 1 in e_ref else None,
        ours_vs_reference_at_last=float(abs(e_ours[-1] - e_ref[last]) / abs(e_ref[last])),
        ours_max_excursion=float(np.abs(e_ours - e_ours[0]).max() / abs(e_ours[0])))
    p_ours = [(m64 * s.velocities.numpy().astype(np.float64)).sum(0) for s in (states[0], states[-1])]
    p_ref = [(m64 * ref[s]["vel"].astype(np.float64)).sum(0) for s in (1, last)] if 1 in ref else None
    out["momentum"] = dict(ours_drift=float(np.linalg.norm(p_ours[1] - p_ours[0])),
                           reference_drift=float(np.linalg.norm(p_ref[1] - p_ref[0])) if p_ref else None,
                           magnitude=float(np.linalg.norm(p_ours[0])))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
