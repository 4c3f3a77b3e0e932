#!/usr/bin/env python3
"""Diagnostic (not part of the product): error / condition number of the pair path and of the directed kernel on the
ill-conditioned bodies of the config4 merger, against the FP64 oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from conftest import rel_rows  # noqa: E402
from test_gpu_pair import _directed_accelerations  # noqa: E402
from test_gpu_shard_emulated import _system  # noqa: E402
from galaxify import simulation  # noqa: E402
from oracle import c_oracle  # noqa: E402

S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pos, vel, mass = _system(n, True)
sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
pair = sim.accelerations.cpu().numpy()
directed = _directed_accelerations(pos, mass)
diff = rel_rows(pair, directed)
rows = np.unique(np.concatenate([np.argsort(diff)[-40:], np.arange(0, n, n // 200)]))
want, kappa = c_oracle.accelerations_cond_f64(pos, mass, S01["g_const"], S01["softening"], rows)
ep, ed = rel_rows(pair[rows], want), rel_rows(directed[rows], want)
order = np.argsort(kappa)[::-1][:25]
print("   row      kappa   pair_err  directed_err  pair/kappa  directed/kappa")
for k in order:
    print(f"{rows[k]:7d} {kappa[k]:9.1f} {ep[k]:10.2e} {ed[k]:10.2e} {ep[k] / kappa[k]:10.2e} {ed[k] / kappa[k]:10.2e}")
print("max err/kappa  pair %.2e  directed %.2e ; median err pair %.2e directed %.2e" % (
    (ep / kappa).max(), (ed / kappa).max(), np.median(ep), np.median(ed)))
