import os, sys, ctypes
ROOT = "/root/repo"
for p in (ROOT, os.path.join(ROOT, "nbody-deep-sim_b200")): sys.path.insert(0, p)
import numpy as np, torch
from galaxify import galaxies, simulation, _native
from galaxify.simulation import _ptr
from oracle import c_oracle
n = 262144
pos, vel, mass = galaxies.generate_disk(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01, seed=n)
sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6, softening=0.05, dt=1e-4, calc_energy=False)
a_single = sim.accelerations.cpu().numpy()
L = _native.lib()
dev = torch.device("cuda")
bodies = torch.zeros((n, 4), device=dev); P = torch.tensor(pos, dtype=torch.float32, device=dev); M = torch.tensor(mass, dtype=torch.float32, device=dev)
bodies[:, :3] = P; bodies[:, 3] = M
half = n // 2
a_shard = torch.zeros((n, 3), device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for r in range(2):
    i0 = r * half
    ws = torch.empty(L.nbody_shard_workspace_bytes(half, n, 2), dtype=torch.uint8, device=dev)
    acc = torch.zeros((half, 3), device=dev)
    parts = [(i0, i0 + half), ((1 - r) * half, (1 - r) * half + half)]
    for part, (j0, j1) in enumerate(parts):
        _native.call("nbody_shard_force_f32", 0, _ptr(bodies), None, n, i0, half, j0, j1, part, 2, None, None, _ptr(acc), None,
                     _native.f32(4.5e-6), _native.f32(0.05**2), 0.0, 0.0, 0, _ptr(ws), ws.numel(), st)
    a_shard[i0:i0 + half] = acc
a_shard = a_shard.cpu().numpy()
den = np.linalg.norm(a_single, axis=1)
rel = np.linalg.norm(a_single - a_shard, axis=1) / den
worst = np.argsort(rel)[-8:][::-1]
print("max rel diff single vs shard", rel.max(), "median", np.median(rel))
for i in worst:
    want = c_oracle.accelerations_f64(pos, mass, 4.5e-6, 0.05, int(i), int(i) + 1)[0]
    es = np.linalg.norm(a_single[i] - want) / np.linalg.norm(want); eh = np.linalg.norm(a_shard[i] - want) / np.linalg.norm(want)
    print(i, "rel diff %.2e | single vs f64 %.2e | shard vs f64 %.2e | |a| %.3e  r=%.3f" % (rel[i], es, eh, np.linalg.norm(want), np.linalg.norm(pos[i])))
# overall error distribution vs f64 on a sample
idx = np.r_[0:256, n//2:n//2+256]
want = np.concatenate([c_oracle.accelerations_f64(pos, mass, 4.5e-6, 0.05, 0, 256), c_oracle.accelerations_f64(pos, mass, 4.5e-6, 0.05, n//2, n//2+256)])
for name, a in (("single", a_single), ("shard", a_shard)):
    e = np.linalg.norm(a[idx] - want, axis=1) / np.linalg.norm(want, axis=1)
    print(name, "sample max %.2e median %.2e" % (e.max(), np.median(e)))
