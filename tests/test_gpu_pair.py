"""The Newton's-third-law ("pair") kernel path, taken for N >= 32,768 (csrc/pair.cuh): ragged sizes around its tile
sizes (I-tiles of 768, J-tiles of 128, 32-body systolic blocks), both integrators, against the FP64 oracle
(<= 1e-5 per particle, north_star) and against the directed kernel of csrc/force.cuh on the same inputs."""

import ctypes

import numpy as np
import pytest
import torch

from conftest import assert_accelerations_agree, rel_rows

pytestmark = pytest.mark.gpu

S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)
KW = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01)


def _rows(n):
    """Sample of i-bodies: both ends, tile boundaries of the pair kernel, and a stride through the rest."""
    edges = [0, 1, 31, 32, 127, 128, 129, 383, 384, 767, 768, 769, 1535, 1536, 1537, n - 1537, n - 769, n - 385, n - 129,
             n - 33, n - 32, n - 2, n - 1]
    return np.unique(np.clip(np.concatenate([edges, np.arange(7, n, n // 160)]), 0, n - 1))


def _oracle_rows(pos, mass, rows):
    from oracle import c_oracle

    return np.stack([c_oracle.accelerations_f64(pos, mass, S01["g_const"], S01["softening"], int(i), int(i) + 1)[0]
                     for i in rows])


def _directed_accelerations(pos, mass):
    """The same system through the sharded building block, which always runs the directed kernel (force.cuh)."""
    from galaxify import _native
    from galaxify.simulation import _ptr

    n = len(mass)
    dev = torch.device("cuda")
    p, m = torch.tensor(pos, dtype=torch.float32, device=dev), torch.tensor(mass, dtype=torch.float32, device=dev)
    bodies = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    acc = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    ws = torch.empty(_native.lib().nbody_shard_workspace_bytes(n, n, 1), dtype=torch.uint8, device=dev)
    _native.call("nbody_shard_prepare_f32", 0, _ptr(p), None, None, _ptr(m), None, _ptr(bodies), 0, n, 0.0, 0.0, None)
    _native.call("nbody_shard_force_f32", 0, _ptr(bodies), None, n, 0, n, 0, n, 0, 0, 0, 1, None, None, _ptr(acc), None,
                 _native.f32(S01["g_const"]), _native.f32(S01["softening"] ** 2), 0.0, 0.0, 0, _ptr(ws), ws.numel(), None)
    return acc.cpu().numpy()


@pytest.mark.parametrize("n", [32768, 32769, 33000, 40001, 65536, 65537, 66000, 70001, 98304 + 383, 131072, 200003])
def test_pair_path_accelerations_vs_oracle_and_directed_kernel(n):
    from galaxify import galaxies, simulation

    pos, vel, mass = galaxies.generate_disk(n_bodies=n, seed=n % 1000, **KW)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    acc = sim.accelerations.cpu().numpy()
    assert np.isfinite(acc).all()
    rows = _rows(n)
    assert rel_rows(acc[rows], _oracle_rows(pos, mass, rows)).max() <= 1e-5
    # every body against the directed kernel: two FP32 evaluations, each within ~1e-6 of exact
    directed = _directed_accelerations(pos, mass)
    assert_accelerations_agree(acc, directed, pos, mass, S01["g_const"], S01["softening"])


@pytest.mark.parametrize("integrator", ["leapfrog", "euler"])
def test_pair_path_trajectory_matches_directed_path(integrator, monkeypatch):
    """A few steps with recording and energies on the pair path vs the same run forced through force.cuh by a
    softening below the pair path's limit... not available: instead compare with the emulated single-rank sharded
    stepping (directed kernel) through the C ABI."""
    from galaxify import _native, galaxies, simulation
    from galaxify.simulation import _ptr

    n, steps = 70001, 3
    a = galaxies.generate_disk(n_bodies=n // 2, seed=5, **KW)
    b = galaxies.generate_disk(n_bodies=n - n // 2, seed=6, offset=(9.0, 2.0, 1.0), initial_vel=(-0.002, 0.0, 0.0),
                               angle=(0.3, 0.1, 0.0), **KW)
    pos, vel, mass = galaxies.merge(a, b)
    cls = simulation.LeapFrogSimulator if integrator == "leapfrog" else simulation.EulerSimulator
    sim = cls(positions=pos, velocities=vel, masses=mass, calc_energy=True, **S01)
    states = sim.run(steps)
    assert [s.step for s in states] == list(range(steps))

    # directed-kernel stepping of the same system: one "rank" owning everything
    code = _native.INTEGRATOR_LEAPFROG if integrator == "leapfrog" else _native.INTEGRATOR_EULER
    dev = torch.device("cuda")
    st = {k: torch.tensor(v, dtype=torch.float32, device=dev) for k, v in (("pos", pos), ("vel", vel), ("mass", mass))}
    st["acc"] = torch.tensor(_directed_accelerations(pos, mass), device=dev)
    st["vhalf"] = torch.zeros_like(st["pos"])
    bodies = [torch.zeros((n, 4), dtype=torch.float32, device=dev) for _ in range(2)]
    ws = torch.empty(_native.lib().nbody_shard_workspace_bytes(n, n, 1), dtype=torch.uint8, device=dev)
    sc = dict(g=_native.f32(S01["g_const"]), eps2=_native.f32(S01["softening"] ** 2), dt=_native.f32(S01["dt"]),
              half=_native.f32(0.5 * S01["dt"]))
    _native.call("nbody_shard_prepare_f32", code, _ptr(st["pos"]), _ptr(st["vel"]), _ptr(st["acc"]), _ptr(st["mass"]),
                 _ptr(st["vhalf"]), _ptr(bodies[0]), 0, n, sc["dt"], sc["half"], None)
    for s in range(steps):
        do_next = 1 if (integrator == "euler" or s + 1 < steps) else 0
        _native.call("nbody_shard_force_f32", code, _ptr(bodies[s & 1]), _ptr(bodies[(s & 1) ^ 1]), n, 0, n, 0, n, 0, 0, 0,
                     1, _ptr(st["pos"]), _ptr(st["vel"]), _ptr(st["acc"]), _ptr(st["vhalf"]), sc["g"], sc["eps2"], sc["dt"],
                     sc["half"], do_next, _ptr(ws), ws.numel(), None)
    last = states[-1]
    for key, got in (("pos", last.positions), ("vel", last.velocities)):
        want = st[key].cpu().numpy()
        assert np.abs(got.numpy() - want).max() <= 1e-6 * np.abs(want).max(), key
    err = rel_rows(last.accelerations.numpy(), st["acc"].cpu().numpy())
    assert np.quantile(err, 0.999) <= 3e-6 and err.max() <= 1e-4, (np.quantile(err, 0.999), err.max())
    # energies of the recorded states are finite and conserved to the reference's level over 3 steps
    e = np.array([[s.u_energy, s.k_energy] for s in states])
    assert np.isfinite(e).all() and abs(e[-1].sum() - e[0].sum()) <= 1e-4 * abs(e[0].sum())
    # simulator state after run() is the last recorded state
    assert torch.equal(sim.positions.cpu(), last.positions) and torch.equal(sim.velocities.cpu(), last.velocities)


def test_pair_path_is_what_runs_at_large_n():
    """Launch accounting: one step at the pair path's lower limit is a pair launch + a finish launch (+ prep and plan),
    and one body fewer takes the directed kernel (prep + one fused force launch)."""
    from galaxify import _native, galaxies, simulation

    n_min = _native.lib().nbody_pair_min_bodies()
    assert n_min == 32768
    pos, vel, mass = galaxies.generate_disk(n_bodies=n_min - 1, seed=1, **KW)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    before = _native.launch_count()
    sim.step()
    torch.cuda.synchronize()
    assert _native.launch_count() - before == 2  # prep, force
    pos, vel, mass = galaxies.generate_disk(n_bodies=n_min, seed=1, **KW)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    before = _native.launch_count()
    sim.step()
    torch.cuda.synchronize()
    assert _native.launch_count() - before == 4  # prep, plan, pair, finish
