"""Batched many-small-systems path (BASELINE.json configs[2]): every system equals the single-system path/oracle."""

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rows

pytestmark = pytest.mark.gpu

S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)


def make_batch(n_systems, n, kind="spiral"):
    from galaxify import galaxies

    gen = galaxies.generate_spiral if kind == "spiral" else galaxies.generate_disk
    ics = [gen(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01,
               seed=100 + s) for s in range(n_systems)]
    return (np.stack([i[0] for i in ics]), np.stack([i[1] for i in ics]), np.stack([i[2] for i in ics]))


@pytest.mark.parametrize("integrator", ["leapfrog", "euler"])
@pytest.mark.parametrize("n", [3, 100, 512, 700, 1500])
def test_batched_equals_single_system_path(integrator, n):
    from galaxify import batched, simulation

    pos, vel, mass = make_batch(5, n)
    bcls = batched.BatchedLeapFrogSimulator if integrator == "leapfrog" else batched.BatchedEulerSimulator
    scls = simulation.LeapFrogSimulator if integrator == "leapfrog" else simulation.EulerSimulator
    steps = 20
    b = bcls(positions=pos, velocities=vel, masses=mass, **S01)
    states = b.run(steps, record_every=5)
    assert [s.step for s in states] == [4, 9, 14, 19]
    assert states[0].positions.shape == (5, n, 3)
    for s_idx in (0, 4):
        single = scls(positions=pos[s_idx], velocities=vel[s_idx], masses=mass[s_idx], calc_energy=False, **S01)
        ref = single.run(steps)
        for k, st in enumerate(states):
            want = ref[st.step]
            for got, w in ((st.positions[s_idx], want.positions), (st.velocities[s_idx], want.velocities)):
                assert np.abs(got.numpy() - w.numpy()).max() <= 1e-6 * max(np.abs(w.numpy()).max(), 1e-30)
            assert rel_rows(st.accelerations[s_idx].numpy(), want.accelerations.numpy()).max() <= 1e-5
    np.testing.assert_array_equal(b.positions.cpu().numpy(), states[-1].positions.numpy())
    np.testing.assert_array_equal(b.accelerations.cpu().numpy(), states[-1].accelerations.numpy())


def test_batched_matches_reference_golden():
    from galaxify import batched

    g = load_golden("spiral_n500_leapfrog")
    reps = 3
    b = batched.BatchedLeapFrogSimulator(positions=np.stack([g["ic_pos"]] * reps), velocities=np.stack([g["ic_vel"]] * reps),
                                         masses=np.stack([g["ic_mass"]] * reps), **g.sim)
    assert rel_rows(b.accelerations[1].cpu().numpy(), g["acc0"]).max() <= 1e-5
    states = b.run(g.steps, record_every=1, to_host=False)
    for k, s in enumerate(g.keep):
        st = states[s]
        for r in range(reps):
            for got, want in ((st.positions[r], g["pos"][k]), (st.velocities[r], g["vel"][k])):
                assert np.abs(got.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()
            assert rel_rows(st.accelerations[r].cpu().numpy(), g["acc"][k]).max() <= 1e-5
    # identical systems in one batch evolve identically (no cross-talk between CTAs)
    np.testing.assert_array_equal(states[-1].positions[0].cpu().numpy(), states[-1].positions[2].cpu().numpy())


def test_batched_zero_softening_and_step():
    from galaxify import batched

    g = load_golden("spiral_n25_eps0_leapfrog")
    b = batched.BatchedLeapFrogSimulator(positions=g["ic_pos"][None], velocities=g["ic_vel"][None],
                                         masses=g["ic_mass"][None], **g.sim)
    assert torch.isfinite(b.accelerations).all()
    assert rel_rows(b.accelerations[0].cpu().numpy(), g["acc0"]).max() <= 1e-5
    for _ in range(10):
        b.step()
    assert np.abs(b.positions[0].cpu().numpy() - g["pos"][-1]).max() <= 1e-6 * np.abs(g["pos"][-1]).max()


def test_batched_energies_match_reference_and_single_path():
    from galaxify import batched, simulation

    g = load_golden("spiral_n500_leapfrog")
    other = load_golden("disk_n500_leapfrog")
    b = batched.BatchedLeapFrogSimulator(positions=np.stack([g["ic_pos"], other["ic_pos"]]),
                                         velocities=np.stack([g["ic_vel"], other["ic_vel"]]),
                                         masses=np.stack([g["ic_mass"], other["ic_mass"]]), calc_energy=True, **g.sim)
    states = b.run(100)
    u = np.array([s.u_energy for s in states])
    k = np.array([s.k_energy for s in states])
    assert u.shape == (100, 2)
    for col, gold in ((0, g), (1, other)):
        assert np.abs(u[:, col] - gold["u"][:100]).max() <= 1e-5 * np.abs(gold["u"]).max()
        assert np.abs(k[:, col] - gold["k"][:100]).max() <= 1e-5 * np.abs(gold["k"]).max()


def test_shard_systems_partitions():
    from galaxify.batched import shard_systems

    for total, world in ((4096, 8), (10, 3), (5, 8)):
        seen = []
        for r in range(world):
            sl = shard_systems(total, r, world)
            seen.extend(range(total)[sl])
        assert seen == list(range(total))
