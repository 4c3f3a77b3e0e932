"""Edge cases of the reference-facing API on the GPU: chunked runs, record_every, tiny systems, in-place mutation,
tensor inputs, empty runs — the behaviours src/s01-dataset-generation.py and ad-hoc users of the reference rely on."""

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rows

pytestmark = pytest.mark.gpu

S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)


def sim_from(g, cls_name=None, **over):
    from galaxify import simulation

    cls = getattr(simulation, cls_name or ("LeapFrogSimulator" if g.integrator == "leapfrog" else "EulerSimulator"))
    kw = dict(g.sim)
    kw.update(over)
    return cls(positions=g["ic_pos"], velocities=g["ic_vel"], masses=g["ic_mass"], **kw)


@pytest.mark.parametrize("tiled", [False, True])
def test_chunked_run_equals_single_chunk(monkeypatch, tiled):
    """run() stages the trajectory through bounded device chunks; the chunk boundary must not show."""
    from galaxify import simulation

    if tiled:
        monkeypatch.setattr(simulation, "PERSISTENT_MAX_N", 0)
    g = load_golden("spiral_n500_leapfrog")
    whole = sim_from(g, calc_energy=True).run(23)
    monkeypatch.setattr(simulation, "TRAJ_CHUNK_BYTES", 36 * 500 * 5)  # 5 states per chunk
    parts = sim_from(g, calc_energy=True).run(23)
    assert [s.step for s in parts] == list(range(23))
    for a, b in zip(whole, parts):
        np.testing.assert_array_equal(a.positions.numpy(), b.positions.numpy())
        np.testing.assert_array_equal(a.velocities.numpy(), b.velocities.numpy())
        np.testing.assert_array_equal(a.accelerations.numpy(), b.accelerations.numpy())
        assert a.u_energy == b.u_energy and a.k_energy == b.k_energy


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("integrator", ["LeapFrogSimulator", "EulerSimulator"])
def test_record_every_subsamples_the_same_trajectory(monkeypatch, tiled, integrator):
    from galaxify import simulation

    if tiled:
        monkeypatch.setattr(simulation, "PERSISTENT_MAX_N", 0)
    g = load_golden("disk_n500_leapfrog")
    every = sim_from(g, integrator, calc_energy=True).run(20)
    sparse_sim = sim_from(g, integrator, calc_energy=True)
    sparse = sparse_sim.run(20, record_every=6)
    assert [s.step for s in sparse] == [5, 11, 17]
    for s in sparse:
        np.testing.assert_array_equal(s.positions.numpy(), every[s.step].positions.numpy())
        np.testing.assert_array_equal(s.accelerations.numpy(), every[s.step].accelerations.numpy())
        assert s.u_energy == every[s.step].u_energy
    # the two trailing, unrecorded steps were still integrated
    np.testing.assert_array_equal(sparse_sim.positions.cpu().numpy(), every[19].positions.numpy())
    with pytest.raises(ValueError):
        sparse_sim.run(3, record_every=0)


def test_zero_and_negative_steps_and_base_class():
    from galaxify import simulation

    g = load_golden("spiral_n25_leapfrog")
    sim = sim_from(g)
    before = sim.positions.clone()
    assert sim.run(0) == [] and sim.run(-3) == []
    assert torch.equal(sim.positions, before)
    base = simulation.BaseSimulator(positions=g["ic_pos"], velocities=g["ic_vel"], masses=g["ic_mass"], **g.sim)
    assert rel_rows(base.accelerations.cpu().numpy(), g["acc0"]).max() <= 1e-5  # __init__ computes forces (:69)
    assert base.run(0) == []
    with pytest.raises(NotImplementedError):
        base.step()
    with pytest.raises(NotImplementedError):
        base.run(2)


@pytest.mark.parametrize("tiled", [False, True])
def test_one_and_two_bodies(monkeypatch, tiled):
    from galaxify import simulation

    if tiled:
        monkeypatch.setattr(simulation, "PERSISTENT_MAX_N", 0)
    one = simulation.LeapFrogSimulator(positions=[[1.0, 2.0, 3.0]], velocities=[[0.5, 0.0, -0.5]], masses=[2.0],
                                       g_const=1.0, softening=0.1, dt=0.01, calc_energy=True)
    assert torch.all(one.accelerations == 0)
    st = one.run(10)[-1]
    np.testing.assert_allclose(st.positions.numpy(), [[1.05, 2.0, 2.95]], rtol=1e-6)
    assert st.u_energy == 0.0 and st.k_energy == pytest.approx(0.5 * 2.0 * 0.5, rel=1e-6)
    two = simulation.EulerSimulator(positions=[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], velocities=np.zeros((2, 3)),
                                    masses=[1.0, 3.0], g_const=1.0, softening=0.0, dt=0.01, calc_energy=True)
    a = two.accelerations.cpu().numpy()
    np.testing.assert_allclose(a, [[3.0, 0, 0], [-1.0, 0, 0]], rtol=1e-6)  # softening 0: self terms masked, finite
    u, k = two.compute_energies()
    assert u == pytest.approx(-3.0, rel=1e-6) and k == 0.0


def test_in_place_mutation_is_seen_and_inputs_are_copied():
    """The reference keeps plain tensors a caller may edit between calls (simulation.py:58-65); nothing is cached."""
    from galaxify import simulation

    g = load_golden("disk_n500_leapfrog")
    pos = np.array(g["ic_pos"])
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=g["ic_vel"], masses=g["ic_mass"], **g.sim)
    pos[:] = 0.0  # the simulator copied its inputs
    assert rel_rows(sim.accelerations.cpu().numpy(), g["acc0"]).max() <= 1e-5
    sim.positions[:, 0] += 1.0  # a rigid shift leaves the forces unchanged
    shifted = sim.compute_accelerations().cpu().numpy()
    assert rel_rows(shifted, g["acc0"]).max() <= 2e-5
    sim.masses *= 2.0  # doubled masses double the accelerations
    assert rel_rows(sim.compute_accelerations().cpu().numpy(), 2.0 * g["acc0"].astype(np.float64)).max() <= 2e-5


def test_tensor_and_float64_inputs():
    from galaxify import simulation

    g = load_golden("spiral_n500_leapfrog")
    sim = simulation.LeapFrogSimulator(positions=torch.tensor(g["ic_pos"]), velocities=torch.tensor(g["ic_vel"]),
                                       masses=torch.tensor(g["ic_mass"], dtype=torch.float64).cuda(), **g.sim)
    assert sim.positions.dtype == torch.float32 and sim.positions.is_cuda and sim.n == 500
    assert sim.device.type == "cuda" and sim.calc_energy is True and sim.dt == g.sim["dt"]
    assert rel_rows(sim.accelerations.cpu().numpy(), g["acc0"]).max() <= 1e-5
    with pytest.raises(ValueError):
        simulation.LeapFrogSimulator(positions=np.zeros((4, 2)), velocities=np.zeros((4, 3)), masses=np.ones(4))
    with pytest.raises(ValueError):
        simulation.LeapFrogSimulator(positions=np.zeros((4, 3)), velocities=np.zeros((4, 3)), masses=np.ones(5))


def test_step_rebinds_accelerations_like_the_reference():
    g = load_golden("spiral_n25_leapfrog")
    sim = sim_from(g)
    old = sim.accelerations
    kept = old.clone()
    sim.step()
    assert sim.accelerations is not old and torch.equal(old, kept)  # simulation.py:168 rebinds, the old tensor survives


def test_launches_are_counted():
    from galaxify import _native

    g = load_golden("disk_n1024_leapfrog")
    sim = sim_from(g, calc_energy=False)
    before = _native.launch_count()
    sim.run(50)
    assert _native.launch_count() - before == 1  # the persistent path: one kernel for the whole run
