"""f4 — the rollout integrator with the force slot left open (trainer.py:217-226, gnn.py:223-253), momentum
diagnostics, and the tiny-softening diagonal. All through the C ABI on a GPU."""

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rows

pytestmark = pytest.mark.gpu


def _numpy_step(pos, vel, acc, acc_new_fn, dt):
    """numpy-FP32 emulation of trainer.py:217-226: every update is fl32(a + fl32(c*b)), c = fl32(0.5*dt) / fl32(dt)."""
    f = np.float32
    vel_ = vel + f(0.5 * dt) * acc
    pos_ = pos + f(dt) * vel_
    acc_ = acc_new_fn(pos_)
    vel_ = vel_ + f(0.5 * dt) * acc_
    return pos_, vel_, acc_


@pytest.mark.parametrize("n", [1, 3, 4, 1000, 4097, 262144])
@pytest.mark.parametrize("dt", [1e-4, 0.01, 0.3])
def test_kick_drift_is_bit_equal_to_numpy_fp32(n, dt):
    from galaxify import rollout

    rng = np.random.default_rng(n)
    pos, vel, acc, acc2 = (rng.standard_normal((n, 3)).astype(np.float32) * s for s in (3.0, 1e-3, 0.1, 0.1))
    want_pos, want_vel, _ = _numpy_step(pos, vel, acc, lambda p: acc2, dt)
    d = [torch.tensor(a, device="cuda") for a in (pos, vel, acc, acc2)]
    pos_, vel_ = rollout.kick_drift(d[0], d[1], d[2], dt)
    assert torch.equal(d[0].cpu(), torch.tensor(pos)) and torch.equal(d[1].cpu(), torch.tensor(vel))  # inputs untouched
    rollout.kick_(vel_, d[3], dt)
    assert np.array_equal(pos_.cpu().numpy(), want_pos)
    assert np.array_equal(vel_.cpu().numpy(), want_vel)
    # the unaligned (non-vectorised) path: views that start 4 bytes into a buffer
    if n > 1:
        big = [torch.zeros(3 * n + 1, device="cuda") for _ in range(3)]
        views = [b[1:].view(n, 3) for b in big]
        for v, a in zip(views, (pos, vel, acc)):
            v.copy_(torch.tensor(a))
        p2, v2 = rollout.kick_drift(*views, dt)
        rollout.kick_(v2, d[3], dt)
        assert np.array_equal(p2.cpu().numpy(), want_pos) and np.array_equal(v2.cpu().numpy(), want_vel)


def test_rollout_with_external_force_matches_numpy_emulation():
    """A model-like force (here: a linear spring field evaluated by torch) in the slot, 50 steps."""
    from galaxify import rollout

    n, dt, steps = 777, 0.01, 50
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32)
    m = rng.uniform(0.5, 1.5, n).astype(np.float32)
    k = np.float32(0.7)

    seen = []

    def predict(p, feats):
        seen.append(tuple(feats.shape))
        return -k * p

    mem = rollout.rollout(predict, torch.tensor(pos, device="cuda"), torch.tensor(vel, device="cuda"),
                          torch.tensor(m, device="cuda"), steps, dt)
    assert len(mem["pos"]) == len(mem["vel"]) == len(mem["acc"]) == steps + 1
    assert seen[0] == (n, 7) and seen[1] == (n, 4)  # gnn.py:247 passes cat(pos, vel, m); trainer.py:223 cat(vel_, m)
    p, v, a = pos, vel, -k * pos
    for _ in range(steps):
        p, v, a = _numpy_step(p, v, a, lambda q: -k * q, dt)
    assert np.array_equal(mem["pos"][-1].cpu().numpy(), p)
    assert np.array_equal(mem["vel"][-1].cpu().numpy(), v)
    assert np.array_equal(mem["acc"][-1].cpu().numpy(), a)


@pytest.mark.parametrize("name", ["disk_n500_leapfrog", "spiral_n1024_leapfrog"])
def test_rollout_with_direct_sum_force_is_the_leapfrog_simulator(name, monkeypatch):
    """With this engine's own force in the slot the rollout IS LeapFrogSimulator.step (simulation.py:153-170): bit
    equal to the fused tiled path, and within the golden tolerance of the reference's trajectory."""
    from galaxify import rollout, simulation

    g = load_golden(name)
    monkeypatch.setattr(simulation, "PERSISTENT_MAX_N", 0)  # same force kernel on both sides
    sim = simulation.LeapFrogSimulator(positions=g["ic_pos"], velocities=g["ic_vel"], masses=g["ic_mass"],
                                       calc_energy=False, **g.sim)
    force = rollout.DirectSumForce(g.sim["g_const"], g.sim["softening"])
    pos = torch.tensor(g["ic_pos"], dtype=torch.float32, device="cuda")
    vel = torch.tensor(g["ic_vel"], dtype=torch.float32, device="cuda")
    m = torch.tensor(g["ic_mass"], dtype=torch.float32, device="cuda")
    steps = 25
    mem = rollout.rollout(force, pos, vel, m, steps, g.sim["dt"])
    assert torch.equal(mem["acc"][0], sim.accelerations)
    states = sim.run(steps)
    assert torch.equal(mem["pos"][-1].cpu(), states[-1].positions)
    assert torch.equal(mem["vel"][-1].cpu(), states[-1].velocities)
    assert torch.equal(mem["acc"][-1].cpu(), states[-1].accelerations)
    if steps - 1 in g.keep:
        k = g.keep.index(steps - 1)
        assert np.abs(mem["pos"][-1].cpu().numpy() - g["pos"][k]).max() <= 1e-6 * np.abs(g["pos"][k]).max()


def test_rollout_rejects_cpu_tensors():
    from galaxify import rollout

    z = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rollout.kick_drift(z, z, z, 0.01)


@pytest.mark.parametrize("n", [3, 1000, 20001, 262144])
def test_momentum_matches_float64_numpy(n):
    from galaxify import galaxies, simulation

    pos, vel, mass = galaxies.generate_disk(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3,
                                            g_const=4.5e-6, black_hole_mass=0.01, seed=3, initial_vel=(0.01, -0.02, 0.003),
                                            offset=(1.0, 2.0, 3.0))
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6, softening=0.05,
                                       dt=1e-4, calc_energy=False)
    p, l = sim.compute_momentum()
    x, v, m = (a.astype(np.float32).astype(np.float64) for a in (pos, vel, mass))
    want_p = (m[:, None] * v).sum(0)
    want_l = (m[:, None] * np.cross(x, v)).sum(0)
    assert np.allclose(p, want_p, rtol=1e-12, atol=1e-18)
    assert np.allclose(l, want_l, rtol=1e-11, atol=1e-16)
    # momentum drift over a short run stays at rounding level (pairwise forces cancel), as in the reference
    sim.run(20)
    p2, _ = sim.compute_momentum()
    assert np.abs(p2 - p).max() <= 1e-6 * max(np.abs(p).max(), (m * np.linalg.norm(v, axis=1)).sum() * 1e-3)


@pytest.mark.parametrize("softening", [1e-15, 1e-10, 3e-9])
@pytest.mark.parametrize("n", [25, 3000])
def test_tiny_nonzero_softening_keeps_the_diagonal_zero(n, softening):
    """fill_diagonal_(0) (simulation.py:85) keeps the self term out even when eps2^-1.5 overflows FP32; the kernels
    must not turn it into 0*inf. Checked against the FP64 oracle on a cloud without close pairs."""
    from galaxify import simulation
    from oracle import c_oracle

    rng = np.random.default_rng(n)
    pos = rng.uniform(-1, 1, (n, 3)) * 10
    mass = rng.uniform(0.5, 1.5, n)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=np.zeros((n, 3)), masses=mass, g_const=1.0,
                                       softening=softening, dt=1e-3, calc_energy=False)
    acc = sim.accelerations.cpu().numpy()
    assert np.all(np.isfinite(acc))
    want = c_oracle.accelerations_f64(pos, mass, 1.0, softening)
    assert rel_rows(acc, want).max() <= 1e-5
