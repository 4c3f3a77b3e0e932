"""i-sharded path on real GPUs: world size 1 always (exercises the shard kernels + NCCL plumbing on one GPU), and
world size 2 when the box has two GPUs. The sharded result must equal the single-GPU path to FP32 summation order."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT, assert_accelerations_agree, load_golden, rel_rows

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, integrator, steps, out_dir, overlap):
    import sys

    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from galaxify import galaxies, sharded

        pos, vel, mass = galaxies.generate_disk(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3,
                                                g_const=4.5e-6, black_hole_mass=0.01, seed=n)
        cls = sharded.ShardedLeapFrogSimulator if integrator == "leapfrog" else sharded.ShardedEulerSimulator
        sim = cls(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6, softening=0.05, dt=1e-4, calc_energy=True,
                  overlap=overlap)
        acc0 = sim.gather_state()[2].numpy()
        e0 = sim.compute_energies()
        states = sim.run(steps)
        p, v, a = (t.numpy() for t in sim.gather_state())
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), acc0=acc0, pos=p, vel=v, acc=a, n_states=len(states), e0=e0,
                     e=np.array([[st.u_energy, st.k_energy] for st in states]),
                     first=states[0].positions.numpy(), n_local=sim.n_local)
    finally:
        dist.destroy_process_group()


def _check(tmp_path, world, n, integrator, steps, overlap=None):
    from galaxify import galaxies, simulation

    mp.spawn(_worker, args=(world, _free_port(), n, integrator, steps, str(tmp_path), overlap), nprocs=world, join=True)
    z = np.load(os.path.join(str(tmp_path), "out.npz"))
    pos, vel, mass = galaxies.generate_disk(n_bodies=n, total_mass=1.0, radial_scale=3.0, height_scale=0.3,
                                            g_const=4.5e-6, black_hole_mass=0.01, seed=n)
    cls = simulation.LeapFrogSimulator if integrator == "leapfrog" else simulation.EulerSimulator
    single = cls(positions=pos, velocities=vel, masses=mass, g_const=4.5e-6, softening=0.05, dt=1e-4, calc_energy=True)
    np.testing.assert_allclose(z["e0"], single.compute_energies(), rtol=1e-6)
    # two FP32 evaluations with different j-split orders: each is within ~1e-6 of exact, bar is the north-star 1e-5
    assert_accelerations_agree(z["acc0"], single.accelerations.cpu().numpy(), pos, mass, 4.5e-6, 0.05)
    ref = single.run(steps)
    assert int(z["n_states"]) == steps
    np.testing.assert_allclose(z["e"], [[st.u_energy, st.k_energy] for st in ref], rtol=1e-6)
    for key, want in (("pos", ref[-1].positions), ("vel", ref[-1].velocities)):
        want = want.numpy()
        assert np.abs(z[key] - want).max() <= 1e-6 * np.abs(want).max(), key
    err = rel_rows(z["acc"], ref[-1].accelerations.numpy())
    assert np.quantile(err, 0.999) <= 3e-6 and err.max() <= 1e-4, (np.quantile(err, 0.999), err.max())
    first = ref[0].positions.numpy()[: int(z["n_local"])]
    assert np.abs(z["first"] - first).max() <= 1e-6 * np.abs(first).max()


@pytest.mark.parametrize("integrator", ["leapfrog", "euler"])
@pytest.mark.parametrize("n", [1000, 20001, 70001])
def test_sharded_world1_equals_single_gpu(tmp_path, n, integrator):
    """n = 70,001 takes the pair path (plan / force / finish through ShardedSimulator, no collective at world 1)."""
    _check(tmp_path, 1, n, integrator, 4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("integrator,n,overlap", [("leapfrog", 20001, True), ("leapfrog", 20001, False),
                                                  ("euler", 4097, True), ("euler", 4096, False),
                                                  ("leapfrog", 262144, True), ("leapfrog", 262144, None),
                                                  ("euler", 70001, None)])
def test_sharded_world2_equals_single_gpu(tmp_path, integrator, n, overlap):
    _check(tmp_path, 2, n, integrator, 3, overlap)
