"""The i-sharded kernel path on ONE GPU: P ranks emulated sequentially on one device, through the C ABI.

Every "rank" has its own state slice and workspace and calls nbody_shard_prepare_f32 / nbody_shard_force_f32 exactly
as galaxify.sharded.ShardedSimulator does on its own GPU (own slice first, then the two-range rest: the overlap path
that the multi-GPU benchmark times); the all-gather is the ranks writing their slots of one shared body array. The
concatenated result must match the FP64 oracle (accelerations, <= 1e-5 per particle) and the single-GPU path
(positions / velocities <= 1e-6 of the maximum). No reference counterpart: the reference is single-device
(simulation.py:46-51); the contract is SURVEY.md §8e + §8c ("1 vs 2/4/8 GPUs").
"""

import ctypes

import numpy as np
import pytest
import torch

from conftest import assert_accelerations_agree, rel_rows

pytestmark = pytest.mark.gpu

S01 = dict(g_const=4.5e-6, softening=0.05, dt=1e-4)


def _system(n, merger=False):
    from galaxify import galaxies

    kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=S01["g_const"], black_hole_mass=0.01)
    if not merger:
        return galaxies.generate_disk(n_bodies=n, seed=n, **kw)
    a = galaxies.generate_disk(n_bodies=n // 2, seed=11, **kw)
    b = galaxies.generate_disk(n_bodies=n - n // 2, seed=12, offset=(12.0, 3.0, 1.0), initial_vel=(-0.002, 0.0, 0.0),
                               angle=(0.4, 0.2, 0.0), clockwise=False, **kw)
    return galaxies.merge(a, b)


class EmulatedRanks:
    """P ShardedSimulator-like ranks on one device, stepping through the same C-ABI calls in rank order."""

    def __init__(self, pos, vel, mass, world, integrator, overlap=True):
        from galaxify import _native, sharded
        from galaxify.simulation import _ptr

        self.nat, self.ptr = _native, _ptr
        self.world, self.integrator = world, integrator
        self.n = len(mass)
        self.n_pad, self.counts = sharded.shard_layout(self.n, world)
        self.total = world * self.n_pad
        dev = torch.device("cuda")
        self.bodies = [torch.zeros((self.total, 4), dtype=torch.float32, device=dev) for _ in range(2)]
        self.parts, self.state, self.ws = [], [], []
        f = np.float32
        for r in range(world):
            lo, cnt = r * self.n_pad, self.counts[r]
            for b in self.bodies:
                b[lo + cnt : lo + self.n_pad, :3] = 1e18
            sl = slice(lo, lo + cnt)
            st = dict(pos=torch.tensor(pos[sl].astype(f), device=dev), vel=torch.tensor(vel[sl].astype(f), device=dev),
                      mass=torch.tensor(mass[sl].astype(f), device=dev))
            st["acc"], st["vhalf"] = torch.zeros_like(st["pos"]), torch.zeros_like(st["pos"])
            self.state.append(st)
            self.parts.append(sharded.step_parts(r, self.n_pad, self.counts, overlap))
            need = _native.lib().nbody_shard_workspace_bytes(cnt, self.total, len(self.parts[r]))
            self.ws.append(torch.empty(need, dtype=torch.uint8, device=dev))
        self.sc = dict(g=_native.f32(S01["g_const"]), eps2=_native.f32(S01["softening"] ** 2), dt=_native.f32(S01["dt"]),
                       half_dt=_native.f32(0.5 * S01["dt"]))
        self._prepare_all(0, self.bodies[0])
        self._force_all(0, self.bodies[0], None, 0)

    def _prepare_all(self, integ, bodies):
        p = self.ptr
        for r, st in enumerate(self.state):
            self.nat.call("nbody_shard_prepare_f32", integ, p(st["pos"]), p(st["vel"]), p(st["acc"]), p(st["mass"]),
                          p(st["vhalf"]), p(bodies), r * self.n_pad, self.counts[r], self.sc["dt"], self.sc["half_dt"], None)

    def _force_all(self, integ, bodies, bodies_next, do_next):
        p, sc = self.ptr, self.sc
        for r, st in enumerate(self.state):
            for part, ((j0, j1), (k0, k1)) in enumerate(self.parts[r]):
                self.nat.call("nbody_shard_force_f32", integ, p(bodies), p(bodies_next), self.total, r * self.n_pad,
                              self.counts[r], j0, j1, k0, k1, part, len(self.parts[r]), p(st["pos"]), p(st["vel"]),
                              p(st["acc"]), p(st["vhalf"]), sc["g"], sc["eps2"], sc["dt"], sc["half_dt"], do_next,
                              p(self.ws[r]), self.ws[r].numel(), None)

    def advance(self, steps):
        integ = self.integrator
        cur = 0
        self._prepare_all(integ, self.bodies[cur])
        for s in range(steps):
            do_next = 1 if (integ == self.nat.INTEGRATOR_EULER or s + 1 < steps) else 0
            self._force_all(integ, self.bodies[cur], self.bodies[cur ^ 1], do_next)
            cur ^= 1
        torch.cuda.synchronize()

    def energies(self, bodies):
        """Sum over ranks of nbody_shard_energies_f32, each with the rank's own (2-part) workspace."""
        p = self.ptr
        total = np.zeros(2)
        for r, st in enumerate(self.state):
            out = torch.empty(2, dtype=torch.float64, device="cuda")
            self.nat.call("nbody_shard_energies_f32", p(bodies), p(st["vel"]), self.total, r * self.n_pad, self.counts[r],
                          self.sc["g"], self.nat.f32(S01["softening"]), p(out), p(self.ws[r]), self.ws[r].numel(), None)
            total += out.cpu().numpy()
        return total

    def gathered(self, key):
        return torch.cat([st[key] for st in self.state]).cpu().numpy()


@pytest.mark.parametrize("world,n,integrator,merger", [
    (2, 20001, "leapfrog", False), (4, 20001, "euler", False), (8, 4099, "leapfrog", False),
    (3, 10007, "leapfrog", False), (8, 65536, "euler", True), (2, 262144, "leapfrog", True),
    (8, 262144, "leapfrog", True), (4, 1000, "leapfrog", False), (2, 3075, "leapfrog", False),
    (4, 12289, "euler", False)])
def test_emulated_ranks_match_oracle_and_single_gpu(world, n, integrator, merger):
    from galaxify import _native, simulation
    from oracle import c_oracle

    pos, vel, mass = _system(n, merger)
    code = _native.INTEGRATOR_LEAPFROG if integrator == "leapfrog" else _native.INTEGRATOR_EULER
    steps = 3
    em = EmulatedRanks(pos, vel, mass, world, code, overlap=True)
    assert all(len(p) == 2 for p in em.parts)  # the two-launch / two-range path is what runs

    # initial accelerations vs the FP64 oracle on sampled rows
    rows = np.unique(np.concatenate([np.arange(0, n, max(1, n // 192)), [0, n - 1, em.n_pad - 1, em.n_pad % n]]))
    want = np.stack([c_oracle.accelerations_f64(pos, mass, S01["g_const"], S01["softening"], int(i), int(i) + 1)[0]
                     for i in rows])
    acc0 = em.gathered("acc")
    assert rel_rows(acc0[rows], want).max() <= 1e-5

    # energies of the initial state through the sharded entry point, with the workspace sized for the 2-part step
    # (the planner may want more splits for one sweep than for two half sweeps, n ~ 3k..16k)
    if n <= 65536:
        u, k = em.energies(em.bodies[0])
        want_u, want_k = c_oracle.energies_f64(pos, vel, mass, S01["g_const"], S01["softening"])
        assert abs(u - want_u) <= 5e-6 * abs(want_u) and abs(k - want_k) <= 5e-6 * abs(want_k)

    cls = simulation.LeapFrogSimulator if integrator == "leapfrog" else simulation.EulerSimulator
    single = cls(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    assert_accelerations_agree(acc0, single.accelerations.cpu().numpy(), pos, mass, S01["g_const"], S01["softening"])
    ref = single.run(steps)[-1]
    em.advance(steps)
    for key, want_t in (("pos", ref.positions), ("vel", ref.velocities)):
        w = want_t.numpy()
        assert np.abs(em.gathered(key) - w).max() <= 1e-6 * np.abs(w).max(), key
    acc = em.gathered("acc")
    err = rel_rows(acc, ref.accelerations.numpy())
    assert np.quantile(err, 0.999) <= 3e-6 and err.max() <= 1e-4, (np.quantile(err, 0.999), err.max())
    # and the final accelerations against the oracle at the final positions
    p_fin = em.gathered("pos") if integrator == "euler" else ref.positions.numpy()
    want = np.stack([c_oracle.accelerations_f64(p_fin, mass, S01["g_const"], S01["softening"], int(i), int(i) + 1)[0]
                     for i in rows])
    if integrator == "leapfrog":  # leapfrog's recorded acceleration belongs to the recorded positions
        assert rel_rows(acc[rows], want).max() <= 1e-5


class EmulatedPairRanks(EmulatedRanks):
    """The same emulation for the pair path (csrc/pair.cuh): every rank accumulates forces AND reactions for the part
    of the interaction matrix it owns into its own full-size FP64 accumulator; the reduce-scatter is a sum of the
    ranks' accumulators slot by slot; nbody_shard_pair_finish_f32 applies the integrator."""

    def __init__(self, pos, vel, mass, world, integrator, split):
        from galaxify import _native, sharded

        self.split = split
        n_pad, _ = sharded.shard_layout(len(mass), world)
        dev = torch.device("cuda")
        self.acc64 = [torch.zeros(world * n_pad * 3, dtype=torch.float64, device=dev) for _ in range(world)]
        self.acc_own = [torch.zeros(n_pad * 3, dtype=torch.float64, device=dev) for _ in range(world)]
        need = _native.lib().nbody_shard_pair_workspace_bytes(world, n_pad)
        self.pws = [torch.empty(need, dtype=torch.uint8, device=dev) for _ in range(world)]
        for r in range(world):
            _native.call("nbody_shard_pair_plan_f32", len(mass), world, n_pad, r, int(split), self.pws[r].data_ptr(),
                         self.pws[r].numel(), None)
        super().__init__(pos, vel, mass, world, integrator, overlap=True)

    def _force_all(self, integ, bodies, bodies_next, do_next):
        p, sc = self.ptr, self.sc
        for r in range(self.world):
            for phase in ((0, 1) if self.split else (0,)):
                self.nat.call("nbody_shard_pair_force_f32", phase, p(bodies), self.world, self.n_pad, sc["eps2"],
                              p(self.acc64[r]), p(self.pws[r]), self.pws[r].numel(), None)
        full = torch.stack(self.acc64).sum(0).view(self.world, self.n_pad * 3)  # what the reduce-scatter delivers
        for r, st in enumerate(self.state):
            self.acc_own[r].copy_(full[r])
            self.nat.call("nbody_shard_pair_finish_f32", integ, p(bodies), p(bodies_next), r * self.n_pad, self.counts[r],
                          p(self.acc_own[r]), p(self.acc64[r]), self.acc64[r].numel(), p(st["pos"]), p(st["vel"]),
                          p(st["acc"]), p(st["vhalf"]), sc["g"], sc["dt"], sc["half_dt"], do_next, self.world, self.n_pad,
                          p(self.pws[r]), self.pws[r].numel(), None)


@pytest.mark.parametrize("world,n,integrator,merger,split", [
    (1, 66000, "leapfrog", False, False), (2, 70001, "leapfrog", False, True), (2, 70001, "euler", False, False),
    (3, 100003, "leapfrog", True, True), (4, 131072, "euler", True, True), (5, 90000, "leapfrog", False, False),
    (8, 65541, "leapfrog", False, True), (8, 262144, "leapfrog", True, True), (6, 262144, "euler", True, False),
    (8, 20001, "leapfrog", False, True)])
def test_emulated_ranks_pair_path(world, n, integrator, merger, split):
    """Every unordered pair exactly once over all ranks: own-slot triangles, cyclic rectangles, and the half
    rectangle of even world sizes; sizes where slots are not multiples of any tile."""
    from galaxify import _native, simulation
    from oracle import c_oracle

    pos, vel, mass = _system(n, merger)
    code = _native.INTEGRATOR_LEAPFROG if integrator == "leapfrog" else _native.INTEGRATOR_EULER
    steps = 3
    em = EmulatedPairRanks(pos, vel, mass, world, code, split)
    rows = np.unique(np.concatenate([np.arange(0, n, max(1, n // 192)), [0, n - 1, em.n_pad - 1, em.n_pad % n]]))
    want = np.stack([c_oracle.accelerations_f64(pos, mass, S01["g_const"], S01["softening"], int(i), int(i) + 1)[0]
                     for i in rows])
    acc0 = em.gathered("acc")
    assert np.isfinite(acc0).all()
    assert rel_rows(acc0[rows], want).max() <= 1e-5

    cls = simulation.LeapFrogSimulator if integrator == "leapfrog" else simulation.EulerSimulator
    single = cls(positions=pos, velocities=vel, masses=mass, calc_energy=False, **S01)
    assert_accelerations_agree(acc0, single.accelerations.cpu().numpy(), pos, mass, S01["g_const"], S01["softening"])
    ref = single.run(steps)[-1]
    em.advance(steps)
    for key, want_t in (("pos", ref.positions), ("vel", ref.velocities)):
        w = want_t.numpy()
        assert np.abs(em.gathered(key) - w).max() <= 1e-6 * np.abs(w).max(), key
    # the accumulators are left clean for the next step
    assert all(float(a.abs().max()) == 0.0 for a in em.acc64)
