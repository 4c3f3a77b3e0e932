"""world_size-2 (and 3) CPU tests of the i-sharded host logic over gloo.

The product path's device calls (`_prepare`, `_force`: C-ABI kernels) are replaced, in this test only, by an
emulation built on the CPU oracle, so that what is exercised is the part that has no GPU in it: the slot layout,
the j-range parts, the in-place all-gather, the step protocol (own part first, epilogue on the last part, the
leapfrog re-opening) and the state recording. The result must equal the single-process oracle run.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _emulated_class(base):
    """Subclass of a sharded simulator whose two device calls are oracle-backed CPU emulations."""
    from oracle import galaxify_oracle as oracle

    f = np.float32

    class Emulated(base):
        def _pick_device(self, device):
            return torch.device("cpu")

        def _alloc_workspace(self):
            return {"sum": None, "seen": 0}

        def _prepare(self, integrator, bodies):
            pos, vel, acc = self.positions.numpy(), self.velocities.numpy(), self.accelerations.numpy()
            if integrator == 1:  # opening half-kick + drift, separately rounded (simulation.py:164-166)
                vh = vel + f(0.5 * self.dt) * acc
                self._vhalf.copy_(torch.from_numpy(vh))
                pos += f(self.dt) * vh
            sl = slice(self.i_begin, self.i_begin + self.n_local)
            bodies[sl, :3] = torch.from_numpy(pos)
            bodies[sl, 3] = self.masses

        def _force(self, integrator, bodies, bodies_next, part, j_ranges, do_next):
            ws = self._workspace
            if part == 0:
                ws["sum"], ws["seen"] = np.zeros((self.n_local, 3)), 0
            b = bodies.numpy()
            own_lo, own_hi = self.i_begin, self.i_begin + self.n_local
            mine = b[own_lo:own_hi]
            own = oracle.accelerations_f64(mine[:, :3], mine[:, 3], 1.0, self.softening)

            def others(lo, hi):  # un-scaled sum over the j slots [lo, hi), none of them the rank's own bodies
                if hi <= lo:
                    return 0.0
                both = np.concatenate([mine, b[lo:hi]])
                full = oracle.accelerations_f64(both[:, :3], both[:, 3], 1.0, self.softening, rows=slice(0, self.n_local))
                return full - own

            # FP64 partial sum of this part's ranges (the kernel: FP32 runs folded into FP64)
            for lo, hi in j_ranges:
                if hi <= lo:
                    continue
                if lo <= own_lo and own_hi <= hi:  # the range contains the own slice: split around it
                    ws["sum"] += others(lo, own_lo) + own + others(own_hi, hi)
                else:
                    assert hi <= own_lo or lo >= own_hi, "ranges may not cut through the own slice"
                    ws["sum"] += others(lo, hi)
            ws["seen"] += 1
            if ws["seen"] < len(self._parts):
                return
            a = (f(self.g_const) * ws["sum"].astype(f)).astype(f)
            self.accelerations.copy_(torch.from_numpy(a))
            if integrator == 0:
                return
            pos, vel = self.positions.numpy(), self.velocities.numpy()
            if integrator == 1:
                h = f(0.5 * self.dt)
                v = self._vhalf.numpy() + h * a
                vel[:] = v
                if not do_next:
                    return
                v = v + h * a
                self._vhalf.copy_(torch.from_numpy(v))
            else:
                v = vel + f(self.dt) * a
                vel[:] = v
            pos += f(self.dt) * v
            sl = slice(self.i_begin, self.i_begin + self.n_local)
            bodies_next[sl, :3] = torch.from_numpy(pos)
            bodies_next[sl, 3] = self.masses

        # ---- pair path (csrc/pair.cuh): plan / force / reduce-scatter / finish, emulated in numpy FP64 from the block
        # list the library itself reports (nbody_shard_pair_blocks is a pure host function)

        def _use_pair(self):
            return False  # the golden systems are far below the size limit; tests ask for the pair path explicitly

        def _setup_pair(self):
            import ctypes

            from galaxify import _native

            total = self.world_size * self.n_pad
            self._acc64 = torch.zeros(total * 3, dtype=torch.float64)
            self._acc_own = self._acc64 if self.world_size == 1 else torch.zeros(self.n_pad * 3, dtype=torch.float64)
            self._pair_split = bool(self._overlap_requested is True and self.world_size > 1)
            buf = (ctypes.c_int * (5 * 16))()
            k = _native.lib().nbody_shard_pair_blocks(self.n, self.world_size, self.n_pad, self.rank, buf, 16)
            assert k >= 1, k
            self._blocks = [tuple(buf[5 * b : 5 * b + 5]) for b in range(k)]
            self.launches_per_step = (2 if self._pair_split else 1) + 1

        def _pair_force(self, phase, bodies):
            b = bodies.numpy().astype(np.float64)
            acc = self._acc64.view(-1, 3).numpy()
            eps2 = float(f(self.softening**2))
            blocks = self._blocks if not self._pair_split else (self._blocks[:1] if phase == 0 else self._blocks[1:])
            for i_lo, i_hi, j_lo, j_hi, tri in blocks:
                xi, mi, xj, mj = b[i_lo:i_hi, :3], b[i_lo:i_hi, 3], b[j_lo:j_hi, :3], b[j_lo:j_hi, 3]
                d = xj[None, :, :] - xi[:, None, :]
                w = ((d * d).sum(-1) + eps2) ** -1.5
                if tri:
                    w = np.triu(w, 1)  # every unordered pair of the slot once
                acc[i_lo:i_hi] += (w[:, :, None] * d * mj[None, :, None]).sum(1)   # force on the i-bodies
                acc[j_lo:j_hi] -= (w[:, :, None] * d * mi[:, None, None]).sum(0)   # reaction on the j-bodies

        def _pair_reduce(self):
            if self.world_size > 1:  # gloo has no reduce_scatter: all-reduce and keep the own slot
                full = self._acc64.clone()
                dist.all_reduce(full, group=self.group)
                self._acc_own.copy_(full.view(self.world_size, -1)[self.rank])

        def _pair_finish(self, integrator, bodies, bodies_next, do_next):
            own = self._acc_own.view(-1, 3)[: self.n_local].numpy().copy()
            self._acc64.zero_()
            self._acc_own.zero_()
            self._workspace = {"sum": own, "seen": 0}
            self._epilogue(integrator, bodies_next, do_next)

        def _epilogue(self, integrator, bodies_next, do_next):
            ws = self._workspace
            a = (f(self.g_const) * ws["sum"].astype(f)).astype(f)
            self.accelerations.copy_(torch.from_numpy(a))
            if integrator == 0:
                return
            pos, vel = self.positions.numpy(), self.velocities.numpy()
            if integrator == 1:
                h = f(0.5 * self.dt)
                v = self._vhalf.numpy() + h * a
                vel[:] = v
                if not do_next:
                    return
                v = v + h * a
                self._vhalf.copy_(torch.from_numpy(v))
            else:
                v = vel + f(self.dt) * a
                vel[:] = v
            pos += f(self.dt) * v
            sl = slice(self.i_begin, self.i_begin + self.n_local)
            bodies_next[sl, :3] = torch.from_numpy(pos)
            bodies_next[sl, 3] = self.masses

    return Emulated


def _worker(rank, world, port, case, integrator, steps, out_dir, overlap, pair=None):
    import sys

    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from galaxify import sharded

        g = load_golden(case)
        base = sharded.ShardedLeapFrogSimulator if integrator == "leapfrog" else sharded.ShardedEulerSimulator
        sim = _emulated_class(base)(positions=g["ic_pos"], velocities=g["ic_vel"], masses=g["ic_mass"], overlap=overlap,
                                    pair=pair, **g.sim)
        assert sim.pair == bool(pair)
        acc0 = sim.gather_state()[2].numpy()
        states = sim.run(steps)
        pos, vel, acc = (t.numpy() for t in sim.gather_state())
        first = states[0].positions.numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), acc0=acc0, pos=pos, vel=vel, acc=acc, first=first,
                 n_states=len(states), i_begin=sim.i_begin, n_local=sim.n_local)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,integrator,overlap", [(2, "spiral_n25_leapfrog", "leapfrog", True),
                                                           (2, "disk_n500_euler", "euler", False),
                                                           (3, "disk_n500_leapfrog", "leapfrog", True),
                                                           (3, "spiral_n500_euler", "euler", None)])
def test_sharded_protocol_matches_single_process_oracle(tmp_path, world, case, integrator, overlap):
    from oracle import galaxify_oracle as oracle

    steps = 5
    mp.spawn(_worker, args=(world, _free_port(), case, integrator, steps, str(tmp_path), overlap), nprocs=world, join=True)
    g = load_golden(case)
    ref, st = oracle.run(g["ic_pos"], g["ic_vel"], g["ic_mass"], integrator=integrator, steps=steps, **g.sim)
    acc0 = oracle.accelerations(g["ic_pos"], g["ic_mass"], g.sim["g_const"], g.sim["softening"]).numpy()
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert int(z["n_states"]) == steps
        scale = lambda a: max(np.abs(a).max(), 1e-30)
        assert np.abs(z["acc0"] - acc0).max() <= 2e-6 * scale(acc0)
        assert np.abs(z["pos"] - ref[steps - 1]["pos"]).max() <= 1e-6 * scale(ref[steps - 1]["pos"])
        assert np.abs(z["vel"] - ref[steps - 1]["vel"]).max() <= 1e-6 * scale(ref[steps - 1]["vel"])
        assert np.abs(z["acc"] - ref[steps - 1]["acc"]).max() <= 2e-6 * scale(ref[steps - 1]["acc"])
        lo, cnt = int(z["i_begin"]), int(z["n_local"])
        n_pad = -(-g.n // world)
        first_rows = slice(r * n_pad, r * n_pad + cnt)
        assert np.abs(z["first"] - ref[0]["pos"][first_rows]).max() <= 1e-6 * scale(ref[0]["pos"])


@pytest.mark.parametrize("world,case,integrator,overlap", [(2, "disk_n500_leapfrog", "leapfrog", None),
                                                           (2, "spiral_n500_euler", "euler", True),
                                                           (3, "disk_n500_leapfrog", "leapfrog", True),
                                                           (4, "spiral_n1024_leapfrog", "leapfrog", None)])
def test_sharded_pair_protocol_matches_single_process_oracle(tmp_path, world, case, integrator, overlap):
    """The pair path's host protocol (plan, one or two force phases, reduce-scatter, finish, accumulators cleared for
    the next step) over gloo, with the device calls emulated from the library's own block list."""
    from oracle import galaxify_oracle as oracle

    steps = 4
    mp.spawn(_worker, args=(world, _free_port(), case, integrator, steps, str(tmp_path), overlap, True), nprocs=world,
             join=True)
    g = load_golden(case)
    ref, st = oracle.run(g["ic_pos"], g["ic_vel"], g["ic_mass"], integrator=integrator, steps=steps, **g.sim)
    acc0 = oracle.accelerations(g["ic_pos"], g["ic_mass"], g.sim["g_const"], g.sim["softening"]).numpy()
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        scale = lambda a: max(np.abs(a).max(), 1e-30)
        assert int(z["n_states"]) == steps
        assert np.abs(z["acc0"] - acc0).max() <= 2e-6 * scale(acc0)
        assert np.abs(z["pos"] - ref[steps - 1]["pos"]).max() <= 1e-6 * scale(ref[steps - 1]["pos"])
        assert np.abs(z["vel"] - ref[steps - 1]["vel"]).max() <= 1e-6 * scale(ref[steps - 1]["vel"])
        assert np.abs(z["acc"] - ref[steps - 1]["acc"]).max() <= 2e-6 * scale(ref[steps - 1]["acc"])


def test_layout_and_parts_cover_every_body_exactly_once():
    from galaxify.sharded import shard_layout, step_parts

    for n, world in ((16, 4), (17, 4), (1 << 20, 8), (5, 2), (1000, 3), (262144, 8), (7, 1)):
        n_pad, counts = shard_layout(n, world)
        assert sum(counts) == n and max(counts) == n_pad
        for rank in range(world):
            for overlap in (True, False):
                parts = step_parts(rank, n_pad, counts, overlap)
                assert len(parts) == (2 if overlap and world > 1 else 1)
                covered = np.zeros(world * n_pad, dtype=int)
                for ranges in parts:
                    for lo, hi in ranges:
                        covered[lo:hi] += 1
                real = np.zeros(world * n_pad, dtype=bool)
                for r, c in enumerate(counts):
                    real[r * n_pad : r * n_pad + c] = True
                np.testing.assert_array_equal(covered[real], 1)  # every body exactly once
                assert covered[~real].max(initial=0) <= 1  # padding entries (massless, far away) at most once
                if overlap and world > 1:
                    assert parts[0] == ((rank * n_pad, rank * n_pad + counts[rank]), (0, 0))
