"""i-sharded large-N simulators: one process per GPU, positions all-gathered every step over NCCL.

An addition with no reference counterpart (the reference is single-device, simulation.py:46-51). The constructor
keeps the reference's keyword arguments (calc_energy defaults to False here); every rank passes the SAME full
arrays and keeps only its slice. Energies are reduced over ranks with one all-reduce of two doubles.

Layout: the body array (x,y,z,m float4) has world_size slots of `n_pad = ceil(n / world_size)` bodies; rank r owns
global indices [r*n_pad, r*n_pad + count_r). Padding entries are massless and parked far away (1e18), so kernels may
sweep whole slots: they contribute exactly zero.

Systems of at least nbody_pair_min_bodies() bodies take the PAIR path (csrc/pair.cuh): every unordered pair of bodies
is evaluated once over all ranks. A rank evaluates the triangle of its own slot (needs no remote data: it runs while
the all-gather is in flight), the rectangles against the next floor((P-1)/2) slots and, for even P, half of the
rectangle against the opposite slot; forces and reactions go to an FP64 accumulator laid out like the body array, a
reduce-scatter (NCCL, float64, 24 B per body) hands every rank the sums of its own bodies, and a finish kernel applies
the integrator. Per step: all-gather (16 B per body) -> one persistent pair launch -> reduce-scatter -> finish; with
overlap=True the own-slot triangle is a launch of its own that starts before the gather is waited for.

Smaller systems take the DIRECTED path below (every rank sums all j for its own i, deterministic split-j reduction).
One step on a rank (leapfrog; Euler differs only in the epilogue), with `overlap` (default for n >= 524,288):
    all_gather(bodies_cur)  [NCCL stream]   ||   force(part 0 = own slice)  [compute stream]
    wait for the gather, force(part 1 = everything before and after the own slot, one launch over two ranges);
    the last part to finish runs the fused kick/kick/drift epilogue and writes the rank's slice of bodies_next,
    which the next step gathers.
Without overlap (smaller systems, where one more launch per step costs more than the gather it hides): wait for the
gather, then a single launch over the whole array.
The force accumulates into the same split-j scratch as the single-GPU path, so a rank computes exactly what the
single-GPU kernel computes for its i-bodies up to FP32 summation order (the j splits differ).
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _native
from .simulation import SimulationState, _ptr


def shard_layout(n: int, world_size: int):
    """(n_pad, counts): slot size and the number of real bodies in every rank's slot."""
    n_pad = (n + world_size - 1) // world_size
    counts = [max(0, min(n_pad, n - r * n_pad)) for r in range(world_size)]
    return n_pad, counts


def step_parts(rank: int, n_pad: int, counts, overlap: bool):
    """j ranges of one step for `rank`, each part a pair of half-open slot ranges ((lo, hi), (lo2, hi2)).

    With overlap, part 0 is the rank's own slice (computed while the all-gather is in flight) and part 1 everything
    before and after the rank's slot; without, a single part sweeps the whole array once the gather is done.
    Ranges run over whole slots, padding entries included: those are massless and parked far away, so they add
    exactly zero."""
    total = len(counts) * n_pad
    lo, hi = rank * n_pad, rank * n_pad + counts[rank]
    if not overlap or len(counts) == 1:
        return [((0, total), (0, 0))]
    return [((lo, hi), (0, 0)), ((0, lo), (lo + n_pad, total))]


class ShardedSimulator:
    _integrator = None

    def __init__(self, *, positions, velocities, masses, g_const: float = 1.0, softening: float = 0.1,
                 dt: float = 0.01, calc_energy: bool = False, device: str = None, group=None, overlap: bool = None,
                 pair: bool = None):
        if device is not None and device not in ["cuda", "cpu"]:
            raise ValueError("device debe ser 'cuda', 'cpu' o None")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.dt, self.g_const, self.softening, self.calc_energy = dt, g_const, softening, calc_energy
        self.device = self._pick_device(device)

        pos = np.asarray(positions, dtype=np.float32)
        vel = np.asarray(velocities, dtype=np.float32)
        mass = np.asarray(masses, dtype=np.float32)
        self.n = pos.shape[0]
        if pos.shape != (self.n, 3) or vel.shape != (self.n, 3) or mass.shape != (self.n,):
            raise ValueError("expected positions (n,3), velocities (n,3), masses (n,)")
        self.n_pad, self.counts = shard_layout(self.n, self.world_size)
        self.n_local = self.counts[self.rank]
        if min(self.counts) < 1:
            raise ValueError(f"n = {self.n} is too small to shard over {self.world_size} ranks")
        self.i_begin = self.rank * self.n_pad
        lo = self.rank * self.n_pad
        sl = slice(lo, lo + self.n_local)
        dev = self.device
        self.positions = torch.tensor(pos[sl], device=dev)
        self.velocities = torch.tensor(vel[sl], device=dev)
        self.masses = torch.tensor(mass[sl], device=dev)
        self.accelerations = torch.zeros_like(self.positions)
        self._vhalf = torch.zeros_like(self.positions)
        self._bodies = [torch.zeros((self.world_size * self.n_pad, 4), dtype=torch.float32, device=dev) for _ in range(2)]
        for b in self._bodies:  # pad entries of the own slot: massless and far away, so that no kernel that sweeps the
            b[self.i_begin + self.n_local : self.i_begin + self.n_pad, :3] = 1e18  # whole array can divide by zero
        # Overlapping the gather with the own-slice force costs one more launch per step: worth it only when a step
        # is long compared to a launch (the gather itself is tens of microseconds either way).
        self._overlap_requested = overlap
        self.overlap = (self.n >= 524288) if overlap is None else bool(overlap)
        self._parts = step_parts(self.rank, self.n_pad, self.counts, self.overlap)
        self._workspace = self._alloc_workspace()
        self.launches_per_step = len(self._parts)
        self.pair = self._use_pair() if pair is None else bool(pair)
        if self.pair:
            self._setup_pair()
        # initial accelerations (simulation.py:69)
        self._prepare(0, self._bodies[0])
        self._gather(self._bodies[0]).wait()
        self._force_all(0, self._bodies[0], None, do_next=0)

    # ------------------------------------------------------------------ device plumbing (overridden by CPU tests)

    def _pick_device(self, device):
        if device == "cpu":
            raise RuntimeError("galaxify (B200 engine) has no CPU path")
        _native.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("galaxify (B200 engine) needs a CUDA device and found none; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def _alloc_workspace(self):
        need = _native.lib().nbody_shard_workspace_bytes(self.n_local, self.world_size * self.n_pad, len(self._parts))
        return torch.empty(need, dtype=torch.uint8, device=self.device)

    def _scalars(self):
        return dict(g=_native.f32(self.g_const), eps2=_native.f32(self.softening**2), dt=_native.f32(self.dt),
                    half_dt=_native.f32(0.5 * self.dt))

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _prepare(self, integrator, bodies):
        s = self._scalars()
        _native.call("nbody_shard_prepare_f32", integrator, _ptr(self.positions), _ptr(self.velocities),
                     _ptr(self.accelerations), _ptr(self.masses), _ptr(self._vhalf), _ptr(bodies), self.i_begin,
                     self.n_local, s["dt"], s["half_dt"], self._stream())

    def _force(self, integrator, bodies, bodies_next, part, j_ranges, do_next):
        s = self._scalars()
        ws = self._workspace
        (j0, j1), (k0, k1) = j_ranges
        _native.call("nbody_shard_force_f32", integrator, _ptr(bodies), _ptr(bodies_next),
                     self.world_size * self.n_pad, self.i_begin, self.n_local, j0, j1, k0, k1, part,
                     len(self._parts), _ptr(self.positions), _ptr(self.velocities), _ptr(self.accelerations),
                     _ptr(self._vhalf), s["g"], s["eps2"], s["dt"], s["half_dt"], do_next, _ptr(ws), ws.numel(),
                     self._stream())

    # ------------------------------------------------------------------ pair path (csrc/pair.cuh)

    def _use_pair(self):
        return self.n >= _native.lib().nbody_pair_min_bodies() and _native.f32(self.softening**2) >= 1e-16

    def _setup_pair(self):
        lib = _native.lib()
        total = self.world_size * self.n_pad
        self._acc64 = torch.zeros(total * 3, dtype=torch.float64, device=self.device)
        self._acc_own = self._acc64 if self.world_size == 1 else torch.zeros(self.n_pad * 3, dtype=torch.float64,
                                                                              device=self.device)
        need = lib.nbody_shard_pair_workspace_bytes(self.world_size, self.n_pad)
        self._pair_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        # Splitting the own-slot triangle off (to run it while the all-gather is in flight) is opt-in: the persistent
        # pair kernel fills every SM, so the NCCL kernel only gets to run in the first launch's tail anyway, and the
        # second launch adds a second tail; measured on 2 GPUs at N = 1M the gather is 0.16 ms of a 179 ms step
        # (tools/diag_sharded_pair.py, profiles/r2_diag_sharded_pair.log).
        self._pair_split = bool(self._overlap_requested is True and self.world_size > 1)
        _native.call("nbody_shard_pair_plan_f32", self.n, self.world_size, self.n_pad, self.rank, int(self._pair_split),
                     _ptr(self._pair_ws), self._pair_ws.numel(), self._stream())
        self.launches_per_step = (2 if self._pair_split else 1) + 1  # pair launch(es) + finish

    def _pair_force(self, phase, bodies):
        _native.call("nbody_shard_pair_force_f32", phase, _ptr(bodies), self.world_size, self.n_pad,
                     _native.f32(self.softening**2), _ptr(self._acc64), _ptr(self._pair_ws), self._pair_ws.numel(),
                     self._stream())

    def _pair_reduce(self):
        """Every rank's accumulator holds partial sums for all slots: sum them slot by slot onto the owners."""
        if self.world_size > 1:
            dist.reduce_scatter_tensor(self._acc_own, self._acc64, group=self.group)

    def _pair_finish(self, integrator, bodies, bodies_next, do_next):
        s = self._scalars()
        clear, clear_n = (self._acc64, self._acc64.numel()) if self.world_size > 1 else (None, 0)
        _native.call("nbody_shard_pair_finish_f32", integrator, _ptr(bodies), _ptr(bodies_next), self.i_begin, self.n_local,
                     _ptr(self._acc_own), _ptr(clear), clear_n, _ptr(self.positions), _ptr(self.velocities),
                     _ptr(self.accelerations), _ptr(self._vhalf), s["g"], s["dt"], s["half_dt"], do_next,
                     self.world_size, self.n_pad, _ptr(self._pair_ws), self._pair_ws.numel(), self._stream())

    def _force_all_pair(self, integrator, bodies, bodies_next, do_next, gather_work):
        if self._pair_split:
            self._pair_force(0, bodies)  # own triangle while the gather is in flight
            if gather_work is not None:
                gather_work.wait()
            self._pair_force(1, bodies)
        else:
            if gather_work is not None:
                gather_work.wait()
            self._pair_force(0, bodies)
        self._pair_reduce()
        self._pair_finish(integrator, bodies, bodies_next, do_next)

    def _local_energies(self, bodies):
        """This rank's (u, k) partial sums as a device tensor of two doubles."""
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        ws = self._workspace
        _native.call("nbody_shard_energies_f32", _ptr(bodies), _ptr(self.velocities), self.world_size * self.n_pad,
                     self.i_begin, self.n_local, _native.f32(self.g_const), _native.f32(self.softening), _ptr(out),
                     _ptr(ws), ws.numel(), self._stream())
        return out

    def _energies_of(self, bodies):
        out = self._local_energies(bodies)
        dist.all_reduce(out, group=self.group)
        u, k = out.tolist()
        return u, k

    def compute_energies(self):
        """(u_energy, k_energy) of the whole system (simulation.py:91-115), identical on every rank."""
        bodies = self._bodies[0]
        self._prepare(0, bodies)
        self._gather(bodies).wait()
        return self._energies_of(bodies)

    def _gather(self, bodies):
        """In-place all-gather of every rank's slot; returns the async work handle."""
        mine = bodies[self.i_begin : self.i_begin + self.n_pad]
        return dist.all_gather_into_tensor(bodies, mine, group=self.group, async_op=True)

    # ------------------------------------------------------------------ stepping

    def _force_all(self, integrator, bodies, bodies_next, do_next, gather_work=None):
        """With overlap: own part while the gather is in flight, then the rest. Without: wait, then one sweep.
        The last launch runs the epilogue."""
        if self.pair:
            self._force_all_pair(integrator, bodies, bodies_next, do_next, gather_work)
            return
        if len(self._parts) == 1:
            if gather_work is not None:
                gather_work.wait()
            self._force(integrator, bodies, bodies_next, 0, self._parts[0], do_next)
            return
        self._force(integrator, bodies, bodies_next, 0, self._parts[0], do_next)
        if gather_work is not None:
            gather_work.wait()
        for part in range(1, len(self._parts)):
            self._force(integrator, bodies, bodies_next, part, self._parts[part], do_next)

    def _advance(self, steps: int, on_state=None):
        if steps <= 0:
            return
        integ = self._integrator
        cur = 0
        self._prepare(integ, self._bodies[cur])
        for s in range(steps):
            work = self._gather(self._bodies[cur])
            do_next = 1 if (integ == _native.INTEGRATOR_EULER or s + 1 < steps) else 0
            self._force_all(integ, self._bodies[cur], self._bodies[cur ^ 1], do_next, work)
            if on_state is not None:
                # leapfrog: the drift of the NEXT step already went into self.positions; state s sits in the
                # buffer this step consumed. Euler: state s is what the epilogue just wrote.
                src = self._bodies[cur] if integ == _native.INTEGRATOR_LEAPFROG else self._bodies[cur ^ 1]
                on_state(s, src)
            cur ^= 1

    def step(self):
        if self._integrator is None:
            raise NotImplementedError("El método step debe ser implementado en la subclase")
        self._advance(1)

    def run(self, steps: int, record_every: int = 1) -> list[SimulationState]:
        """Like BaseSimulator.run (simulation.py:117-146); the recorded tensors hold this rank's slice only."""
        if self._integrator is None:
            raise NotImplementedError("El método step debe ser implementado en la subclase")
        states = []

        def record(s, bodies):
            if (s + 1) % record_every:
                return
            u = k = None
            if self.calc_energy:
                if self._integrator == _native.INTEGRATOR_EULER:  # the epilogue wrote only this rank's slice
                    self._gather(bodies).wait()
                u, k = self._energies_of(bodies)
            states.append(SimulationState(step=s, step_time=float("nan"),
                                          positions=bodies[self.i_begin : self.i_begin + self.n_local, :3].cpu(),
                                          velocities=self.velocities.cpu(), accelerations=self.accelerations.cpu(),
                                          u_energy=u, k_energy=k))

        self._advance(steps, record)
        return states

    def gather_state(self):
        """Full (positions, velocities, accelerations) on every rank, as CPU tensors in the original body order."""
        out = []
        for t in (self.positions, self.velocities, self.accelerations):
            padded = torch.zeros((self.n_pad, 3), dtype=t.dtype, device=t.device)
            padded[: self.n_local] = t
            full = torch.empty((self.world_size * self.n_pad, 3), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(full, padded, group=self.group)
            rows = [full[r * self.n_pad : r * self.n_pad + c] for r, c in enumerate(self.counts)]
            out.append(torch.cat(rows).cpu())
        return tuple(out)


class ShardedLeapFrogSimulator(ShardedSimulator):
    _integrator = _native.INTEGRATOR_LEAPFROG


class ShardedEulerSimulator(ShardedSimulator):
    _integrator = _native.INTEGRATOR_EULER
