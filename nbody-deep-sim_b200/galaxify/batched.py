"""Many independent small systems stepped together (dataset generation, BASELINE.json configs[2]).

An addition with no reference counterpart: src/s01-dataset-generation.py:130-214 builds and runs its scenes one at a
time. The constructors keep the reference's keyword arguments; arrays carry a leading system dimension:
positions / velocities (S, n, 3), masses (S, n). Every system evolves exactly as LeapFrogSimulator / EulerSimulator
would evolve it alone, up to FP32 summation order. All steps of a `run` execute inside ONE persistent kernel
(one CTA per system, bodies in shared memory), see csrc/batched.cuh.

Multi-GPU: shard by system index (`shard_systems`) and give each rank its own slice: there is no communication.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native
from .simulation import SimulationState, _as_f32_cuda, _ptr


def shard_systems(n_systems: int, rank: int, world_size: int) -> slice:
    """Contiguous, balanced slice of system indices for `rank` (first ranks take the remainder)."""
    base, extra = divmod(n_systems, world_size)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


class BatchedSimulator:
    _integrator = None

    def __init__(self, *, positions, velocities, masses, g_const: float = 1.0, softening: float = 0.1,
                 dt: float = 0.01, calc_energy: bool = False, device: str = None):
        if device is not None and device not in ["cuda", "cpu"]:
            raise ValueError("device debe ser 'cuda', 'cpu' o None")
        if device == "cpu":
            raise RuntimeError("galaxify (B200 engine) has no CPU path")
        _native.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("galaxify (B200 engine) needs a CUDA device and found none; there is no CPU fallback")
        self.device = torch.device("cuda")
        self.dt, self.g_const, self.softening, self.calc_energy = dt, g_const, softening, calc_energy
        self.positions = _as_f32_cuda(positions, self.device)
        self.velocities = _as_f32_cuda(velocities, self.device)
        self.masses = _as_f32_cuda(masses, self.device)
        if self.positions.dim() != 3 or self.positions.shape[2] != 3:
            raise ValueError("positions must have shape (n_systems, n, 3)")
        self.n_systems, self.n = self.positions.shape[0], self.positions.shape[1]
        if self.velocities.shape != self.positions.shape or self.masses.shape != (self.n_systems, self.n):
            raise ValueError("velocities must match positions and masses must have shape (n_systems, n)")
        if self.n > _native.lib().nbody_batched_max_n():
            raise ValueError(f"batched systems are limited to {_native.lib().nbody_batched_max_n()} bodies")
        self.accelerations = self.compute_accelerations()

    def _scalars(self):
        return (_native.f32(self.g_const), _native.f32(self.softening**2), _native.f32(self.dt),
                _native.f32(0.5 * self.dt))

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def compute_accelerations(self):
        acc = torch.empty_like(self.positions)
        g, eps2, _, _ = self._scalars()
        _native.call("nbody_batched_accel_f32", _ptr(self.positions), _ptr(self.masses), _ptr(acc), self.n_systems,
                     self.n, g, eps2, self._stream())
        return acc

    def _integrate(self, steps, record_every, traj):
        g, eps2, dt, half_dt = self._scalars()
        _native.call("nbody_batched_integrate_f32", self._integrator, _ptr(self.positions), _ptr(self.velocities),
                     _ptr(self.accelerations), _ptr(self.masses), self.n_systems, self.n, g, eps2, dt, half_dt, steps,
                     record_every, _ptr(traj), self._stream())

    def step(self):
        if self._integrator is None:
            raise NotImplementedError("El método step debe ser implementado en la subclase")
        self.accelerations = self.accelerations.clone()
        self._integrate(1, 1, None)

    def run(self, steps: int, record_every: int = 1, to_host: bool = True):
        """Returns the recorded states; tensors have shape (n_systems, n, 3). With to_host=False they stay on the
        GPU (views of one trajectory buffer) for a consumer that writes them out itself. With calc_energy=True,
        u_energy / k_energy are numpy arrays of shape (n_systems,)."""
        if self._integrator is None:
            raise NotImplementedError("El método step debe ser implementado en la subclase")
        if record_every < 1:
            raise ValueError("record_every must be >= 1")
        slots = steps // record_every
        if steps <= 0:
            return []
        traj = None
        if slots:
            traj = torch.empty((slots, 3, self.n_systems, self.n, 3), dtype=torch.float32, device=self.device)
        self.accelerations = self.accelerations.clone()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        self._integrate(steps, record_every, traj)
        end.record()
        if not slots:
            return []
        energies = None
        if self.calc_energy:
            energies = torch.empty((slots, self.n_systems, 2), dtype=torch.float64, device=self.device)
            for lo in range(0, slots, 65535):  # grid.y limit of the energy kernel
                cnt = min(65535, slots - lo)
                _native.call("nbody_traj_energies_f32", _ptr(traj[lo]), _ptr(self.masses), cnt, self.n_systems, self.n,
                             _native.f32(self.g_const), _native.f32(self.softening), _ptr(energies[lo]), self._stream())
            energies = energies.cpu().numpy()
        if to_host:
            host = torch.empty(traj.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(traj, non_blocking=True)
            torch.cuda.synchronize()
            traj = host
        else:
            end.synchronize()
        per_step = start.elapsed_time(end) * 1e-3 / steps
        return [SimulationState(step=(j + 1) * record_every - 1, step_time=per_step, positions=traj[j, 0],
                                velocities=traj[j, 1], accelerations=traj[j, 2],
                                u_energy=energies[j, :, 0] if energies is not None else None,
                                k_energy=energies[j, :, 1] if energies is not None else None) for j in range(slots)]


class BatchedLeapFrogSimulator(BatchedSimulator):
    _integrator = _native.INTEGRATOR_LEAPFROG


class BatchedEulerSimulator(BatchedSimulator):
    _integrator = _native.INTEGRATOR_EULER
