"""galaxify.simulation — same public API as the reference module, computed by sm_100a CUDA kernels.

Mirrors (reference, read-only) src/galaxify/simulation.py: `SimulationState` (:8-18), `BaseSimulator` (:21-150),
`LeapFrogSimulator` (:153-170), `EulerSimulator` (:173-187). Names, keyword arguments, attribute names, return
types, the `ValueError` for a bad `device` string and the `NotImplementedError` of `BaseSimulator.step` are kept
so that src/s01-dataset-generation.py:192-241 runs unchanged on top of this module.

What differs, on purpose:
  * All arithmetic happens in libnbody_b200.so (include/nbody_b200.h). There is no CPU path: `device="cpu"`, or no
    usable B200, raises RuntimeError instead of silently computing somewhere else.
  * `run()` executes every step inside one C call (integrator fused into the force kernel, trajectory written to a
    device buffer, one bulk device->host copy per chunk) instead of a Python loop with three `.cpu()` copies per
    step (simulation.py:126-146). `step_time` is the step's device time from CUDA events; the reference stores an
    unsynchronised wall-clock (simulation.py:127-129); on the persistent small-system path it is the mean over the
    call, because the steps never leave the kernel.
  * Optional, additive keywords that the reference does not have: `run(steps, record_every=1)`.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _native

# Upper bound of one trajectory chunk on the device (and of its pinned host mirror).
TRAJ_CHUNK_BYTES = 2 << 30
# Systems up to this many bodies are stepped by the persistent one-cluster-per-system kernel (csrc/batched.cuh): all
# steps of a run() in ONE launch, energies of the recorded states evaluated afterwards, all states in parallel. Above
# it one fused force/integrator launch per step uses the whole GPU (csrc/force.cuh). The dataset-generation sizes of
# the reference's experiments (3..500 bodies, gnn_experiment.py:34) all take the persistent path.
PERSISTENT_MAX_N = 1024


@dataclass
class SimulationState:
    """State of the simulation after one step (field order as simulation.py:8-18)."""

    step: int
    step_time: float
    positions: torch.Tensor
    velocities: torch.Tensor
    accelerations: torch.Tensor
    u_energy: float = None
    k_energy: float = None


def _as_f32_cuda(x, device: torch.device) -> torch.Tensor:
    """New contiguous FP32 tensor on `device`; never aliases the caller's data (simulation.py:58-65)."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=device, dtype=torch.float32, copy=True).contiguous()
    return torch.tensor(np.asarray(x), dtype=torch.float32, device=device).contiguous()


def _ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


class BaseSimulator:
    _integrator = None  # set by the subclasses

    def __init__(
        self,
        *,
        positions: np.ndarray | torch.Tensor,
        velocities: np.ndarray | torch.Tensor,
        masses: np.ndarray | torch.Tensor,
        g_const: float = 1.0,
        softening: float = 0.1,
        dt: float = 0.01,
        calc_energy: bool = True,
        device: str = None,
    ):
        """Same parameters as the reference constructor (simulation.py:22-69).

        :raises ValueError: if device is not 'cuda', 'cpu' or None (same message as the reference), or if the
            array shapes are not (n,3), (n,3), (n,).
        :raises RuntimeError: if device is 'cpu', or no CUDA device is present: this engine has no CPU path.
        """
        if device is not None and device not in ["cuda", "cpu"]:
            raise ValueError("device debe ser 'cuda', 'cpu' o None")
        if device == "cpu":
            raise RuntimeError(
                "galaxify (B200 engine) has no CPU path: device='cpu' is not available, use device='cuda' or None"
            )
        _native.lib()  # fail loudly before touching the GPU if the extension is missing
        if not torch.cuda.is_available():
            raise RuntimeError("galaxify (B200 engine) needs a CUDA device and found none; there is no CPU fallback")
        self.device = torch.device("cuda")
        self._device_index = torch.cuda.current_device()

        self.dt = dt
        self.g_const = g_const
        self.softening = softening
        self.calc_energy = calc_energy

        self.positions = _as_f32_cuda(positions, self.device)
        self.velocities = _as_f32_cuda(velocities, self.device)
        self.accelerations = None
        self.masses = _as_f32_cuda(masses, self.device)

        self.n = self.positions.shape[0]
        if self.positions.shape != (self.n, 3) or self.velocities.shape != (self.n, 3) or self.masses.shape != (self.n,):
            raise ValueError(
                f"expected positions (n,3), velocities (n,3), masses (n,); got {tuple(self.positions.shape)}, "
                f"{tuple(self.velocities.shape)}, {tuple(self.masses.shape)}"
            )
        self._workspace = None

        self.accelerations = self.compute_accelerations()

    # ------------------------------------------------------------------ native plumbing

    def _scalars(self):
        """FP32 roundings the reference's tensor ops apply to its Python doubles."""
        return dict(
            g=_native.f32(self.g_const),
            eps2=_native.f32(self.softening**2),  # simulation.py:82
            eps=_native.f32(self.softening),  # simulation.py:105
            dt=_native.f32(self.dt),  # simulation.py:166
            half_dt=_native.f32(0.5 * self.dt),  # simulation.py:164
        )

    def _ws(self) -> torch.Tensor:
        need = _native.lib().nbody_workspace_bytes(self.n, self.n)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    @staticmethod
    def _stream() -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _persistent(self) -> bool:
        return 0 < self.n <= PERSISTENT_MAX_N

    def _integrate_persistent(self, steps, record_every, traj, energies, step_ms):
        """All `steps` in one launch of the batched kernel (one system), then the energies of the recorded slots.
        step_ms receives the mean device time of a step (the steps never leave the kernel); like the tiled path and
        the reference (simulation.py:127-129) it covers the integrator only, not the energy evaluation."""
        s = self._scalars()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(self._device_index):
            start.record()
            _native.call("nbody_batched_integrate_f32", self._integrator, _ptr(self.positions), _ptr(self.velocities),
                         _ptr(self.accelerations), _ptr(self.masses), 1, self.n, s["g"], s["eps2"], s["dt"],
                         s["half_dt"], steps, record_every, _ptr(traj), self._stream())
            end.record()
            if energies is not None and traj is not None:
                _native.call("nbody_traj_energies_f32", _ptr(traj), _ptr(self.masses), steps // record_every, 1,
                             self.n, s["g"], s["eps"], _ptr(energies), self._stream())
        if step_ms is not None:
            # like nbody_integrate_f32 with step_ms: everything queued by this call (the energies too) is complete
            # when it returns, which is what run() relies on before it stages the chunk on its copy stream
            torch.cuda.current_stream().synchronize()
            step_ms[:] = start.elapsed_time(end) / max(steps, 1)

    def _integrate(self, steps, record_every, traj, energies, step_ms):
        if self._persistent():
            self._integrate_persistent(steps, record_every, traj, energies, step_ms)
            return
        s = self._scalars()
        ws = self._ws()
        with torch.cuda.device(self._device_index):
            _native.call(
                "nbody_integrate_f32", self._integrator, _ptr(self.positions), _ptr(self.velocities),
                _ptr(self.accelerations), _ptr(self.masses), self.n, s["g"], s["eps2"], s["eps"], s["dt"],
                s["half_dt"], steps, record_every, _ptr(traj), _ptr(energies),
                step_ms.ctypes.data_as(ctypes.c_void_p) if step_ms is not None else None, _ptr(ws), ws.numel(),
                self._stream(),
            )

    # ------------------------------------------------------------------ reference API

    def compute_accelerations(self):
        """a_i = G * sum_{j != i} m_j (r_j - r_i) / (|r_j - r_i|^2 + softening^2)^(3/2)   (simulation.py:71-89).

        :return: new (n_bodies, 3) FP32 CUDA tensor.
        """
        acc = torch.empty_like(self.positions)
        if self.n == 0:
            return acc
        s = self._scalars()
        if self._persistent():
            with torch.cuda.device(self._device_index):
                _native.call("nbody_batched_accel_f32", _ptr(self.positions), _ptr(self.masses), _ptr(acc), 1, self.n,
                             s["g"], s["eps2"], self._stream())
            return acc
        ws = self._ws()
        with torch.cuda.device(self._device_index):
            _native.call("nbody_accel_f32", _ptr(self.positions), _ptr(self.masses), _ptr(acc), self.n, s["g"], s["eps2"],
                         _ptr(ws), ws.numel(), self._stream())
        return acc

    def compute_energies(self):
        """Total potential and kinetic energy, defined as in simulation.py:91-115.

        :return: (u_energy, k_energy) as Python floats.
        """
        if self.n == 0:
            return 0.0, 0.0
        s = self._scalars()
        ws = self._ws()
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self._device_index):
            _native.call("nbody_energies_f32", _ptr(self.positions), _ptr(self.velocities), _ptr(self.masses), self.n,
                         s["g"], s["eps"], _ptr(out), _ptr(ws), ws.numel(), self._stream())
        u, k = out.tolist()
        return u, k

    def compute_momentum(self):
        """(p, l): total linear momentum sum m v and angular momentum sum m (x cross v) as two float64 numpy arrays of
        shape (3,), summed on the device in FP64. An addition (the reference has no momentum function); its drift is
        one of the conservation metrics of SURVEY.md 8a."""
        if self.n == 0:
            return np.zeros(3), np.zeros(3)
        out = torch.empty(6, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self._device_index):
            _native.call("nbody_momentum_f32", _ptr(self.positions), _ptr(self.velocities), _ptr(self.masses), self.n,
                         _ptr(out), self._stream())
        v = out.cpu().numpy()
        return v[:3].copy(), v[3:].copy()

    def run(self, steps: int, record_every: int = 1) -> list[SimulationState]:
        """Runs `steps` steps and returns the recorded states (simulation.py:117-146).

        State k describes the system after step k+1; the initial state is not recorded; `step` is 0-based.
        With record_every > 1 (an addition) only every record_every-th step is recorded.
        The tensors of the returned states are views into one pinned host buffer per chunk (up to TRAJ_CHUNK_BYTES):
        keeping any one of them alive keeps its chunk alive; `.clone()` a state's tensors to hold on to them alone.
        """
        if record_every < 1:
            raise ValueError("record_every must be >= 1")
        states: list[SimulationState] = []
        if steps <= 0:
            return states
        if self._integrator is None:
            self.step()  # raises NotImplementedError, like the reference on the first loop iteration
        if self.n == 0:
            return states
        n = self.n
        chunk_slots = max(1, min(TRAJ_CHUNK_BYTES // (36 * n), 65535))
        chunk_steps = chunk_slots * record_every
        copy_stream = torch.cuda.Stream(device=self.device)
        # fresh tensor, as the reference rebinds self.accelerations every step (simulation.py:168)
        self.accelerations = self.accelerations.clone()

        done = 0
        pending = []
        while done < steps:
            k = min(chunk_steps, steps - done)
            slots = k // record_every
            traj = torch.empty((max(slots, 1), 3, n, 3), dtype=torch.float32, device=self.device)
            energies = (
                torch.empty((max(slots, 1), 2), dtype=torch.float64, device=self.device) if self.calc_energy else None
            )
            step_ms = np.zeros(k, dtype=np.float32)
            self._integrate(k, record_every, traj if slots else None, energies if slots else None, step_ms)
            # the call synchronised (step_ms): stage the chunk to pinned host memory while the next chunk computes
            if slots:
                host = torch.empty((slots, 3, n, 3), dtype=torch.float32, pin_memory=True)
                with torch.cuda.stream(copy_stream):
                    host.copy_(traj[:slots], non_blocking=True)
                    host_en = energies[:slots].to("cpu", non_blocking=False) if energies is not None else None
                traj.record_stream(copy_stream)
                pending.append((done, slots, host, host_en, step_ms))
            done += k
        copy_stream.synchronize()

        for start, slots, host, host_en, step_ms in pending:
            en = host_en.tolist() if host_en is not None else None
            for j in range(slots):
                s_idx = start + (j + 1) * record_every - 1
                states.append(
                    SimulationState(
                        positions=host[j, 0],
                        velocities=host[j, 1],
                        accelerations=host[j, 2],
                        step=s_idx,
                        step_time=float(step_ms[s_idx - start]) * 1e-3,
                        u_energy=en[j][0] if en is not None else None,
                        k_energy=en[j][1] if en is not None else None,
                    )
                )
        return states

    def step(self):
        """Advances the simulation by one step."""
        if self._integrator is None:
            raise NotImplementedError("El método step debe ser implementado en la subclase")
        if self.n == 0:
            return
        self.accelerations = self.accelerations.clone()
        self._integrate(1, 1, None, None, None)


class LeapFrogSimulator(BaseSimulator):
    """Kick-drift-kick leapfrog (simulation.py:153-170); the half-kicks are two separately rounded updates."""

    _integrator = _native.INTEGRATOR_LEAPFROG


class EulerSimulator(BaseSimulator):
    """Semi-implicit Euler: a(x_t), v += dt*a, x += dt*v_new (simulation.py:173-187)."""

    _integrator = _native.INTEGRATOR_EULER
