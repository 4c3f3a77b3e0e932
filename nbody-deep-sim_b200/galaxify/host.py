"""Host-buffer entry points: numpy in, numpy out, host<->device copies inside the call.

These wrap nbody_accel_host_f32 / nbody_integrate_host_f32 (include/nbody_b200.h), the form of the C ABI a
reference-side binding would call with the arrays `galaxify.galaxies` returns. They are what
bench.py times for the end-to-end number. No torch involved.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _native


def _f32c(a, shape):
    out = np.ascontiguousarray(a, dtype=np.float32)
    if out.shape != shape:
        raise ValueError(f"expected shape {shape}, got {out.shape}")
    return out


def accelerations_host(positions, masses, *, g_const: float = 1.0, softening: float = 0.1, device: int = 0):
    """compute_accelerations (simulation.py:71-89) on host arrays. Returns ((n,3) float32, h2d_bytes, d2h_bytes)."""
    n = len(masses)
    pos, mass = _f32c(positions, (n, 3)), _f32c(masses, (n,))
    acc = np.empty((n, 3), dtype=np.float32)
    h2d, d2h = ctypes.c_uint64(), ctypes.c_uint64()
    _native.call("nbody_accel_host_f32", pos.ctypes.data, mass.ctypes.data, acc.ctypes.data, n,
                 _native.f32(g_const), _native.f32(softening**2), device, ctypes.byref(h2d), ctypes.byref(d2h))
    return acc, h2d.value, d2h.value


def integrate_host(integrator: str, positions, velocities, accelerations, masses, *, steps: int, g_const: float = 1.0,
                   softening: float = 0.1, dt: float = 0.01, record_every: int = 0, calc_energy: bool = False,
                   device: int = 0):
    """`steps` leapfrog / Euler steps (simulation.py:153-187) on host arrays, updated IN PLACE (float32, C-order).

    Returns dict(traj=(slots,3,n,3) or None, energies=(slots,2) or None, step_ms, h2d_bytes, d2h_bytes).
    """
    code = {"leapfrog": _native.INTEGRATOR_LEAPFROG, "euler": _native.INTEGRATOR_EULER}[integrator]
    n = len(masses)
    for name, a, shape in (("positions", positions, (n, 3)), ("velocities", velocities, (n, 3)),
                           ("accelerations", accelerations, (n, 3)), ("masses", masses, (n,))):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags.c_contiguous and a.shape == shape):
            raise ValueError(f"{name} must be a C-contiguous float32 array of shape {shape}")
    slots = steps // record_every if record_every > 0 else 0
    traj = np.empty((slots, 3, n, 3), dtype=np.float32) if slots else None
    energies = np.empty((slots, 2), dtype=np.float64) if (slots and calc_energy) else None
    step_ms = np.zeros(steps, dtype=np.float32)
    h2d, d2h = ctypes.c_uint64(), ctypes.c_uint64()
    _native.call("nbody_integrate_host_f32", code, positions.ctypes.data, velocities.ctypes.data,
                 accelerations.ctypes.data, masses.ctypes.data, n, _native.f32(g_const), _native.f32(softening**2),
                 _native.f32(softening), _native.f32(dt), _native.f32(0.5 * dt), steps, max(record_every, 1),
                 traj.ctypes.data if traj is not None else None,
                 energies.ctypes.data if energies is not None else None, step_ms.ctypes.data, device,
                 ctypes.byref(h2d), ctypes.byref(d2h))
    return dict(traj=traj, energies=energies, step_ms=step_ms, h2d_bytes=h2d.value, d2h_bytes=d2h.value)
