"""ctypes binding of the C ABI in include/nbody_b200.h (libnbody_b200.so).

There is no CPU fallback: if the shared library is missing, or a call fails, this module raises. PyTorch is used
by the callers only to own device memory and streams; the signatures below carry plain pointers and sizes.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libnbody_b200.so")

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_WORKSPACE = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_UNSUPPORTED = -5

INTEGRATOR_LEAPFROG = 1
INTEGRATOR_EULER = 2


class NativeError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, fn: str, status: int, message: str):
        super().__init__(f"{fn} failed with status {status}: {message}")
        self.status = status


# name -> (restype, argtypes); every symbol include/nbody_b200.h declares appears here.
SIGNATURES = {
    "nbody_version": (c_int, []),
    "nbody_status_string": (c_char_p, [c_int]),
    "nbody_last_error": (c_char_p, []),
    "nbody_launch_count": (c_uint64, []),
    "nbody_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nbody_plan_f32": (c_int, [c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "nbody_accel_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "nbody_integrate_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_int,
         c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "nbody_energies_f32": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "nbody_momentum_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "nbody_kick_drift_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]),
    "nbody_kick_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "nbody_shard_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "nbody_shard_prepare_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p],
    ),
    "nbody_shard_force_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
         c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "nbody_shard_energies_f32": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "nbody_pair_min_bodies": (c_int, []),
    "nbody_shard_pair_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nbody_shard_pair_blocks": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int), c_int]),
    "nbody_shard_pair_plan_f32": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "nbody_shard_pair_force_f32": (c_int, [c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nbody_shard_pair_finish_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
         c_float, c_float, c_float, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "nbody_batched_max_n": (c_int, []),
    "nbody_batched_integrate_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_int, c_int,
         c_void_p, c_void_p],
    ),
    "nbody_batched_accel_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p]),
    "nbody_traj_energies_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "nbody_accel_host_f32": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, POINTER(c_uint64), POINTER(c_uint64)],
    ),
    "nbody_integrate_host_f32": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_float, c_int,
         c_int, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_uint64), POINTER(c_uint64)],
    ),
    "nbody_host_cache_release": (c_int, []),
    "nbody_probe_fp32_peak": (c_int, [c_int, c_int, POINTER(c_double)]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Loads libnbody_b200.so once. Raises ImportError with the build recipe if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found. This engine has no CPU or PyTorch fallback: build the sm_100a library with "
                f"`make -C {os.path.join(_PKG_DIR, 'csrc')}` or `python -c 'import __graft_entry__ as g; g.build()'`."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(fn_name: str, status: int) -> None:
    if status == OK:
        return
    msg = lib().nbody_last_error().decode(errors="replace") or lib().nbody_status_string(status).decode()
    if status == ERR_INVALID_ARGUMENT:
        raise ValueError(f"{fn_name}: {msg}")
    raise NativeError(fn_name, status, msg)


def call(fn_name: str, *args) -> None:
    check(fn_name, getattr(lib(), fn_name)(*args))


def f32(x: float) -> float:
    """Rounds a Python double to FP32, as torch does when a Python scalar meets a float32 tensor."""
    return c_float(x).value


def launch_count() -> int:
    return int(lib().nbody_launch_count())
