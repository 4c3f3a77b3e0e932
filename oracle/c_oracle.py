"""ctypes loader of the plain-C FP64 oracle (nbody_oracle.c). TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_DIR, "liboracle_nbody.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _DIR, "-s"], check=True)
    return _PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        h = ctypes.CDLL(_PATH)
        h.oracle_accel_f64.restype = None
        h.oracle_accel_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        h.oracle_accel_cond_f64.restype = None
        h.oracle_accel_cond_f64.argtypes = h.oracle_accel_f64.argtypes + [ctypes.c_void_p]
        h.oracle_energies_f64.restype = None
        h.oracle_energies_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        _lib = h
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def accelerations_f64(pos, mass, g_const, softening, lo=0, hi=None):
    """FP64 accelerations of rows [lo, hi) from FP32-rounded inputs; eps2 and G rounded to FP32 as the kernels get them."""
    p, m = _f32(pos), _f32(mass)
    n = p.shape[0]
    hi = n if hi is None else hi
    out = np.empty((hi - lo, 3), dtype=np.float64)
    lib().oracle_accel_f64(p.ctypes.data, m.ctypes.data, n, lo, hi, float(np.float32(g_const)),
                           float(np.float32(softening**2)), out.ctypes.data)
    return out


def accelerations_cond_f64(pos, mass, g_const, softening, rows):
    """(accelerations, kappa) of the bodies whose indices are in `rows`: FP64 accelerations and the condition number
    sum_j |term_ij| / |a_i| of each body's sum."""
    p, m = _f32(pos), _f32(mass)
    n = p.shape[0]
    acc = np.empty((len(rows), 3), dtype=np.float64)
    kappa = np.empty(len(rows), dtype=np.float64)
    one, k1 = np.empty(3, dtype=np.float64), np.empty(1, dtype=np.float64)
    for r, i in enumerate(rows):
        lib().oracle_accel_cond_f64(p.ctypes.data, m.ctypes.data, n, int(i), int(i) + 1, float(np.float32(g_const)),
                                    float(np.float32(softening**2)), one.ctypes.data, k1.ctypes.data)
        acc[r], kappa[r] = one, k1[0]
    return acc, kappa


def energies_f64(pos, vel, mass, g_const, softening):
    p, v, m = _f32(pos), _f32(vel), _f32(mass)
    out = np.empty(2, dtype=np.float64)
    lib().oracle_energies_f64(p.ctypes.data, v.ctypes.data, m.ctypes.data, p.shape[0], float(np.float32(g_const)),
                              float(np.float32(softening)), out.ctypes.data)
    return float(out[0]), float(out[1])
