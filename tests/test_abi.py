"""The C-ABI shared library: loads, exports every symbol include/nbody_b200.h declares, argument checks and the
pure-host helpers behave. No compute call is made here (no GPU in the CPU suite)."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from galaxify import _native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbody_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    names = declared_symbols()
    assert len(names) >= 29
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_native.SIGNATURES) == names  # the binding covers exactly the header


def test_version_and_status_strings():
    lib = _native.lib()
    assert lib.nbody_version() == 101
    assert lib.nbody_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert lib.nbody_status_string(code) not in (b"ok", b"unknown status")
    assert lib.nbody_status_string(-99) == b"unknown status"


def test_workspace_bytes_is_pure_and_monotone():
    lib = _native.lib()
    assert lib.nbody_workspace_bytes(0, 0) == 0 and lib.nbody_workspace_bytes(5, 3) == 0
    sizes = [lib.nbody_workspace_bytes(n, n) for n in (1, 3, 1000, 16384, 262144, 1 << 20)]
    assert all(s > 0 for s in sizes)
    assert sizes[-1] >= 2 * 16 * (1 << 20)  # two (x,y,z,m) arrays at least
    assert sizes[-1] < 256 << 20
    assert lib.nbody_shard_workspace_bytes(1 << 17, 1 << 20, 8) > 0
    assert lib.nbody_shard_workspace_bytes(1 << 17, 1 << 20, 0) == 0
    assert lib.nbody_batched_max_n() == 2048


def test_launch_plan_fills_the_machine():
    """Plans are pure functions of the sizes; every plan's CTA count should waste little of its last wave."""
    lib = _native.lib()
    for n, j in ((1 << 20, 1 << 20), (262144, 262144), (131072, 349525), (65536, 65536), (16384, 16384), (1000, 1000)):
        large, tiles, splits = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.nbody_plan_f32(n, j, ctypes.byref(large), ctypes.byref(tiles), ctypes.byref(splits)) == 0
        tile_i = 2048 if large.value else 512
        slots = 148 * (1 if large.value else 2)
        assert tiles.value == -(-n // tile_i) and 1 <= splits.value <= 16
        ctas = tiles.value * splits.value
        if n >= 16384:
            assert -(-ctas // slots) * slots / ctas <= 1.16, (n, large.value, tiles.value, splits.value)
    # the headline and the 8-GPU shard both take the 4-bodies-per-thread shape
    for n, j in ((1 << 20, 1 << 20), (131072, 349525)):
        assert lib.nbody_plan_f32(n, j, ctypes.byref(large), ctypes.byref(tiles), ctypes.byref(splits)) == 0
        assert large.value == 1
    assert lib.nbody_plan_f32(0, 5, ctypes.byref(large), ctypes.byref(tiles), ctypes.byref(splits)) == -1


def test_argument_errors_are_reported_without_touching_a_device():
    lib = _native.lib()
    buf = np.zeros(64, dtype=np.float32)
    p = buf.ctypes.data
    assert lib.nbody_accel_f32(None, p, p, 4, 1.0, 0.01, p, 1 << 20, None) == _native.ERR_INVALID_ARGUMENT
    assert b"null" in lib.nbody_last_error()
    assert lib.nbody_accel_f32(p, p, p, 0, 1.0, 0.01, p, 1 << 20, None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_integrate_f32(7, p, p, p, p, 4, 1.0, 0.01, 0.1, 0.01, 0.005, 1, 1, None, None, None, p, 1 << 20,
                                   None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_integrate_f32(1, p, p, p, p, 4, 1.0, 0.01, 0.1, 0.01, 0.005, -1, 1, None, None, None, p, 1 << 20,
                                   None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_batched_integrate_f32(1, p, p, p, p, 2, 4096, 1.0, 0.01, 0.01, 0.005, 1, 1, None,
                                           None) == _native.ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        _native.call("nbody_accel_f32", None, None, None, 4, 1.0, 0.01, None, 0, None)


def test_f32_rounding_helper_matches_numpy():
    for x in (0.5 * 1e-4, 0.05**2, 4.5e-6, 0.1, 1e-40, 0.0):
        assert _native.f32(x) == float(np.float32(x))


def test_simulator_api_without_a_gpu():
    """Same names and error behaviour as simulation.py:22-51,148-150; and a loud failure instead of a CPU fallback."""
    import dataclasses
    import inspect

    import torch

    from galaxify import simulation

    fields = [f.name for f in dataclasses.fields(simulation.SimulationState)]
    assert fields == ["step", "step_time", "positions", "velocities", "accelerations", "u_energy", "k_energy"]
    sig = inspect.signature(simulation.BaseSimulator.__init__)
    assert [p for p in sig.parameters][1:] == ["positions", "velocities", "masses", "g_const", "softening", "dt",
                                               "calc_energy", "device"]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for n, p in sig.parameters.items() if n != "self")
    d = {n: p.default for n, p in sig.parameters.items()}
    assert (d["g_const"], d["softening"], d["dt"], d["calc_energy"], d["device"]) == (1.0, 0.1, 0.01, True, None)
    for cls in (simulation.LeapFrogSimulator, simulation.EulerSimulator):
        assert issubclass(cls, simulation.BaseSimulator)
    kw = dict(positions=np.zeros((2, 3)), velocities=np.zeros((2, 3)), masses=np.ones(2))
    with pytest.raises(ValueError, match="device debe ser 'cuda', 'cpu' o None"):
        simulation.LeapFrogSimulator(device="cuda:0", **kw)
    with pytest.raises(RuntimeError, match="no CPU path"):
        simulation.LeapFrogSimulator(device="cpu", **kw)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            simulation.LeapFrogSimulator(**kw)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "libnbody_b200.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _native.lib()


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under nbody-deep-sim_b200/ may import, load or execute it."""
    import glob

    from conftest import PKG

    for path in glob.glob(os.path.join(PKG, "**", "*"), recursive=True):
        if os.path.isdir(path) or not path.endswith((".py", ".cu", ".cuh", ".h")):
            continue
        text = open(path).read()
        assert "oracle" not in text.lower(), path


def test_binding_argument_counts_match_the_header():
    """Every ctypes signature has as many arguments as the C declaration (a mismatch corrupts the stack silently)."""
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict(re.findall(r"\b(nbody_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(decls) == set(_native.SIGNATURES)
    for name, params in decls.items():
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_native.SIGNATURES[name][1]), (name, n, len(_native.SIGNATURES[name][1]))
    # and every slot has the right kind: pointer, float, size_t or int
    for name, params in decls.items():
        params = " ".join(params.split())
        if params in ("", "void"):
            continue
        for ctype, decl in zip(_native.SIGNATURES[name][1], params.split(",")):
            if "*" in decl:
                assert ctype is ctypes.c_void_p or issubclass(ctype, ctypes._Pointer), (name, decl)
            elif "float" in decl:
                assert ctype is ctypes.c_float, (name, decl)
            elif "size_t" in decl:
                assert ctype is ctypes.c_size_t, (name, decl)
            elif "long long" in decl:
                assert ctype is ctypes.c_longlong, (name, decl)
            else:
                assert ctype is ctypes.c_int, (name, decl)
