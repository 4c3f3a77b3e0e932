"""Host-buffer C-ABI entry points (what bench.py's e2e leg and a reference-side binding call)."""

import numpy as np
import pytest

from conftest import load_golden, rel_rows

pytestmark = pytest.mark.gpu


def test_accel_host_matches_reference():
    from galaxify import host

    g = load_golden("disk_n1024_leapfrog")
    acc, h2d, d2h = host.accelerations_host(g["ic_pos"], g["ic_mass"], g_const=g.sim["g_const"],
                                            softening=g.sim["softening"])
    assert rel_rows(acc, g["acc0"]).max() <= 1e-5
    assert (h2d, d2h) == (1024 * 16, 1024 * 12)


@pytest.mark.parametrize("name", ["spiral_n500_leapfrog", "disk_n500_euler"])
def test_integrate_host_matches_reference(name):
    from galaxify import host

    g = load_golden(name)
    pos = np.ascontiguousarray(g["ic_pos"], dtype=np.float32)
    vel = np.ascontiguousarray(g["ic_vel"], dtype=np.float32)
    mass = np.ascontiguousarray(g["ic_mass"], dtype=np.float32)
    acc = np.ascontiguousarray(g["acc0"], dtype=np.float32)
    steps = 100
    r = host.integrate_host(g.integrator, pos, vel, acc, mass, steps=steps, record_every=1, calc_energy=True, **g.sim)
    assert r["traj"].shape == (steps, 3, g.n, 3) and r["energies"].shape == (steps, 2)
    for k, s in enumerate(g.keep):
        if s >= steps:
            continue
        assert np.abs(r["traj"][s, 0] - g["pos"][k]).max() <= 1e-6 * np.abs(g["pos"][k]).max()
        assert np.abs(r["traj"][s, 1] - g["vel"][k]).max() <= 1e-6 * np.abs(g["vel"][k]).max()
        assert rel_rows(r["traj"][s, 2], g["acc"][k]).max() <= 1e-5
    np.testing.assert_array_equal(pos, r["traj"][-1, 0])
    assert np.abs(r["energies"][:, 0] - g["u"][:steps]).max() <= 1e-5 * np.abs(g["u"]).max()
    assert np.abs(r["energies"][:, 1] - g["k"][:steps]).max() <= 1e-5 * np.abs(g["k"]).max()
    assert (r["step_ms"] > 0).all()
    assert r["h2d_bytes"] == g.n * 40


def test_host_api_rejects_bad_arrays():
    from galaxify import host

    with pytest.raises(ValueError):
        host.integrate_host("leapfrog", np.zeros((4, 3)), np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32),
                            np.ones(4, np.float32), steps=1)
    with pytest.raises(KeyError):
        host.integrate_host("rk4", *(np.zeros((4, 3), np.float32),) * 3, np.ones(4, np.float32), steps=1)
