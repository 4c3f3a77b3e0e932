"""Host-side initial conditions: same API, seeds and draw order as the reference (galaxies.py), checked against the
initial conditions stored in the golden files (which the unmodified reference generated)."""

import json

import numpy as np
import pytest

from galaxify import galaxies


def test_reproduces_reference_initial_conditions(golden):
    ic = golden.meta["ic"]
    gen = galaxies.generate_disk if golden.meta["kind"] == "disk" else galaxies.generate_spiral
    with np.errstate(divide="ignore", invalid="ignore"):
        pos, vel, mass = gen(n_bodies=golden.n, seed=golden.meta["seed"], **ic)
    assert pos.dtype == vel.dtype == mass.dtype == np.float64
    np.testing.assert_array_equal(pos, golden["ic_pos"])
    np.testing.assert_array_equal(mass, golden["ic_mass"])
    # disk speeds use a prefix sum instead of the reference's O(n^2) masked sums: summation order only
    np.testing.assert_allclose(vel, golden["ic_vel"], rtol=1e-12, atol=0)


def test_disk_structure():
    pos, vel, mass = galaxies.generate_disk(n_bodies=2000, total_mass=2.0, radial_scale=3.0, height_scale=0.3,
                                            g_const=1.0, black_hole_mass=0.05, seed=5)
    assert pos.shape == (2000, 3) and vel.shape == (2000, 3) and mass.shape == (2000,)
    assert np.all(pos[0] == 0) and np.all(vel[0] == 0) and mass[0] == pytest.approx(0.1)
    assert mass.sum() == pytest.approx(2.0, rel=1e-12)
    r = np.hypot(pos[1:, 0], pos[1:, 1])
    # circular orbits: velocity perpendicular to the radius vector, no vertical motion
    assert np.abs((pos[1:, :2] * vel[1:, :2]).sum(1)).max() <= 1e-12 * (r * np.linalg.norm(vel[1:, :2], axis=1)).max()
    assert np.all(vel[:, 2] == 0)


def test_disk_offset_rotation_direction():
    kw = dict(n_bodies=200, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=1.0, black_hole_mass=0.01, seed=9)
    p0, v0, m0 = galaxies.generate_disk(**kw)
    p1, v1, m1 = galaxies.generate_disk(offset=(1, 2, 3), initial_vel=(0.1, 0.2, 0.3), clockwise=False, **kw)
    np.testing.assert_allclose(p1, p0 + [1, 2, 3])
    np.testing.assert_allclose(v1[:, :2], -v0[:, :2] + [0.1, 0.2])
    np.testing.assert_array_equal(m0, m1)
    p2, v2, _ = galaxies.generate_disk(angle=(np.pi / 2, 0, 0), **kw)  # about x: y -> z
    np.testing.assert_allclose(p2[:, 2], p0[:, 1], atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(p2, axis=1), np.linalg.norm(p0, axis=1))


def test_hernquist():
    r = np.array([0.0, 1.0, 2.0])
    d = galaxies.spherical_hernquist_distribution(r=r, r0=1, total_mass=1)
    assert d[1] == pytest.approx(1 / (2 * np.pi) / 8) and d[2] == pytest.approx(1 / (2 * np.pi) / 54)
    assert np.isfinite(d[0]) and d[0] > d[1]
    with pytest.raises(ValueError):
        galaxies.spherical_hernquist_distribution(r=r, avoid_distance_zero=False)
    assert galaxies.BodyType.BLACK_HOLE.value == "black hole" and galaxies.BodyType.STAR.value == "star"


def test_spiral_small_and_large_are_consistent():
    kw = dict(total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=4.5e-6, black_hole_mass=0.01, seed=11)
    p, v, m = galaxies.generate_spiral(n_bodies=1, **kw)
    assert p.shape == (1, 3) and m[0] == pytest.approx(0.01)
    p, v, m = galaxies.generate_spiral(n_bodies=3000, **kw)
    assert np.all(m[1:] == m[1]) and m.sum() == pytest.approx(1.0)
    assert np.isfinite(p).all() and np.isfinite(v).all()


def test_plummer_and_merge():
    p, v, m = galaxies.generate_plummer(n_bodies=20000, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=3)
    assert p.shape == (20000, 3) and m.sum() == pytest.approx(1.0)
    # virial ratio 2K/|W| ~ 1 for a Plummer sphere: K = 3 pi/64, W = -3 pi/32 in these units
    k = 0.5 * (m[:, None] * v * v).sum()
    assert k == pytest.approx(3 * np.pi / 64, rel=0.05)
    p2, v2, m2 = galaxies.generate_plummer(n_bodies=20000, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=3)
    np.testing.assert_array_equal(p, p2)
    a = galaxies.generate_disk(n_bodies=10, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=1.0,
                               black_hole_mass=0.01, seed=1)
    b = galaxies.generate_disk(n_bodies=7, total_mass=1.0, radial_scale=3.0, height_scale=0.3, g_const=1.0,
                               black_hole_mass=0.01, seed=2, offset=(10, 0, 0))
    mp, mv, mm = galaxies.merge(a, b)
    assert mp.shape == (17, 3) and mv.shape == (17, 3) and mm.shape == (17,)
    np.testing.assert_array_equal(mp[10:], b[0])
    with pytest.raises(ValueError):
        galaxies.merge()
