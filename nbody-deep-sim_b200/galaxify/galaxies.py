"""galaxify.galaxies — initial conditions with the reference's API, seeds and random-draw order.

Mirrors (reference, read-only) src/galaxify/galaxies.py: `BodyType` (:6-8), `spherical_hernquist_distribution`
(:11-51), `generate_disk` (:54-192), `generate_spiral` (:195-296). Host-side numpy code, run once per scene; the
kernels never see it. Every generator returns (positions (n,3), velocities (n,3), masses (n,)) as float64.

The legacy global numpy RNG is consumed in exactly the reference's order, so a seeded call gives the same galaxy.
What is restated differently is the arithmetic around the draws:
  * generate_disk: the enclosed mass of each star, an O(n^2) Python loop in the reference (galaxies.py:143-152), is
    a sort + exclusive prefix sum here (O(n log n)); it differs from the reference only in summation order (~1e-15).
  * generate_spiral: the reference draws and computes body by body (galaxies.py:245-294); here only the draws stay
    in a loop (their order interleaves distributions and cannot be batched), the geometry is vectorised.

Additions (no reference counterpart): `generate_plummer`, `merge`.
"""

from __future__ import annotations

import enum

import numpy as np


class BodyType(enum.Enum):
    BLACK_HOLE = "black hole"
    STAR = "star"


def spherical_hernquist_distribution(
    *,
    r: float | np.ndarray,
    r0: float = 1,
    total_mass: float = 1,
    avoid_distance_zero: bool = True,
) -> float | np.ndarray:
    """Hernquist density rho(r) = M/(2 pi) * r0 / (r (r0 + r)^3)   (galaxies.py:11-51).

    :param avoid_distance_zero: replace r == 0 by float32 eps instead of raising.
    :raises ValueError: if r contains a zero and avoid_distance_zero is False.
    """
    radius = np.asarray(r)
    zero = radius == 0
    if avoid_distance_zero:
        radius = np.where(zero, np.finfo(np.float32).eps, radius)
    elif np.any(zero):
        raise ValueError("r contiene cero(s) y avoid_distance_zero es False")
    return (total_mass / (2 * np.pi)) * (r0 / (radius * (r0 + radius) ** 3))


def _axis_rotations(angle) -> list[np.ndarray]:
    """Rotation matrices about x, y, z for the Euler angles `angle` (galaxies.py:159-181)."""
    ax, ay, az = (float(a) for a in np.array(angle))
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    return [
        np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]),
        np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]),
        np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]),
    ]


def _enclosed_mass(distances: np.ndarray, masses: np.ndarray) -> np.ndarray:
    """For every body, the total mass of the bodies strictly closer to the centre (galaxies.py:146)."""
    order = np.argsort(distances, kind="stable")
    sorted_d = distances[order]
    below = np.concatenate(([0.0], np.cumsum(masses[order])))
    return below[np.searchsorted(sorted_d, distances, side="left")]


def generate_disk(
    *,
    n_bodies: int,
    total_mass: float,
    radial_scale: float,
    height_scale: float,
    g_const: float,
    black_hole_mass: float,
    offset=(0, 0, 0),
    initial_vel=(0, 0, 0),
    clockwise=True,
    angle=(0, 0, 0),
    seed: int = None,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exponential disk with a central black hole carrying `black_hole_mass` of the total mass (galaxies.py:54-192).

    Body 0 is the black hole at the centre, at rest. Star masses are Hernquist-weighted and normalised so that all
    masses sum to `total_mass`; stars move on circular orbits around the mass enclosed by their radius.
    """
    np.random.seed(seed)
    n = n_bodies
    is_star = np.ones(n, dtype=bool)
    is_star[0] = False

    # three draws, in this order (galaxies.py:99-114)
    u_radius = np.random.uniform(low=np.finfo(np.float32).eps, high=1.0, size=n)
    u_height = np.random.uniform(-1.0, 1.0, size=n)
    phi = np.random.rand(n) * 2 * np.pi

    distances = -radial_scale * np.log(1 - u_radius)
    distances[~is_star] = 0
    zs = u_height * height_scale * (1 - np.sqrt(distances))
    zs[~is_star] = 0
    positions = np.array((np.cos(phi) * distances, np.sin(phi) * distances, zs)).T

    mass_bh = total_mass * black_hole_mass
    masses = np.empty(n)
    masses[0] = mass_bh
    weights = spherical_hernquist_distribution(r=distances[is_star], r0=1, total_mass=total_mass)
    masses[is_star] = weights * ((total_mass - mass_bh) / weights.sum())

    velocities = np.zeros((n, 3))
    stars = np.flatnonzero(is_star)
    if stars.size:
        speed = np.sqrt(g_const * _enclosed_mass(distances, masses)[stars] / distances[stars])
        velocities[stars, 0] = speed * np.cos(phi[stars] + np.pi / 2)
        velocities[stars, 1] = speed * np.sin(phi[stars] + np.pi / 2)
    if clockwise:
        velocities[:, 0] = -velocities[:, 0]
        velocities[:, 1] = -velocities[:, 1]

    for rot in _axis_rotations(angle):
        positions = positions @ rot.T
        velocities = velocities @ rot.T
    positions += np.array(offset)
    velocities += np.array(initial_vel)
    return positions, velocities, masses


def generate_spiral(
    *,
    n_bodies: int,
    total_mass: float,
    radial_scale: float,
    height_scale: float,
    g_const: float,
    black_hole_mass: float,
    n_arms: int = 2,
    pitch_angle: float = -np.pi / 6,
    arm_strength: float = 0.3,
    seed: int = None,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Spiral galaxy with a central black hole and equal-mass stars (galaxies.py:195-296)."""
    np.random.seed(seed)
    n = n_bodies
    positions = np.zeros((n, 3))
    velocities = np.zeros((n, 3))
    mass_bh = total_mass * black_hole_mass
    masses = np.empty(n)
    masses[0] = mass_bh
    if n > 1:
        masses[1:] = (total_mass - mass_bh) / (n - 1)
    if n <= 1:
        return positions, velocities, masses

    # Per star, in this order: gamma radius, uniform azimuth, then four normals (height, v_R, v_phi, v_z). The normals
    # are drawn with unit scale: the legacy generator computes loc + scale * gauss, so scaling afterwards is identical.
    ns = n - 1
    r = np.empty(ns)
    u = np.empty(ns)
    g = np.empty((ns, 4))
    gamma, rand, normal = np.random.gamma, np.random.rand, np.random.normal
    for i in range(ns):
        r[i] = gamma(shape=2, scale=radial_scale)
        u[i] = rand()
        g[i, 0] = normal(0, 1.0)
        g[i, 1] = normal(0, 1.0)
        g[i, 2] = normal(0, 1.0)
        g[i, 3] = normal(0, 1.0)

    phi = 2 * np.pi * u
    safe_r = np.where(r > 0, r, 1.0)
    swirl = phi + arm_strength * np.sin(n_arms * (phi - np.log(safe_r / radial_scale) / np.tan(pitch_angle)))
    phi_spiral = np.where(r > 0, swirl, phi)
    cos_p, sin_p = np.cos(phi_spiral), np.sin(phi_spiral)
    positions[1:, 0] = r * cos_p
    positions[1:, 1] = r * sin_p
    positions[1:, 2] = 0 + height_scale * g[:, 0]

    m_enc = total_mass * (1 - np.exp(-r / radial_scale) * (1 + r / radial_scale))
    v_circ = np.where(r < 1e-8, 0.0, np.sqrt(g_const * m_enc / safe_r))
    v_r = 0 + (0.1 * v_circ) * g[:, 1]
    v_phi = v_circ + (0 + (0.07 * v_circ) * g[:, 2])
    v_z = 0 + (0.05 * v_circ) * g[:, 3]
    velocities[1:, 0] = v_r * cos_p - v_phi * sin_p
    velocities[1:, 1] = v_r * sin_p + v_phi * cos_p
    velocities[1:, 2] = v_z
    return positions, velocities, masses


def generate_plummer(
    *,
    n_bodies: int,
    total_mass: float,
    scale_radius: float,
    g_const: float,
    seed: int = None,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Equal-mass Plummer sphere in virial equilibrium (Aarseth, Henon & Wielen 1974 sampling).

    New helper with no reference counterpart: benchmark configurations ask for a Plummer model and the reference
    has none. Uses its own `np.random.default_rng(seed)`, so it does not disturb the global stream.
    """
    rng = np.random.default_rng(seed)
    n = n_bodies
    masses = np.full(n, total_mass / n)
    radius = scale_radius / np.sqrt(np.maximum(rng.uniform(1e-10, 1.0, n) ** (-2.0 / 3.0) - 1.0, 1e-12))

    def isotropic(count):
        cos_t = rng.uniform(-1.0, 1.0, count)
        sin_t = np.sqrt(1.0 - cos_t * cos_t)
        az = rng.uniform(0.0, 2 * np.pi, count)
        return np.stack((sin_t * np.cos(az), sin_t * np.sin(az), cos_t), axis=1)

    positions = radius[:, None] * isotropic(n)
    # speed fraction q of the escape speed, pdf ~ q^2 (1 - q^2)^(7/2), by rejection
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.uniform(0.0, 1.0, todo.size)
        y = rng.uniform(0.0, 0.1, todo.size)
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    v_escape = np.sqrt(2.0 * g_const * total_mass) * (radius * radius + scale_radius * scale_radius) ** -0.25
    velocities = (q * v_escape)[:, None] * isotropic(n)
    return positions, velocities, masses


def merge(*galaxies) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Concatenates (positions, velocities, masses) triples into one system, e.g. two `generate_disk` galaxies
    placed with `offset` / `initial_vel` / `angle` for a merger. New helper with no reference counterpart."""
    if not galaxies:
        raise ValueError("merge needs at least one galaxy")
    return (
        np.concatenate([g[0] for g in galaxies], axis=0),
        np.concatenate([g[1] for g in galaxies], axis=0),
        np.concatenate([g[2] for g in galaxies], axis=0),
    )
