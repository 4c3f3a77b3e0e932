"""The CPU oracle against the reference's golden vectors (tests/golden/, made by importing the unmodified reference).

The oracle restates the reference with the same torch operators, so on the machine that generated the vectors it
is bit-identical; the asserted bound is 2e-6 so that the suite also holds on hosts whose torch CPU kernels
vectorise reductions differently.
"""

import numpy as np
import pytest

from conftest import load_golden, rel_rows
from oracle import c_oracle, galaxify_oracle as oracle

TOL = 2e-6


def test_initial_accelerations(golden):
    acc = oracle.accelerations(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"], golden.sim["softening"])
    assert rel_rows(acc.numpy(), golden["acc0"]).max() <= TOL


def test_row_chunking_does_not_change_results(golden):
    full = oracle.accelerations(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"], golden.sim["softening"])
    chunked = oracle.accelerations(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"],
                                   golden.sim["softening"], chunk=7)
    assert rel_rows(chunked.numpy(), full.numpy()).max() <= TOL
    part = oracle.accelerations(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"], golden.sim["softening"],
                                rows=slice(1, golden.n))
    assert rel_rows(part.numpy(), full.numpy()[1:]).max() <= TOL if golden.n > 1 else part.shape == (0, 3)


@pytest.mark.parametrize("name", ["disk_n500_leapfrog", "spiral_n500_leapfrog", "disk_n500_euler", "spiral_n500_euler",
                                  "disk_n25_defaults_leapfrog", "spiral_n500_defaults_euler",
                                  "spiral_n25_eps0_leapfrog", "disk_n1_leapfrog", "disk_n3_leapfrog"])
def test_trajectories_and_energies(name):
    g = load_golden(name)
    steps = min(g.steps, 100)
    keep = [s for s in g.keep if s < steps]
    out, _ = oracle.run(g["ic_pos"], g["ic_vel"], g["ic_mass"], integrator=g.integrator, steps=steps,
                        calc_energy=True, keep=keep, **g.sim)
    for k, s in enumerate(g.keep):
        if s >= steps:
            continue
        for key in ("pos", "vel"):
            want = g[key][k]
            assert np.abs(out[s][key] - want).max() <= TOL * max(np.abs(want).max(), 1e-30), (name, s, key)
        assert rel_rows(out[s]["acc"], g["acc"][k]).max() <= 10 * TOL, (name, s)
        assert abs(out[s]["u"] - g["u"][s]) <= 1e-5 * abs(g["u"][s]) + 1e-30
        assert abs(out[s]["k"] - g["k"][s]) <= 1e-5 * abs(g["k"][s]) + 1e-30


def test_energy_chunked_matches_unchunked():
    g = load_golden("disk_n1024_leapfrog")
    args = (g["pos"][0], g["vel"][0], g["ic_mass"], g.sim["g_const"], g.sim["softening"])
    u, k = oracle.energies(*args)
    uc, kc = oracle.energies(*args, chunk=100)
    assert abs(u - uc) <= 1e-5 * abs(u) and k == kc
    assert abs(u - float(g["u"][0])) <= 1e-5 * abs(u)


def test_fp64_oracles_agree_with_reference_to_fp32_noise(golden):
    """The FP64 restatements (numpy and C) sit within the reference's own FP32 noise (3-4e-7)."""
    if golden.sim["softening"] == 0.0 and golden.n > 1:
        tol = 5e-6
    else:
        tol = 2e-6
    a_np = oracle.accelerations_f64(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"], golden.sim["softening"])
    a_c = c_oracle.accelerations_f64(golden["ic_pos"], golden["ic_mass"], golden.sim["g_const"],
                                     golden.sim["softening"])
    assert rel_rows(a_np, a_c).max() <= 1e-12
    assert rel_rows(golden["acc0"], a_c).max() <= tol
    u, k = c_oracle.energies_f64(golden["ic_pos"], golden["ic_vel"], golden["ic_mass"], golden.sim["g_const"],
                                 golden.sim["softening"])
    if golden.n > 1 and golden.sim["softening"] > 0:
        assert abs(u - float(golden["u0"])) <= 1e-5 * abs(u)
        assert abs(k - float(golden["k0"])) <= 1e-5 * abs(k)


def test_c_oracle_row_range():
    g = load_golden("spiral_n500_leapfrog")
    full = c_oracle.accelerations_f64(g["ic_pos"], g["ic_mass"], 4.5e-6, 0.05)
    part = c_oracle.accelerations_f64(g["ic_pos"], g["ic_mass"], 4.5e-6, 0.05, 100, 217)
    np.testing.assert_array_equal(part, full[100:217])


def _random_system(seed, n):
    rng = np.random.default_rng(seed)
    pos = rng.standard_normal((n, 3)) * rng.choice([0.1, 1.0, 10.0])
    vel = rng.standard_normal((n, 3)) * 1e-2
    mass = rng.random(n) ** 3 + 1e-6
    return pos, vel, mass


@pytest.mark.parametrize("seed", range(8))
def test_fp32_restatement_tracks_fp64_on_random_clouds(seed):
    """Beyond the golden vectors: on random clouds the FP32 torch restatement stays within the FP32 conditioning
    bound of the FP64 C oracle, for accelerations (2e-7 * kappa, floor 2e-6) and for both energies."""
    n = [2, 7, 50, 200, 333, 1000, 1500, 64][seed]
    softening = [0.01, 0.1][seed % 2]
    pos, vel, mass = _random_system(seed, n)
    a32 = oracle.accelerations(pos, mass, 0.7, softening).numpy()
    a64, kappa = c_oracle.accelerations_cond_f64(pos, mass, 0.7, softening, np.arange(n))
    err = rel_rows(a32, a64)
    assert np.all(err <= np.maximum(2e-6, 2e-7 * kappa)), (err.max(), kappa.max())
    u32, k32 = oracle.energies(pos, vel, mass, 0.7, softening)
    u64, k64 = c_oracle.energies_f64(pos, vel, mass, 0.7, softening)
    assert abs(u32 - u64) <= 1e-5 * abs(u64) and abs(k32 - k64) <= 1e-5 * abs(k64)


def test_condition_number_oracle_matches_plain_accelerations():
    pos, vel, mass = _random_system(3, 120)
    rows = np.array([0, 5, 119])
    acc, kappa = c_oracle.accelerations_cond_f64(pos, mass, 1.0, 0.05, rows)
    full = c_oracle.accelerations_f64(pos, mass, 1.0, 0.05)
    np.testing.assert_allclose(acc, full[rows], rtol=1e-14)
    assert np.all(kappa >= 1.0 - 1e-12)  # a 1-norm over a norm of the sum is never below one
