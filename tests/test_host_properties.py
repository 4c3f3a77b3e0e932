"""Property tests (hypothesis) of the pure host logic: launch plans, workspace sizing, shard layout, system sharding."""

import ctypes

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from galaxify import _native
from galaxify.batched import shard_systems
from galaxify.sharded import shard_layout, step_parts


@settings(max_examples=300, deadline=None)
@given(n=st.integers(1, 1 << 22), j=st.integers(1, 1 << 22))
def test_launch_plan_is_always_valid(n, j):
    lib = _native.lib()
    large, tiles, splits = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.nbody_plan_f32(n, j, ctypes.byref(large), ctypes.byref(tiles), ctypes.byref(splits)) == 0
    tile_i = 2048 if large.value else 512
    tile_j = 1024 if large.value else 512
    assert tiles.value == -(-n // tile_i) >= 1
    assert 1 <= splits.value <= 16
    assert splits.value == 1 or j // splits.value >= 2 * tile_j  # every split keeps at least two j tiles


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 1 << 21), extra=st.integers(0, 1 << 21), parts=st.integers(1, 4))
def test_workspace_sizes_are_consistent(n, extra, parts):
    lib = _native.lib()
    total = n + extra
    single = lib.nbody_workspace_bytes(n, total)
    shard = lib.nbody_shard_workspace_bytes(n, total, parts)
    assert single >= 2 * 16 * total + 12 * n  # two body arrays and the half-kick velocities live in it
    assert 0 < shard < single + (1 << 30)
    assert single % 256 == 0 and shard % 256 == 0
    # more i-bodies never need less scratch
    if n > 1:
        assert lib.nbody_workspace_bytes(n - 1, total) <= single + 64 * 1024 * 1024


@settings(max_examples=300, deadline=None)
@given(n=st.integers(1, 5_000_000), world=st.integers(1, 16), overlap=st.booleans())
def test_shard_layout_and_parts(n, world, overlap):
    n_pad, counts = shard_layout(n, world)
    assert sum(counts) == n and len(counts) == world and max(counts) == n_pad
    assert all(c >= 0 for c in counts) and counts == sorted(counts, reverse=True)
    total = world * n_pad
    for rank in {0, world // 2, world - 1}:
        if counts[rank] == 0:
            continue
        parts = step_parts(rank, n_pad, counts, overlap)
        covered = []
        for ranges in parts:
            for lo, hi in ranges:
                assert 0 <= lo <= hi <= total
                if hi > lo:
                    covered.append((lo, hi))
        covered.sort()
        for (a0, a1), (b0, b1) in zip(covered, covered[1:]):
            assert a1 <= b0  # disjoint
        real = sum(min(hi, r * n_pad + c) - max(lo, r * n_pad) for lo, hi in covered for r, c in enumerate(counts)
                   if min(hi, r * n_pad + c) > max(lo, r * n_pad))
        assert real == n  # every real body exactly once


@settings(max_examples=300, deadline=None)
@given(total=st.integers(0, 100_000), world=st.integers(1, 64))
def test_shard_systems_is_a_balanced_partition(total, world):
    sizes, nxt = [], 0
    for r in range(world):
        sl = shard_systems(total, r, world)
        assert sl.start == nxt and sl.stop >= sl.start
        sizes.append(sl.stop - sl.start)
        nxt = sl.stop
    assert nxt == total and max(sizes) - min(sizes) <= 1


@settings(max_examples=150, deadline=None)
@given(n=st.integers(1, 1 << 21), world=st.integers(1, 16))
def test_pair_blocks_partition_the_work(n, world):
    """The pair path's blocks (nbody_shard_pair_blocks, a pure host function): over all ranks they contain exactly
    n(n-1)/2 unordered pairs, stay inside the real bodies, and the sharded pair workspace is sized for them."""
    if world > n:
        return
    lib = _native.lib()
    n_pad, counts = shard_layout(n, world)
    pairs = 0
    buf = (ctypes.c_int * (5 * 16))()
    for rank in range(world):
        k = lib.nbody_shard_pair_blocks(n, world, n_pad, rank, buf, 16)
        assert 1 <= k <= 16
        lo = rank * n_pad
        assert (buf[0], buf[1], buf[2], buf[3], buf[4]) == (lo, lo + counts[rank], lo, lo + counts[rank], 1)
        pairs += counts[rank] * (counts[rank] - 1) // 2
        for b in range(1, k):
            i_lo, i_hi, j_lo, j_hi, tri = (buf[5 * b + c] for c in range(5))
            assert tri == 0 and lo <= i_lo < i_hi <= lo + counts[rank]  # i-bodies are the rank's own
            s = j_lo // n_pad
            assert s != rank and s * n_pad <= j_lo < j_hi <= s * n_pad + counts[s]  # j-bodies of ONE other slot
            pairs += (i_hi - i_lo) * (j_hi - j_lo)
    assert pairs == n * (n - 1) // 2
    assert lib.nbody_shard_pair_workspace_bytes(world, n_pad) > 0


def test_rollout_and_integrator_entry_points_reject_bad_arguments_without_a_device():
    """f4 (trainer.py:217-226): argument checks of the kick/drift entry points and of galaxify.rollout on the CPU."""
    import pytest
    import torch

    from galaxify import rollout

    lib = _native.lib()
    buf = np.zeros(16, dtype=np.float32)
    p = buf.ctypes.data
    assert lib.nbody_kick_drift_f32(None, p, p, p, p, 4, 0.01, 0.005, None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_kick_drift_f32(p, p, p, p, p, -1, 0.01, 0.005, None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_kick_f32(p, None, p, 4, 0.005, None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_momentum_f32(p, p, p, 0, p, None) == _native.ERR_INVALID_ARGUMENT
    assert lib.nbody_pair_min_bodies() == 32768
    z = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rollout.kick_drift(z, z, z, 0.01)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rollout.step(lambda pos, feats: pos, z, z, torch.ones((4, 1)), z, 0.01)
