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
