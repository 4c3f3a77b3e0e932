"""The pair path's decomposition of the interaction matrix over ranks (csrc/capi.cu: shard_pair_geometry, exported as
the pure host function nbody_shard_pair_blocks): every unordered pair of bodies exactly once over all ranks, and equal
work per rank. No reference counterpart (the reference is single-device); the contract is SURVEY.md 8e."""

import ctypes

import numpy as np
import pytest

from galaxify import _native
from galaxify.sharded import shard_layout


def blocks_of(n, world, rank):
    n_pad, _ = shard_layout(n, world)
    buf = (ctypes.c_int * (5 * 16))()
    k = _native.lib().nbody_shard_pair_blocks(n, world, n_pad, rank, buf, 16)
    assert k >= 1, (k, _native.lib().nbody_last_error())
    return [tuple(buf[5 * b : 5 * b + 5]) for b in range(k)]


@pytest.mark.parametrize("n,world", [(1, 1), (50, 1), (50, 2), (51, 3), (64, 4), (4000, 4), (4001, 5), (3000, 6), (3333, 7),
                                     (4097, 8), (6200, 8), (5000, 16)])
def test_every_unordered_pair_exactly_once(n, world):
    if world > n:
        pytest.skip("more ranks than bodies")
    cover = np.zeros((n, n), dtype=np.int32)
    for rank in range(world):
        bl = blocks_of(n, world, rank)
        assert bl[0][4] == 1 and all(b[4] == 0 for b in bl[1:])  # the own triangle first, then rectangles
        for i_lo, i_hi, j_lo, j_hi, tri in bl:
            assert 0 <= i_lo <= i_hi <= n and 0 <= j_lo <= j_hi <= n  # never a padding entry
            if tri:
                assert (i_lo, i_hi) == (j_lo, j_hi)
                cover[i_lo:i_hi, i_lo:i_hi] += np.triu(np.ones((i_hi - i_lo, i_hi - i_lo), dtype=np.int32), 1)
            else:
                cover[i_lo:i_hi, j_lo:j_hi] += 1
    sym = cover + cover.T
    off = ~np.eye(n, dtype=bool)
    assert (sym[off] == 1).all() and (np.diag(sym) == 0).all()


@pytest.mark.parametrize("n,world", [(262144, 2), (262144, 4), (262144, 8), (1 << 20, 8), (1 << 20, 3), (1000003, 6)])
def test_work_is_balanced_over_ranks(n, world):
    work = []
    for rank in range(world):
        w = 0
        for i_lo, i_hi, j_lo, j_hi, tri in blocks_of(n, world, rank):
            w += (i_hi - i_lo) * (i_hi - i_lo - 1) // 2 if tri else (i_hi - i_lo) * (j_hi - j_lo)
        work.append(w)
    assert sum(work) == n * (n - 1) // 2
    assert max(work) <= 1.02 * min(work), work


def test_block_query_argument_errors():
    lib = _native.lib()
    buf = (ctypes.c_int * 80)()
    assert lib.nbody_shard_pair_blocks(100, 4, 25, 4, buf, 16) == _native.ERR_INVALID_ARGUMENT  # slot out of range
    assert lib.nbody_shard_pair_blocks(100, 4, 20, 0, buf, 16) == _native.ERR_INVALID_ARGUMENT  # slots do not hold n
    assert lib.nbody_shard_pair_blocks(100, 8, 13, 0, buf, 2) == _native.ERR_WORKSPACE          # too few output slots
    assert lib.nbody_shard_pair_blocks(100, 8, 13, 0, None, 16) == _native.ERR_INVALID_ARGUMENT
