"""Property-based checks (hypothesis) of the size-independent host logic: block layouts, segment sums, row
partitions, block-diagonal-sparse products, the tile schedule of the symmetric kernels."""

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from cagpjax_b200 import _fused
from cagpjax_b200.operators import BlockDiagonalSparse
from oracle import cagp_oracle as O

DT = torch.float64
shapes = st.integers(1, 400).flatmap(lambda n: st.tuples(st.just(n), st.integers(1, n)))  # (n, k <= n)


@settings(max_examples=200, deadline=None)
@given(shapes)
def test_block_layout_covers_every_row_exactly_once(nk):
    n, k = nk
    block_size, n_main_blocks, n_main = _fused.block_layout(n, k)
    assert block_size == n // k and n_main == n_main_blocks * block_size and n_main <= n
    idx = O.block_index(n, k).numpy()
    assert idx.min() == 0 and idx.max() == k - 1 and np.all(np.diff(idx) >= 0)
    counts = np.bincount(idx, minlength=k)
    assert np.all(counts[:-1] == block_size) and counts[-1] == n - (k - 1) * block_size  # last block absorbs n % k
    starts = O.block_starts(n, k)
    assert starts[0] == 0 and np.all(np.diff(starts) == counts)


@settings(max_examples=100, deadline=None)
@given(shapes, st.integers(0, 2**31 - 1))
def test_segment_sum_and_bds_products_match_dense(nk, seed):
    n, k = nk
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, generator=g, dtype=DT)
    S = O.bds_dense(v, k)
    assert torch.allclose(_fused.segment_sum(v, k), S.sum(dim=0))
    op = BlockDiagonalSparse(v, k)
    X = torch.randn(k, 3, generator=g, dtype=DT)
    Y = torch.randn(2, n, generator=g, dtype=DT)
    assert op.shape == (n, k)
    assert torch.allclose(op @ X, S @ X) and torch.allclose(Y @ op, Y @ S)
    assert torch.allclose(op.to_dense(), S)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 16))
def test_even_row_partition_is_exact(n, world):
    shards = [_fused.RowShard.for_rank(n, r, world) for r in range(world)]
    assert shards[0].begin == 0 and shards[-1].end == n
    assert all(a.end == b.begin for a, b in zip(shards, shards[1:]))
    sizes = [s.end - s.begin for s in shards]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 16).flatmap(lambda w: st.tuples(st.just(w), st.integers(w, 200))).flatmap(
    lambda wk: st.tuples(st.just(wk[0]), st.just(wk[1]), st.integers(wk[1], 5000))))
def test_block_aligned_partition_is_exact_and_aligned(wkn):
    world, k, n = wkn
    shards = [_fused.RowShard.for_rank_blocks(n, k, r, world) for r in range(world)]
    bs = n // k
    assert shards[0].begin == 0 and shards[-1].end == n
    for a, b in zip(shards, shards[1:]):
        assert a.end == b.begin and a.a_end == b.a_begin and a.end == a.a_end * bs
    assert shards[0].a_begin == 0 and shards[-1].a_end == k
    blocks = [s.a_end - s.a_begin for s in shards]
    assert max(blocks) - min(blocks) <= 1
    for r, s in enumerate(shards):
        assert s.peer_rows(r) == (s.begin, s.end)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 60))
def test_symmetric_tile_schedule_covers_every_block_pair_once(k):
    """Tiles (a, (a + off) mod k), off = 0 .. k/2, with the even-k rule (offset k/2 only for a < k/2)."""
    seen = {}
    for a in range(k):
        for off in range(k // 2 + 1):
            if off and 2 * off == k and a >= off:
                continue
            b = (a + off) % k
            key = (min(a, b), max(a, b))
            seen[key] = seen.get(key, 0) + 1
    assert len(seen) == k * (k + 1) // 2 and set(seen.values()) == {1}
