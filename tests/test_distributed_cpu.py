"""World-size-2 gloo test of the row-sharding logic (partition, packed all-reduce of the k x k partials
in the forward, all-reduce of the n-sized cotangents in the backward).  The LOCAL compute is injected:
here an oracle-backed torch implementation stands in for the native kernels, which need a GPU; the
plumbing under test (``_fused.ProjectStats`` / ``RowShard``) is the product code."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cagpjax_b200 import _fused
from oracle import cagp_oracle as O

DT = torch.float64


def _local_stats(family, X, ls, s, yt, k, shard):
    """Differentiable local statistics over the rows [begin, end) (oracle arithmetic)."""
    one = torch.tensor(1.0, dtype=DT)
    K = O.cross_covariance(family, X[shard.begin:shard.end], X, ls, one)
    M = O.bds_rmatmat(s, k, K)  # (m, k)
    blk = O.block_index(shard.n, k)[shard.begin:shard.end]
    A = torch.zeros(k, k, dtype=DT).index_add(0, blk, s[shard.begin:shard.end, None] * M)
    return A, M.T @ M, M.T @ yt[shard.begin:shard.end]


def oracle_local_fwd(family, X, ls, s, yt, k, shard):
    with torch.no_grad():
        A, B, c = _local_stats(family, X, ls, s, yt, k, shard)
    return A, B, c, None


def oracle_local_bwd(family, X, ls, s, yt, k, shard, saved, Abar, Bbar, cbar):
    with torch.enable_grad():
        ls_r, s_r, yt_r = (t.detach().clone().requires_grad_(True) for t in (ls, s, yt))
        A, B, c = _local_stats(family, X, ls_r, s_r, yt_r, k, shard)
        loss = (A * Abar).sum() + (B * Bbar).sum() + (c * cbar).sum()
        gl, gs, gy = torch.autograd.grad(loss, [ls_r, s_r, yt_r])
    return gs, gy[shard.begin:shard.end], gl


def _objective(A, B, c):
    k = A.shape[0]
    P = A + 0.5 * torch.eye(k, dtype=DT)
    L = torch.linalg.cholesky(P)
    return torch.logdet(P) + (torch.cholesky_solve(B, L)).diagonal().sum() + c @ torch.cholesky_solve(c[:, None], L)[:, 0]


def _run(shard, family, X, ls, s, yt, k):
    ls_r, s_r, yt_r = (t.detach().clone().requires_grad_(True) for t in (ls, s, yt))
    A, B, c = _fused.project_stats(family, X, ls_r, s_r, yt_r, k, shard, None, oracle_local_fwd, oracle_local_bwd)
    val = _objective(A, B, c)
    gl, gs, gy = torch.autograd.grad(val, [ls_r, s_r, yt_r])
    return val.detach(), gl, gs, gy


def _worker(rank, world, port, n, d, k, family, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        X = torch.rand(n, d, generator=g, dtype=DT) * 3
        s = torch.randn(n, generator=g, dtype=DT)
        yt = torch.randn(n, generator=g, dtype=DT)
        ls = torch.tensor([0.8, 1.1][:d], dtype=DT)
        shard = _fused.RowShard.for_rank(n, rank, world, dist.group.WORLD)
        sharded = _run(shard, family, X, ls, s, yt, k)
        full = _run(_fused.RowShard.full(n), family, X, ls, s, yt, k)
        errs = [float((a - b).abs().max() / (b.abs().max() + 1e-300)) for a, b in zip(sharded, full)]
        out[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,d,k,family", [(157, 2, 6, "rbf"), (64, 1, 8, "matern32")])
def test_row_sharded_statistics_and_gradients_match_unsharded(n, d, k, family):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, d, k, family, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for rank in range(world):
        assert max(out[rank]) < 1e-12, (rank, out[rank])


# ---- symmetric Gram path: tile ownership and the two exchanges ----------------------------------------
def _tile_owner_is_row_block(a: int, b: int, k: int) -> bool:
    """True if entry (b, i in block a) is a ROW sum of tile (a, b) in the schedule of csrc/ks_sym.cuh."""
    off = (b - a) % k
    if off == 0:
        return True
    if off > k // 2:
        return False
    if 2 * off == k:
        return a < off
    return True


def _sym_worker(rank, world, port, n, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        Q = torch.randn(k, n, generator=g, dtype=DT)   # the complete M'^T every rank should end up sharing
        G = torch.randn(k, n, generator=g, dtype=DT)
        shard = _fused.RowShard.for_rank_blocks(n, k, rank, world, dist.group.WORLD)
        blk = O.block_index(n, k)
        owner_of_block = torch.empty(k, dtype=torch.long)
        for r in range(world):
            b0, b1 = _fused.RowShard.block_range(k, r, world)
            owner_of_block[b0:b1] = r
        # entries of Q this rank's tiles produce
        mask = torch.zeros(k, n, dtype=torch.bool)
        for b in range(k):
            for a in range(k):
                cols = blk == a
                row_sum = _tile_owner_is_row_block(a, b, k)
                producer = owner_of_block[a] if row_sum else owner_of_block[b]
                if producer == rank:
                    mask[b, cols] = True
        Mfull = torch.where(mask, Q, torch.zeros_like(Q))
        Mt = _fused.exchange_tiles_to_rows(Mfull, shard)
        ok_fwd = torch.equal(Mt, Q[:, shard.begin:shard.end])
        Gfull = _fused.gather_cotangent_for_tiles(G[:, shard.begin:shard.end].contiguous(), shard,
                                                  out=torch.full((k, n), float("nan"), dtype=DT))
        ok_bwd = torch.equal(Gfull[:, shard.begin:shard.end], G[:, shard.begin:shard.end]) and \
            torch.equal(Gfull[shard.a_begin:shard.a_end], G[shard.a_begin:shard.a_end])
        # every entry is produced by exactly one rank
        cnt = mask.to(torch.int32)
        dist.all_reduce(cnt)
        out[rank] = (bool(ok_fwd), bool(ok_bwd), bool((cnt == 1).all()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k,world", [(103, 7, 2), (64, 8, 2), (50, 5, 3)])
def test_symmetric_tile_exchange(n, k, world):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sym_worker, args=(world, port, n, k, out), nprocs=world, join=True)
    for rank in range(world):
        assert out[rank] == (True, True, True), (rank, out[rank])


# ---- dense action matrices: row-sharded statistics (DenseProjectStats) -------------------------------------
class _OracleDenseKernels:
    """Stand-in for the three native calls ``DenseProjectStats`` makes; "scaled inputs" are just X^T here."""

    @staticmethod
    def scale_inputs(family, X, lengthscale, origin=None):
        return X.T.contiguous()

    @staticmethod
    def ks_dense_fwd(family, xs1, xs2, d, V, lengthscale):
        return O.cross_covariance(family, xs1.T, xs2.T, lengthscale, torch.tensor(1.0, dtype=DT)) @ V

    @staticmethod
    def ks_dense_bwd_ls(family, xs1, xs2, d, V, Ybar, lengthscale):
        with torch.enable_grad():
            ls = lengthscale.detach().clone().requires_grad_(True)
            Y = O.cross_covariance(family, xs1.T, xs2.T, ls, torch.tensor(1.0, dtype=DT)) @ V
            (g,) = torch.autograd.grad((Y * Ybar).sum(), ls)
        return g.reshape(-1)


def _run_dense(shard, family, X, ls, S, yt):
    ls_r, S_r, yt_r = (t.detach().clone().requires_grad_(True) for t in (ls, S, yt))
    A, B, c = _fused.dense_project_stats(family, X, ls_r, S_r, yt_r, shard, None, _OracleDenseKernels)
    val = _objective(A, B, c)
    gl, gS, gy = torch.autograd.grad(val, [ls_r, S_r, yt_r])
    return val.detach(), gl, gS, gy


def _dense_worker(rank, world, port, n, d, k, family, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        X = torch.rand(n, d, generator=g, dtype=DT) * 3
        S = torch.randn(n, k, generator=g, dtype=DT) / n**0.5
        yt = torch.randn(n, generator=g, dtype=DT)
        ls = torch.tensor([0.8, 1.1][:d], dtype=DT)
        shard = _fused.RowShard.for_rank(n, rank, world, dist.group.WORLD)
        sharded = _run_dense(shard, family, X, ls, S, yt)
        full = _run_dense(_fused.RowShard.full(n), family, X, ls, S, yt)
        # independent reference of the unsharded statistics (plain autograd through the dense formulas)
        ls_r, S_r, yt_r = (t.detach().clone().requires_grad_(True) for t in (ls, S, yt))
        M = O.cross_covariance(family, X, X, ls_r, torch.tensor(1.0, dtype=DT)) @ S_r
        ref_val = _objective(S_r.T @ M, M.T @ M, M.T @ yt_r)
        ref = (ref_val.detach(),) + torch.autograd.grad(ref_val, [ls_r, S_r, yt_r])
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-300))
        out[rank] = [rel(a, b) for a, b in zip(sharded, full)] + [rel(a, b) for a, b in zip(full, ref)]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,d,k,family", [(101, 2, 5, "rbf"), (64, 1, 7, "matern52")])
def test_row_sharded_dense_statistics_and_gradients(n, d, k, family):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dense_worker, args=(world, port, n, d, k, family, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for rank in range(world):
        assert max(out[rank]) < 1e-11, (rank, out[rank])
