"""Parity of the raw C-ABI kernels (through ctypes) against the CPU oracle.  GPU only."""

import numpy as np
import pytest
import torch

from cagpjax_b200 import _native as N
from oracle import cagp_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = {torch.float64: 2e-12, torch.float32: 2e-4}


def _data(n, m, d, k, seed, ard):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(m, d, generator=g, dtype=torch.float64) * 1.5
    x2 = torch.randn(n, d, generator=g, dtype=torch.float64) * 1.5
    s = O.normalize_by_blocks(torch.randn(n, generator=g, dtype=torch.float64), k)
    ls = (0.8 + 0.6 * torch.rand(d if ard else 1, generator=g, dtype=torch.float64))
    return x1, x2, s, ls


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,m,d,k,ard", [(9, 9, 1, 4, False), (5, 7, 2, 2, False), (6, 6, 3, 3, True),
                                         (1037, 1037, 3, 10, False), (1500, 700, 5, 7, True),
                                         (2000, 333, 8, 64, False), (777, 1111, 16, 5, True),
                                         (3, 4, 2, 5, False), (600, 600, 20, 3, False)])
def test_ks_fwd_matches_oracle(dtype, family, n, m, d, k, ard):
    x1, x2, s, ls = _data(n, m, d, k, 1, ard)
    ref = O.projected_kernel_blocksparse(family, x1.numpy(), x2.numpy(), ls.numpy(), s.numpy(), k)
    X1, X2, S, LS = (t.to(DEV, dtype) for t in (x1, x2, s, ls))
    xs1 = N.scale_inputs(family, X1, LS)
    xs2 = N.scale_inputs(family, X2, LS)
    Mt = N.ks_blocksparse_fwd(family, xs1, xs2, d, S, k, LS)
    assert Mt.shape == (k, m)
    assert _rel(Mt.T.cpu().numpy(), ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,m,d,k,ard", [(9, 9, 1, 4, False), (6, 6, 3, 3, True), (1037, 1037, 3, 10, False),
                                         (1500, 700, 5, 7, True), (900, 333, 8, 64, False), (777, 500, 16, 5, True),
                                         (3, 4, 2, 5, False)])
def test_ks_bwd_matches_autograd(dtype, family, n, m, d, k, ard):
    x1, x2, s, ls = _data(n, m, d, k, 2, ard)
    g = torch.Generator().manual_seed(5)
    Gt = torch.randn(k, m, generator=g, dtype=torch.float64)
    # oracle: L = sum_{i,b} Gt[b,i] * (K'(x1,x2) S)[i,b], reverse mode w.r.t. s and lengthscale
    s_r = s.clone().requires_grad_(True)
    ls_r = ls.clone().requires_grad_(True)
    K = O.cross_covariance(family, x1, x2, ls_r, torch.tensor(1.0, dtype=torch.float64))
    M = O.bds_rmatmat(s_r, k, K)  # (m, k) = K @ S
    L = (Gt.T * M).sum()
    gs, gl = torch.autograd.grad(L, [s_r, ls_r])
    X1, X2, S, LS, G = (t.to(DEV, dtype).contiguous() for t in (x1, x2, s, ls, Gt))
    xs1 = N.scale_inputs(family, X1, LS)
    xs2 = N.scale_inputs(family, X2, LS)
    s_bar, ls_bar = N.ks_blocksparse_bwd(family, xs1, xs2, d, S, k, G, LS)
    tol = TOL[dtype] * 5
    assert _rel(s_bar.cpu().numpy(), gs.numpy()) < tol
    assert _rel(ls_bar.cpu().numpy(), gl.numpy()) < tol * 20  # sums of signed terms; relative to max


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n_total,row_offset,m,k", [
    (1037, 0, 1037, 10), (1037, 300, 500, 10), (9, 0, 9, 4), (50, 7, 43, 64),
    # float64 tensor-core kernels (csrc/gemm_dmma.cuh): several 128 x 128 tiles incl. partial ones, odd leading
    # dimensions (8-byte copy path), several m-splits of the Gram kernel, k odd
    (3001, 0, 3001, 130), (5000, 1000, 2999, 257), (4096, 0, 4096, 256), (70000, 0, 70000, 300), (2050, 2, 2047, 129)])
def test_row_projections(dtype, n_total, row_offset, m, k):
    g = torch.Generator().manual_seed(3)
    Mt = torch.randn(k, m, generator=g, dtype=torch.float64)
    s = torch.randn(n_total, generator=g, dtype=torch.float64)
    yt = torch.randn(n_total, generator=g, dtype=torch.float64)
    Abar = torch.randn(k, k, generator=g, dtype=torch.float64)
    cbar = torch.randn(k, generator=g, dtype=torch.float64)
    W = torch.randn(k, k, generator=g, dtype=torch.float64)
    blk = O.block_index(n_total, k)[row_offset:row_offset + m]
    sl, yl = s[row_offset:row_offset + m], yt[row_offset:row_offset + m]
    A_ref = torch.zeros(k, k, dtype=torch.float64).index_add_(0, blk, sl[:, None] * Mt.T)
    c_ref = Mt @ yl
    B_ref = Mt @ Mt.T
    G_ref = W @ Mt + Abar[blk].T * sl[None, :] + cbar[:, None] * yl[None, :]
    sdir_ref = (Abar[blk] * Mt.T).sum(1)
    ytbar_ref = Mt.T @ cbar
    d = lambda t: t.to(DEV, dtype).contiguous()
    A, c = N.project_rows(d(Mt), row_offset, n_total, d(sl), d(yl))
    B = N.gram_mtm(d(Mt))
    G = N.assemble_cotangent(d(Mt), row_offset, n_total, d(W), d(Abar), d(cbar), d(sl), d(yl))
    sdir, ytbar = N.project_rows_bwd(d(Mt), row_offset, n_total, d(Abar), d(cbar))
    tol = TOL[dtype] * 5
    for got, ref in [(A, A_ref), (c, c_ref), (B, B_ref), (G, G_ref), (sdir, sdir_ref), (ytbar, ytbar_ref)]:
        assert _rel(got.cpu().numpy(), ref.numpy()) < tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("m,k", [(700, 12), (3001, 130), (4096, 256), (2047, 257)])
def test_predict_moments(dtype, m, k):
    g = torch.Generator().manual_seed(4)
    Mt = torch.randn(k, m, generator=g, dtype=torch.float64) * 0.1
    w = torch.randn(k, generator=g, dtype=torch.float64)
    R = torch.randn(k, k, generator=g, dtype=torch.float64)
    Pinv = R @ R.T / k
    var, mu = torch.tensor(1.3, dtype=torch.float64), torch.tensor(0.4, dtype=torch.float64)
    mean_ref = mu + var * (Mt.T @ w)
    var_ref = var - var**2 * ((Mt.T @ Pinv) * Mt.T).sum(1)
    d = lambda t: t.to(DEV, dtype).contiguous()
    mean, v = N.predict_moments(d(Mt), d(w), d(Pinv), d(var), d(mu))
    assert _rel(mean.cpu().numpy(), mean_ref.numpy()) < TOL[dtype] * 5
    assert _rel(v.cpu().numpy(), var_ref.numpy()) < TOL[dtype] * 5


SYM_SHAPES = [(9, 1, 4, False), (5, 2, 2, False), (6, 3, 3, True), (1037, 3, 10, False), (1500, 5, 7, True),
              (2000, 8, 64, False), (777, 16, 5, True), (3, 2, 5, False), (600, 20, 3, False), (4100, 3, 3, False),
              (2500, 3, 2, True), (64, 2, 64, False), (1000, 1, 1, False),
              # SPREAD shape corner cases: even-k half offset, k not a multiple of the rows per thread, a last
              # block longer than the CTA (atomics), blocks of exactly 256 rows, two blocks only
              (300, 3, 2, False), (513, 2, 2, False), (1799, 3, 7, False), (2048, 4, 8, True), (1500, 8, 6, False),
              (700, 3, 3, False)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,d,k,ard", SYM_SHAPES)
def test_ks_sym_fwd_matches_oracle(dtype, family, n, d, k, ard):
    x1, x2, s, ls = _data(n, n, d, k, 1, ard)
    ref = O.projected_kernel_blocksparse(family, x2.numpy(), x2.numpy(), ls.numpy(), s.numpy(), k)
    X2, S, LS = (t.to(DEV, dtype) for t in (x2, s, ls))
    xs = N.scale_inputs(family, X2, LS)
    Mt = N.ks_blocksparse_fwd_sym(family, xs, d, S, k, LS)
    assert _rel(Mt.T.cpu().numpy(), ref) < TOL[dtype]
    # a two-"rank" split of the action blocks covers every entry exactly once
    half = k // 2
    Ma = N.ks_blocksparse_fwd_sym(family, xs, d, S, k, LS, 0, half)
    Mb = N.ks_blocksparse_fwd_sym(family, xs, d, S, k, LS, half, k)
    assert _rel((Ma + Mb).T.cpu().numpy(), ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,d,k,ard", SYM_SHAPES)
def test_ks_sym_bwd_matches_nonsymmetric(dtype, family, n, d, k, ard):
    x1, x2, s, ls = _data(n, n, d, k, 2, ard)
    g = torch.Generator().manual_seed(5)
    Gt = torch.randn(k, n, generator=g, dtype=torch.float64)
    X2, S, LS, G = (t.to(DEV, dtype).contiguous() for t in (x2, s, ls, Gt))
    xs = N.scale_inputs(family, X2, LS)
    s_ref, l_ref = N.ks_blocksparse_bwd(family, xs, xs, d, S, k, G, LS)  # checked against autograd above
    s_bar, ls_bar = N.ks_blocksparse_bwd_sym(family, xs, d, S, k, G, LS)
    tol = TOL[dtype] * 20
    assert _rel(s_bar.cpu().numpy(), s_ref.cpu().numpy()) < tol
    assert _rel(ls_bar.cpu().numpy(), l_ref.cpu().numpy()) < tol * 20
    half = k // 2
    sa, la = N.ks_blocksparse_bwd_sym(family, xs, d, S, k, G, LS, 0, half)
    sb, lb = N.ks_blocksparse_bwd_sym(family, xs, d, S, k, G, LS, half, k)
    assert _rel((sa + sb).cpu().numpy(), s_ref.cpu().numpy()) < tol
    assert _rel((la + lb).cpu().numpy(), l_ref.cpu().numpy()) < tol * 20


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,m,d,p,ard", [(9, 5, 2, 1, False), (7, 10, 2, 10, False), (1037, 700, 3, 1, True),
                                         (900, 1100, 8, 17, False), (2300, 130, 16, 256, True), (700, 700, 5, 300, False),
                                         (5000, 64, 3, 4, False)])
def test_ks_dense_fwd_bwd_match_oracle(dtype, family, n, m, d, p, ard):
    """LazyKernel @ dense (reference tests/test_operators.py:262-313): values and reverse mode."""
    x1, x2, _, ls = _data(n, m, d, 1, 3, ard)
    g = torch.Generator().manual_seed(7)
    V = torch.randn(n, p, generator=g, dtype=torch.float64)
    Yb = torch.randn(m, p, generator=g, dtype=torch.float64)
    ls_r, V_r = ls.clone().requires_grad_(True), V.clone().requires_grad_(True)
    Yref = O.cross_covariance(family, x1, x2, ls_r, torch.tensor(1.0, dtype=torch.float64)) @ V_r
    gl, gV = torch.autograd.grad((Yref * Yb).sum(), [ls_r, V_r])
    X1, X2, LS, Vd, Ybd = (t.to(DEV, dtype).contiguous() for t in (x1, x2, ls, V, Yb))
    xs1, xs2 = N.scale_inputs(family, X1, LS), N.scale_inputs(family, X2, LS)
    Y = N.ks_dense_fwd(family, xs1, xs2, d, Vd, LS)
    tol = TOL[dtype] * 10
    assert _rel(Y.cpu().numpy(), Yref.detach().numpy()) < tol
    Vbar = N.ks_dense_fwd(family, xs2, xs1, d, Ybd, LS)
    assert _rel(Vbar.cpu().numpy(), gV.numpy()) < tol
    lbar = N.ks_dense_bwd_ls(family, xs1, xs2, d, Vd, Ybd, LS)
    assert _rel(lbar.cpu().numpy(), gl.numpy()) < tol * 50


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("family,n,d,k", [("rbf", 1037, 3, 10), ("matern32", 2000, 8, 64), ("rbf", 4100, 3, 4), ("rbf", 700, 2, 7)])
def test_ks_sym_fwd_peer_panels_emulated_on_one_gpu(world, family, n, d, k):
    """The peer-memory forward (compute + exchange in one kernel): ranks emulated sequentially on one GPU,
    every rank's panel being an ordinary buffer of this device.  After all ranks' tiles have run, panel h
    holds the complete (K' S)^T restricted to the rows of rank h."""
    from cagpjax_b200._fused import RowShard

    x1, x2, s, ls = _data(n, n, d, k, 4, False)
    ref = O.projected_kernel_blocksparse(family, x2.numpy(), x2.numpy(), ls.numpy(), s.numpy(), k)  # (n, k)
    X2, S, LS = (t.to(DEV) for t in (x2, s, ls))
    xs = N.scale_inputs(family, X2, LS, X2[0])
    bb = N.bounding_box(xs)
    shards = [RowShard.for_rank_blocks(n, k, r, world) for r in range(world)]
    row_begin = [sh.begin for sh in shards] + [n]
    ld = max(sh.end - sh.begin for sh in shards)
    panels = [torch.zeros(k, ld, dtype=torch.float64, device=DEV) for _ in range(world)]
    for me, sh in enumerate(shards):
        N.ks_blocksparse_fwd_sym_peer(family, xs, d, S, k, LS, sh.a_begin, sh.a_end, [p.data_ptr() for p in panels],
                                      row_begin, me, ld, bbox=bb)
    for h, sh in enumerate(shards):
        got = panels[h][:, :sh.end - sh.begin].T.cpu().numpy()
        assert _rel(got, ref[sh.begin:sh.end]) < 2e-12
