"""Parity at the EXACT kernel instances bench.py measures: the small-size parity tests reach
other template instances than the benchmark does (C3: float64 RBF fast path with 976-row blocks and the 1552-row
overhang block split over CTAs; C2: SPREAD shape, 195-row blocks, d = 8 Matern-3/2; C4: rectangular forward at
1M x 1M in float64 and float32).  Here the kernels run at the benchmark's (n, k) on bench.py's own synthetic inputs and
are checked

  * against the CPU oracle on a ROW SAMPLE: row i of M' = K'(X, X) S holds, for all k blocks, both kinds of output
    of the symmetric kernels (row sums where blk(i) is the tile's row block, column sums where it is the column
    block), so a sample of rows covers every code path that writes M';
  * against the independent rectangular kernels for the full-size reverse mode (promoted from tools/c3_crosscheck.py),
    plus an oracle sample of s_bar.

The oracle restates the reference, it is not the reference; what pins it against the reference's own source:
oracle/cagp_oracle.py header and tests/test_reference_golden*.py (incl. mid-size cases at these kernel shapes)."""

import numpy as np
import pytest
import torch

import bench as B
from cagpjax_b200 import _fused, _native as N
from cagpjax_b200.policies.block_sparse import _normalize_by_blocks
from oracle import cagp_oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(name, dtype=torch.float64):
    cfg = B.CONFIGS[name]
    x, y, s = B.synth(cfg, "cpu", torch.float64)
    s = _normalize_by_blocks(s, cfg["k"])
    ls = torch.tensor([cfg.get("lengthscale", 1.0)], dtype=torch.float64)
    return cfg, x, y, s, ls


def _scaled(family, x, ls, dtype=torch.float64):
    X = x.to(DEV, dtype)
    l = ls.to(DEV, dtype)
    xs = N.scale_inputs(family, X, l, _fused._box_centre(X))  # exactly what native_local_forward does
    return X, l, xs, N.bounding_box(xs)


def _sample_rows(n, k, stride, extra_blocks=()):
    """Every stride-th row, the first / last rows of the listed blocks and of the last (overhang) block."""
    bs = n // k
    rows = set(range(0, n, stride))
    for b in tuple(extra_blocks) + (k - 1,):
        lo, hi = b * bs, (n if b == k - 1 else (b + 1) * bs)
        rows.update([lo, lo + 1, lo + (hi - lo) // 2, hi - 2, hi - 1])
        rows.update(range(lo + 1020, min(hi, lo + 1030)))  # around the 1024-row CTA boundary of a split block
    return np.array(sorted(r for r in rows if 0 <= r < n))


def _oracle_rows(family, x, ls, s, k, rows, chunk=32):
    return O.projected_kernel_blocksparse(family, x[rows].numpy(), x.numpy(), ls.numpy(), s.numpy(), k, chunk=chunk)


def _oracle_sbar(family, x, ls, Gt_rows, k, cols):
    """s_bar[j] = sum_i K'(x_i, x_j) Gt[blk(j)][i] for the sampled j (all i)."""
    n = x.shape[0]
    bs = n // k
    out = np.empty(len(cols))
    xl = (x / ls).numpy()
    for t, j in enumerate(cols):
        b = min(j // bs, k - 1)
        Kj, _ = O._np_kernel_unit(family, xl[j:j + 1], xl)
        out[t] = float(Kj[0] @ Gt_rows[b])
    return out


def test_c3_symmetric_forward_at_benchmark_size_matches_oracle_rows():
    cfg, x, y, s, ls = _inputs("c3")
    n, k, d = cfg["n"], cfg["k"], cfg["d"]
    X, l, xs, bbox = _scaled("rbf", x, ls)
    assert N.sym_fast_path("rbf", xs, d, k, l, bbox), "C3 must run the fast-path (expanded form) kernel"
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, d, s.to(DEV), k, l, bbox=bbox)
    rows = _sample_rows(n, k, 7919, extra_blocks=(0, 511, 512))
    ref = _oracle_rows("rbf", x, ls, s, k, rows)
    got = Mt[:, torch.from_numpy(rows).to(DEV)].T
    assert rel_err(got, ref) < 1e-11
    # the general kernel (fast path off: no bounding box) agrees on the same sample
    Mt0 = N.ks_blocksparse_fwd_sym("rbf", xs, d, s.to(DEV), k, l)
    assert rel_err(Mt0[:, torch.from_numpy(rows).to(DEV)].T, ref) < 1e-11
    assert rel_err(Mt, Mt0) < 1e-12


def test_c3_symmetric_backward_at_benchmark_size_matches_rectangular_and_oracle():
    cfg, x, y, s, ls = _inputs("c3")
    n, k, d = cfg["n"], cfg["k"], cfg["d"]
    X, l, xs, bbox = _scaled("rbf", x, ls)
    assert N.sym_fast_path("rbf", xs, d, k, l, bbox, backward=True)
    sd = s.to(DEV)
    Gt = torch.randn(k, n, dtype=torch.float64, device=DEV, generator=torch.Generator(DEV).manual_seed(3)) / n**0.5
    sb, lb = N.ks_blocksparse_bwd_sym("rbf", xs, d, sd, k, Gt, l, bbox=bbox)           # fast path
    sb0, lb0 = N.ks_blocksparse_bwd_sym("rbf", xs, d, sd, k, Gt, l)                    # general symmetric kernel
    sbr, lbr = N.ks_blocksparse_bwd("rbf", xs, xs, d, sd, k, Gt, l)                    # independent rectangular kernel
    for a, b in [(sb, sbr), (sb0, sbr), (lb, lbr), (lb0, lbr)]:
        assert rel_err(a, b) < 1e-11
    bs = n // k
    cols = np.array([0, 5, bs - 1, 511 * bs + 7, 512 * bs, (k - 1) * bs, (k - 1) * bs + 1025, n - 1])
    blocks = sorted({min(j // bs, k - 1) for j in cols})
    Gt_rows = {b: Gt[b].cpu().numpy() for b in blocks}
    ref = _oracle_sbar("rbf", x, ls, Gt_rows, k, cols)
    assert rel_err(sb[torch.from_numpy(cols).to(DEV)], ref) < 1e-11


def test_c2_spread_matern_at_benchmark_size():
    cfg, x, y, s, ls = _inputs("c2")
    n, k, d, fam = cfg["n"], cfg["k"], cfg["d"], cfg["family"]
    X, l, xs, bbox = _scaled(fam, x, ls)
    sd = s.to(DEV)
    Mt = N.ks_blocksparse_fwd_sym(fam, xs, d, sd, k, l, bbox=bbox)
    rows = _sample_rows(n, k, 499, extra_blocks=(0, 1, 255, 256))
    ref = _oracle_rows(fam, x, ls, s, k, rows)
    assert rel_err(Mt[:, torch.from_numpy(rows).to(DEV)].T, ref) < 1e-11
    Gt = torch.randn(k, n, dtype=torch.float64, device=DEV, generator=torch.Generator(DEV).manual_seed(4)) / n**0.5
    sb, lb = N.ks_blocksparse_bwd_sym(fam, xs, d, sd, k, Gt, l, bbox=bbox)
    sbr, lbr = N.ks_blocksparse_bwd(fam, xs, xs, d, sd, k, Gt, l)
    assert rel_err(sb, sbr) < 1e-11 and rel_err(lb, lbr) < 1e-11
    bs = n // k
    cols = np.array([0, bs - 1, bs, 255 * bs + 3, (k - 1) * bs, n - 1])
    Gt_rows = {b: Gt[b].cpu().numpy() for b in sorted({min(j // bs, k - 1) for j in cols})}
    assert rel_err(sb[torch.from_numpy(cols).to(DEV)], _oracle_sbar(fam, x, ls, Gt_rows, k, cols)) < 1e-11


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
def test_c4_predict_cross_covariance_rows_at_benchmark_size(dtype, tol):
    """K'(X*, X) S at m = n = 1M, k = 1024 (rectangular forward, BASELINE config 4): a sample of test rows."""
    cfg, x, y, s, ls = _inputs("c3")
    n, k, d = cfg["n"], cfg["k"], cfg["d"]
    z = torch.rand(n, d, generator=torch.Generator().manual_seed(31), dtype=torch.float64) * 10.0
    X = x.to(DEV, dtype)
    Z = z.to(DEV, dtype)
    l = ls.to(DEV, dtype)
    origin = _fused._box_centre(X)
    xs2 = N.scale_inputs("rbf", X, l, origin)
    xs1 = N.scale_inputs("rbf", Z, l, origin)
    Mt = N.ks_blocksparse_fwd("rbf", xs1, xs2, d, s.to(DEV, dtype), k, l, N.union_bounding_box(xs1, xs2))
    rows = np.arange(0, n, 7919)
    # the oracle sees the same (dtype-rounded) inputs, in float64 arithmetic
    rd = lambda t: t.to(dtype).double()
    ref = O.projected_kernel_blocksparse("rbf", rd(z[rows]).numpy(), rd(x).numpy(), ls.numpy(), rd(s).numpy(), k, chunk=32)
    assert rel_err(Mt[:, torch.from_numpy(rows).to(DEV)].T.double(), ref) < tol


def test_c3_model_step_sym_vs_rect_at_benchmark_size():
    """Whole ELBO + gradient step at C3 through the public API: symmetric (fast-path) kernels vs the rectangular ones."""
    from cagpjax_b200 import gpx

    cfg = B.CONFIGS["c3"]
    x, y, s = B.synth(cfg, DEV, torch.float64)
    res = []
    try:
        for sym in (True, False):
            _fused.USE_SYMMETRIC = sym
            model, data = B.build_model(cfg, x, y, s, torch.device(DEV), torch.float64)
            res.append(gpx.value_and_grad(B.neg_elbo, model, data))
            del model, data
            torch.cuda.empty_cache()
    finally:
        _fused.USE_SYMMETRIC = True
    (v1, g1), (v2, g2) = res
    assert rel_err(v1, v2) < 1e-11
    for name in g1:
        assert rel_err(g1[name], g2[name]) < 1e-9, name
