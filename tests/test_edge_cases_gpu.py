"""Edge cases of the CUDA path: degenerate block layouts, far-apart points (underflow / clamp), tiny and
huge lengthscales, single action, float32 predict, distribution surface (log_prob / sample)."""

import math

import numpy as np
import pytest
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import _native as N, gpx
from cagpjax_b200.distributions import GaussianDistribution
from oracle import cagp_oracle as O
from tests.util import make_problem, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,k", [(3, 5), (7, 1), (1, 1), (40, 40), (41, 40), (2049, 2)])
def test_degenerate_block_layouts(n, k):
    """More blocks than points (empty blocks), a single block, one row per block, a 2-block split of a
    long vector (blocks longer than a CTA's rows -> split tiles / atomics)."""
    d = 2
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 3
    s = torch.randn(n, generator=g, dtype=torch.float64)
    ls = torch.tensor([0.9], dtype=torch.float64)
    ref = O.projected_kernel_blocksparse("rbf", x.numpy(), x.numpy(), ls.numpy(), s.numpy(), k)
    xs = N.scale_inputs("rbf", x.to(DEV), ls.to(DEV))
    for Mt in (N.ks_blocksparse_fwd("rbf", xs, xs, d, s.to(DEV), k, ls.to(DEV)),
               N.ks_blocksparse_fwd_sym("rbf", xs, d, s.to(DEV), k, ls.to(DEV))):
        assert rel_err(Mt.T, ref) < 1e-12
    model, data, p, x64, y64 = make_problem("rbf", n, d, k, seed=1)
    val, grads = gpx.value_and_grad(lambda m, dd: m.elbo(m.init(dd)), model, data)
    oval, ograds = O.elbo_and_grad_literal("rbf", x64, y64, p, k)
    assert rel_err(val, oval) < 1e-10
    gs, go = grads["BlockSparsePolicy.nz_values"].cpu(), ograds["actions"]
    # (k = n: S is square and invertible, the ELBO is then independent of S up to the jitter, so the
    #  gradient is a ~1e-6 remainder of O(100) cancelling terms: compare absolutely as well)
    assert rel_err(gs, go) < 1e-8 or float((gs - go).abs().max()) < 1e-9


@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
def test_far_apart_points_underflow_to_zero(family):
    """Distances far beyond the kernel's range (exp underflow in the reference): zeros / ~1e-301, never NaN
    or inf; includes a point 1e9 lengthscales away (high-word overflow of the range reduction -> clamp)."""
    x = torch.tensor([[0.0], [1.0], [50.0], [2000.0], [4e4], [1e9]], dtype=torch.float64)
    s = torch.ones(6, dtype=torch.float64)
    ls = torch.tensor([1.0], dtype=torch.float64)
    ref = O.projected_kernel_blocksparse(family, x.numpy(), x.numpy(), ls.numpy(), s.numpy(), 6)
    xs = N.scale_inputs(family, x.to(DEV), ls.to(DEV), x[0].to(DEV))
    for Mt in (N.ks_blocksparse_fwd(family, xs, xs, 1, s.to(DEV), 6, ls.to(DEV)).cpu(),
               N.ks_blocksparse_fwd_sym(family, xs, 1, s.to(DEV), 6, ls.to(DEV)).cpu()):
        assert torch.isfinite(Mt).all()
        assert np.max(np.abs(Mt.T.numpy() - ref)) < 1e-15  # absolute: the reference underflows to 0 as well


def test_common_offset_is_removed_before_scaling():
    """A data set far from zero (offset 1e7, spacing 0.25): the origin shift keeps differences exact."""
    x = 1e7 + 0.25 * torch.arange(64, dtype=torch.float64)[:, None]
    s = torch.ones(64, dtype=torch.float64)
    ls = torch.tensor([0.7], dtype=torch.float64)
    ref = O.projected_kernel_blocksparse("rbf", (x - 1e7).numpy(), (x - 1e7).numpy(), ls.numpy(), s.numpy(), 4)
    xs = N.scale_inputs("rbf", x.to(DEV), ls.to(DEV), x[0].to(DEV))
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, 1, s.to(DEV), 4, ls.to(DEV))
    assert rel_err(Mt.T, ref) < 1e-13


@pytest.mark.parametrize("ls", [1e-3, 1e3])
def test_extreme_lengthscales(ls):
    n, d, k = 500, 3, 5
    model, data, p, x, y = make_problem("rbf", n, d, k)
    model.posterior.prior.kernel.lengthscale = torch.tensor(ls, dtype=torch.float64, device=DEV)
    p.lengthscale = torch.tensor(ls, dtype=torch.float64)
    with torch.no_grad():
        elbo = model.elbo(model.init(data))
    ref = O.elbo_literal("rbf", x, y, p, k)
    assert torch.isfinite(elbo) and rel_err(elbo, ref) < 1e-9


def test_float32_predict_matches_float64_oracle():
    model, data, p, x, y = make_problem("rbf", 2000, 3, 16, dtype=torch.float32, obs_stddev=0.5)
    g = torch.Generator().manual_seed(5)
    z = torch.rand(300, 3, generator=g, dtype=torch.float64) * 4
    with torch.no_grad():
        st = model.init(data)
        q = model.predict(st, z.to(DEV, torch.float32))
        assert q.mean.dtype == torch.float32 and q.variance.dtype == torch.float32
    # north star: 1e-4 in float32; the oracle sees the same float32-rounded inputs (float64 arithmetic)
    p.actions = gpx.unwrap(model.policy.nz_values).detach().double().cpu()
    ost = O.cagp_init("rbf", data.X.double().cpu(), data.y.double().cpu().reshape(-1), p, 16)
    om, ov = O.cagp_predict("rbf", ost, p, z.to(torch.float32).double())
    assert rel_err(q.mean, om) < 1e-4 and rel_err(q.variance, ov) < 1e-4


def test_gaussian_distribution_surface():
    """distributions.py:48-106: mean / variance / stddev / covariance / log_prob / sample on the lazy
    predictive covariance."""
    model, data, p, x, y = make_problem("rbf", 200, 2, 8)
    g = torch.Generator().manual_seed(6)
    z = torch.rand(25, 2, generator=g, dtype=torch.float64) * 4
    with torch.no_grad():
        st = model.init(data)
        q = model.predict(st, z.to(DEV))
        assert isinstance(q, GaussianDistribution)
        cov = q.covariance().to_dense()
        assert torch.allclose(q.variance, torch.diagonal(cov), atol=1e-12)
        assert torch.allclose(q.stddev**2, q.variance)
        # operator products agree with the dense matrix
        v = torch.randn(25, 3, generator=g, dtype=torch.float64).to(DEV)
        assert torch.allclose(q.scale @ v, cov @ v, atol=1e-10)
        ost = O.cagp_init("rbf", x, y, p, 8)
        om, ov, ocov = O.cagp_predict("rbf", ost, p, z, full_cov=True)
        value = (om + 0.1).to(DEV)
        lp = q.log_prob(value)
        Lc = torch.linalg.cholesky(ocov + 1e-6 * torch.eye(25, dtype=torch.float64))
        diff = value.cpu() - om
        ref = -0.5 * (25 * math.log(2 * math.pi) + 2 * torch.log(torch.diagonal(Lc)).sum()
                      + diff @ torch.cholesky_solve(diff[:, None], Lc)[:, 0])
        assert rel_err(lp, ref) < 1e-8
        samples = q.sample(3, (4000,))
        assert samples.shape == (4000, 25)
        assert torch.allclose(samples.mean(0).cpu(), om, atol=0.05)
        assert torch.allclose(samples.var(0).cpu(), ov, rtol=0.2, atol=1e-3)


@pytest.mark.parametrize("ls", [0.9, 0.2, 0.17, 0.16, 0.05])
@pytest.mark.parametrize("n,k,d", [(1500, 3, 3), (700, 64, 2)])
def test_rbf_range_modes_agree_with_oracle(ls, n, k, d):
    """The fp64 RBF kernels pick one of three arithmetic forms from the bounding box of the scaled inputs:
    expanded / factored distance (|x|^2 < 300 everywhere; SINGLE-shape kernels and the rectangular forward),
    clamp-free exp2 (all q < 990; what the SPREAD-shape kernels of the k = 64 case use), general.  X in [0, 4]^d:
    the lengthscale moves the same data across the thresholds (d = 3: expanded up to ls = 0.17, general from 0.16);
    every form must match the oracle, forward and reverse."""
    g = torch.Generator().manual_seed(4)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 4
    s = torch.randn(n, generator=g, dtype=torch.float64)
    Gt = torch.randn(k, n, generator=g, dtype=torch.float64)
    lst = torch.tensor([ls], dtype=torch.float64)
    lso = lst.clone().requires_grad_(True)
    so = s.clone().requires_grad_(True)
    Ko = O.cross_covariance("rbf", x, x, lso, torch.tensor(1.0, dtype=torch.float64))
    Mo = Ko @ O.bds_dense(so, k)  # (n, k)
    gs, gl = torch.autograd.grad((Mo.T * Gt).sum(), [so, lso])

    xd, sd, lsd = x.to(DEV), s.to(DEV), lst.to(DEV)
    xs = N.scale_inputs("rbf", xd, lsd, 0.5 * (xd.amin(0) + xd.amax(0)))
    bbox = N.bounding_box(xs)
    Mt_sym = N.ks_blocksparse_fwd_sym("rbf", xs, d, sd, k, lsd, bbox=bbox)
    Mt_rect = N.ks_blocksparse_fwd("rbf", xs, xs, d, sd, k, lsd, bbox)
    assert rel_err(Mt_sym.T, Mo) < 1e-11
    assert rel_err(Mt_rect.T, Mo) < 1e-11
    sb, lb = N.ks_blocksparse_bwd_sym("rbf", xs, d, sd, k, Gt.to(DEV), lsd, bbox=bbox)
    assert rel_err(sb, gs) < 1e-10
    assert abs(float(lb[0]) - float(gl[0])) <= 1e-10 * max(abs(float(gl[0])), float((Mo.detach().abs().T * Gt.abs()).sum()))


def test_expanded_form_falls_back_for_wide_data():
    """Same data, shifted far from the origin and spread over hundreds of lengthscales: the centred coordinates
    exceed the |x|^2 < 300 proof, the kernels must take the difference-based forms and stay exact."""
    n, d, k = 900, 3, 5
    g = torch.Generator().manual_seed(5)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 400 + 1e5  # ~340 scaled units wide
    x[: n // 2] = x[n // 2:][: n // 2] + 0.3 * torch.randn(n // 2, d, generator=g, dtype=torch.float64)  # near pairs
    s = torch.randn(n, generator=g, dtype=torch.float64)
    ls = torch.tensor([1.0], dtype=torch.float64)
    ref = O.projected_kernel_blocksparse("rbf", x.numpy(), x.numpy(), ls.numpy(), s.numpy(), k)
    xd = x.to(DEV)
    xs = N.scale_inputs("rbf", xd, ls.to(DEV), 0.5 * (xd.amin(0) + xd.amax(0)))
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, d, s.to(DEV), k, ls.to(DEV), bbox=N.bounding_box(xs))
    assert rel_err(Mt.T, ref) < 1e-10


@pytest.mark.parametrize("family,n,d,k", [("rbf", 2300, 3, 2), ("rbf", 1500, 2, 1), ("rbf", 977, 3, 5), ("matern32", 900, 8, 5)])
def test_symmetric_kernels_stay_inside_their_outputs(family, n, d, k):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with canaries: the output panel
    sits between two guard rows of NaN inside one allocation, inputs are followed by NaN padding, and every
    result must be finite while the guards stay untouched (fast-path and general kernels, split blocks included)."""
    g = torch.Generator().manual_seed(9)
    X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 3).to(DEV)
    s_pad = torch.full((n + 64,), float("nan"), dtype=torch.float64, device=DEV)
    s_pad[:n] = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    s = s_pad[:n]
    ls = torch.ones(1, dtype=torch.float64, device=DEV)
    xs = N.scale_inputs(family, X, ls, X[0])
    bb = N.bounding_box(xs)
    buf = torch.full((k + 2, n), float("nan"), dtype=torch.float64, device=DEV)
    out = buf[1:k + 1]
    out.zero_()
    Mt = N.ks_blocksparse_fwd_sym(family, xs, d, s, k, ls, out=out, bbox=bb)
    assert Mt.data_ptr() == out.data_ptr() and torch.isfinite(out).all()
    assert torch.isnan(buf[0]).all() and torch.isnan(buf[k + 1]).all()
    ref = N.ks_blocksparse_fwd(family, xs, xs, d, s, k, ls)
    assert rel_err(out, ref) < 1e-11
    G = torch.randn(k, n, generator=g, dtype=torch.float64).to(DEV)
    sb, lb = N.ks_blocksparse_bwd_sym(family, xs, d, s, k, G, ls, bbox=bb)
    sbr, lbr = N.ks_blocksparse_bwd(family, xs, xs, d, s, k, G, ls)
    assert torch.isfinite(sb).all() and rel_err(sb, sbr) < 1e-11 and rel_err(lb, lbr) < 1e-11
    assert torch.isnan(s_pad[n:]).all()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_nan_coordinate_poisons_the_bounding_box(dtype):
    """A NaN input coordinate must not let the launch predicates pick an unclamped fast path from a box computed
    over the remaining finite points: the box itself becomes NaN (every predicate is then false), the kernels run
    their general forms and the NaN shows up in the rows / columns it touches, as in the reference (jnp propagates)."""
    n, d, k = 700, 3, 4
    g = torch.Generator().manual_seed(2)
    X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 3).to(dtype).to(DEV)
    ls = torch.ones(1, dtype=dtype, device=DEV)
    s = torch.randn(n, generator=g, dtype=torch.float64).to(dtype).to(DEV)
    xs = N.scale_inputs("rbf", X, ls, X[0])
    assert torch.isfinite(N.bounding_box(xs)).all()
    X[123, 1] = float("nan")
    xs = N.scale_inputs("rbf", X, ls, X[0])
    bb = N.bounding_box(xs)
    dp = xs.shape[0]
    assert torch.isnan(bb[1]) and torch.isnan(bb[dp + 1])
    assert torch.isfinite(bb[0]) and torch.isfinite(bb[dp]) and torch.isfinite(bb[2]) and torch.isfinite(bb[dp + 2])
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, d, s, k, ls, bbox=bb)
    assert torch.isnan(Mt[:, 123]).all()          # row 123 of M' = K'S: every entry sees the NaN point
    blk = 123 // (n // k)
    assert torch.isnan(Mt[blk]).all()             # its own block's column sums over the NaN point
    other = [b for b in range(k) if b != blk]
    cols = torch.ones(n, dtype=torch.bool, device=DEV)
    cols[123] = False
    assert torch.isfinite(Mt[other][:, cols]).all()


@pytest.mark.parametrize("width,level", [(20.0, 1), (36.0, 2), (60.0, 0)])
def test_expanded_form_levels_against_the_oracle(width, level):
    """The three range regimes of the float64 RBF symmetric kernels (kernel_math.cuh, launch_expand_level): every
    scaled |x|^2 < 300 (expanded form, no clamp), < 900 (expanded form, exponent clamp: width 36 has pairs with
    u = |b|^2 - 2 a.b up to ~2100 > 1000) and beyond (difference form).  Forward against the oracle's dense product,
    backward against the rectangular general kernel, all three within 1e-11."""
    n, d, k = 1800, 3, 6
    g = torch.Generator().manual_seed(17)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * width
    x[0], x[1] = torch.zeros(d, dtype=torch.float64), torch.full((d,), width, dtype=torch.float64)  # the two far corners
    x[n // 2:] = x[: n - n // 2] + 0.2 * torch.randn(n - n // 2, d, generator=g, dtype=torch.float64)  # near pairs
    x.clamp_(0.0, width)
    s = torch.randn(n, generator=g, dtype=torch.float64)
    ls = torch.ones(1, dtype=torch.float64)
    ref = O.projected_kernel_blocksparse("rbf", x.numpy(), x.numpy(), ls.numpy(), s.numpy(), k)
    xd, sd, lsd = x.to(DEV), s.to(DEV), ls.to(DEV)
    xs = N.scale_inputs("rbf", xd, lsd, 0.5 * (xd.amin(0) + xd.amax(0)))
    bb = N.bounding_box(xs)
    assert N.sym_fast_path("rbf", xs, d, k, lsd, bb) == level
    assert N.sym_fast_path("rbf", xs, d, k, lsd, bb, backward=True) == level
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, d, sd, k, lsd, bbox=bb)
    assert torch.isfinite(Mt).all() and rel_err(Mt.T, ref) < 1e-11
    G = torch.randn(k, n, generator=g, dtype=torch.float64).to(DEV)
    sb, lb = N.ks_blocksparse_bwd_sym("rbf", xs, d, sd, k, G, lsd, bbox=bb)
    sbr, lbr = N.ks_blocksparse_bwd("rbf", xs, xs, d, sd, k, G, lsd)
    assert torch.isfinite(sb).all() and rel_err(sb, sbr) < 1e-11 and rel_err(lb, lbr) < 1e-11


def test_ard_forward_takes_the_fast_path():
    """ARD lengthscales only change the input scaling, so the float64 RBF forward runs the expanded-form kernel for
    them too (the backward needs per-dimension differences and stays on the general kernel); both against the oracle."""
    n, d, k = 2100, 3, 7
    g = torch.Generator().manual_seed(23)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 6
    s = torch.randn(n, generator=g, dtype=torch.float64)
    ls = torch.tensor([0.7, 1.3, 2.1], dtype=torch.float64)
    ref = O.projected_kernel_blocksparse("rbf", x.numpy(), x.numpy(), ls.numpy(), s.numpy(), k)
    xd, sd, lsd = x.to(DEV), s.to(DEV), ls.to(DEV)
    xs = N.scale_inputs("rbf", xd, lsd, 0.5 * (xd.amin(0) + xd.amax(0)))
    bb = N.bounding_box(xs)
    assert N.sym_fast_path("rbf", xs, d, k, lsd, bb) == 1
    assert N.sym_fast_path("rbf", xs, d, k, lsd, bb, backward=True) == 0
    Mt = N.ks_blocksparse_fwd_sym("rbf", xs, d, sd, k, lsd, bbox=bb)
    assert rel_err(Mt.T, ref) < 1e-12
    G = torch.randn(k, n, generator=g, dtype=torch.float64).to(DEV)
    sb, lb = N.ks_blocksparse_bwd_sym("rbf", xs, d, sd, k, G, lsd, bbox=bb)
    sbr, lbr = N.ks_blocksparse_bwd("rbf", xs, xs, d, sd, k, G, lsd)
    assert lb.numel() == d and rel_err(sb, sbr) < 1e-11 and rel_err(lb, lbr) < 1e-11
