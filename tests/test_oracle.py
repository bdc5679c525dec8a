"""CPU tests of the oracle: the relational identities the reference's own tests assert,
independent anchors (exact GP, finite differences, literal vs structured) and the golden fixtures."""

import math
import os

import numpy as np
import pytest
import torch

from oracle import cagp_oracle as O
from tests.golden.make_golden import CASES, make_case

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "cagp_golden.npz"))
DT = torch.float64


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=DT)


# ---- BlockDiagonalSparse (reference tests/test_operators.py:117-154) ----------------------------
@pytest.mark.parametrize("n,k", [(5, 2), (9, 4), (6, 3), (1037, 10), (3, 5)])
def test_block_diagonal_sparse_consistency(n, k):
    s = _rand(n)
    dense = O.bds_dense(s, k)
    assert dense.shape == (n, k)
    # exactly one non-zero per row at blk(i)
    blk = O.block_index(n, k)
    ref = torch.zeros(n, k, dtype=DT); ref[torch.arange(n), blk] = s
    assert torch.equal(dense, ref)
    # I @ op == op @ I == to_dense ; transpose consistency
    assert torch.allclose(O.bds_rmatmat(s, k, torch.eye(n, dtype=DT)), dense)
    X = _rand(4, n, seed=1)
    assert torch.allclose(O.bds_rmatmat(s, k, X), X @ dense)
    Y = _rand(k, 3, seed=2)
    assert torch.allclose(O.bds_matmat(s, k, Y), dense @ Y)


def test_block_layout_matches_reference_rule():
    assert O.block_layout(1_000_000, 1024) == (976, 1023, 998_448)
    assert O.block_layout(100_000, 512) == (195, 511, 99_645)
    assert O.block_layout(1000, 10) == (100, 10, 1000)
    starts = O.block_starts(1_000_000, 1024)
    assert starts[-1] - starts[-2] == 1552 and len(starts) == 1025


# ---- normalisation (reference tests/test_policies.py:233-273) -----------------------------------
@pytest.mark.parametrize("n,k", [(12, 2), (13, 3), (20, 3)])
def test_normalize_by_blocks(n, k):
    v = O.normalize_by_blocks(_rand(n) * 3.0, k)
    starts = O.block_starts(n, k)
    for b in range(k):
        assert math.isclose(float(v[starts[b]:starts[b + 1]].norm()), 1.0, rel_tol=1e-12)
    z = O.normalize_by_blocks(torch.zeros(n, dtype=DT), k)
    assert torch.equal(z, torch.zeros(n, dtype=DT))  # zero-norm guard


# ---- sparse congruence (reference tests/test_linalg.py:69-85) -----------------------------------
@pytest.mark.parametrize("n,k", [(7, 3), (10, 2)])
def test_congruence_blocksparse_diagonal(n, k):
    s, dvals = _rand(n), _rand(n, seed=3)
    S = O.bds_dense(s, k)
    assert torch.allclose(torch.diag(O.congruence_bds_diag(s, k, dvals)), S.T @ torch.diag(dvals) @ S)
    assert torch.allclose(torch.diag(O.congruence_bds_diag(s, k, torch.tensor(0.3, dtype=DT))), 0.3 * S.T @ S)


# ---- LazyKernel (reference tests/test_operators.py:183-273) -------------------------------------
@pytest.mark.parametrize("shape", [(9, 5), (7, 10)])
@pytest.mark.parametrize("batch_size", [1, 3, None])
@pytest.mark.parametrize("max_memory_mb", [0.0, 32])
def test_lazy_kernel_matches_dense(shape, batch_size, max_memory_mb):
    x1, x2 = _rand(shape[0], 2, seed=4), _rand(shape[1], 2, seed=5)
    one = torch.tensor(1.0, dtype=DT)
    dense = O.cross_covariance("rbf", x1, x2, one, one)
    bs = None if batch_size is None else batch_size * max(shape)
    got = O.lazy_kernel_matmat("rbf", x1, x2, one, one, torch.eye(shape[1], dtype=DT), max_memory_mb=max_memory_mb, batch_size=bs)
    assert torch.allclose(got, dense)
    got = O.lazy_kernel_rmatmat("rbf", x1, x2, one, one, torch.eye(shape[0], dtype=DT), max_memory_mb=max_memory_mb, batch_size=bs)
    assert torch.allclose(got, dense)
    r, c = O.lazy_batch_sizes(shape, 8, max_memory_mb, bs)
    if bs is None:
        assert (r, c) == (1, 1) if max_memory_mb == 0.0 else (r > 1 and c > 1)
    else:
        assert r == c == bs
    assert O.lazy_batch_sizes((1_000_000, 1_000_000), 8)[0] == 134


def test_kernel_formulas_closed_form_values():
    x = torch.tensor([[0.0, 0.0]], dtype=DT); z = torch.tensor([[3.0, 4.0]], dtype=DT)  # distance 5
    ls, var = torch.tensor(2.0, dtype=DT), torch.tensor(1.5, dtype=DT)
    tau = 2.5
    assert math.isclose(float(O.cross_covariance("rbf", x, z, ls, var)), 1.5 * math.exp(-0.5 * tau**2), rel_tol=1e-15)
    assert math.isclose(float(O.cross_covariance("matern12", x, z, ls, var)), 1.5 * math.exp(-tau), rel_tol=1e-15)
    r3 = math.sqrt(3) * tau
    assert math.isclose(float(O.cross_covariance("matern32", x, z, ls, var)), 1.5 * (1 + r3) * math.exp(-r3), rel_tol=1e-15)
    r5 = math.sqrt(5) * tau
    assert math.isclose(float(O.cross_covariance("matern52", x, z, ls, var)), 1.5 * (1 + r5 + 5 / 3 * tau**2) * math.exp(-r5), rel_tol=1e-15)
    for fam in O.KERNEL_FAMILIES:  # zero distance -> variance
        assert math.isclose(float(O.cross_covariance(fam, x, x, ls, var)), 1.5, rel_tol=1e-15)


# ---- Cholesky solver (reference tests/test_solvers.py:104-244) ----------------------------------
@pytest.mark.parametrize("n", [4, 10])
def test_cholesky_solver_identities(n):
    R = _rand(n, n, seed=6)
    A = R @ R.T + 0.1 * torch.eye(n, dtype=DT)
    L = O.lower_cholesky_dense(A, 1e-3)
    Aj = A + 1e-3 * torch.eye(n, dtype=DT)
    assert torch.allclose(L @ L.T, Aj)
    b = _rand(n, seed=7)
    assert torch.allclose(O.chol_solve(L, b), torch.linalg.solve(Aj, b), rtol=1e-8 * n)
    B = _rand(n, 3, seed=8)
    assert torch.allclose(O.chol_inv_congruence(L, B), B.T @ torch.linalg.solve(Aj, B), rtol=1e-10)
    assert torch.allclose(O.chol_inv_quad(L, b), b @ torch.linalg.solve(Aj, b), rtol=1e-10)
    assert torch.allclose(O.chol_logdet(L), torch.linalg.slogdet(Aj)[1])
    R2 = _rand(n, n, seed=9); Bm = R2 @ R2.T + torch.eye(n, dtype=DT)
    assert torch.allclose(O.chol_trace_solve(L, torch.linalg.cholesky(Bm)), torch.trace(torch.linalg.solve(Aj, Bm)), rtol=1e-12 * n * 100)
    dq = torch.rand(n, dtype=DT, generator=torch.Generator().manual_seed(1)) + 0.1
    assert torch.allclose(O.lower_cholesky_diag(dq, 1e-3), torch.sqrt(dq + 1e-3))


# ---- model-level identities (reference tests/test_models.py:193-308) ----------------------------
def _problem(family="rbf", n=30, d=1, k=5, seed=0, normalize=True):
    x, y, s, ls = make_case(family, n, d, k, False, normalize, seed)
    return x, y, O.make_params(ls, 1.0, 0.1, 3.0, s)


@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
@pytest.mark.parametrize("n,k", [(10, 3), (30, 5)])
def test_prior_kl_nonnegative_and_explicit(family, n, k):
    x, y, p = _problem(family, n, 1, k)
    st = O.cagp_init(family, x, y, p, k)
    kl = O.cagp_prior_kl(st)
    assert torch.isfinite(kl) and kl >= 0
    # explicit KL(q || p) between the predictive at the training inputs and the prior (+ jitter)
    mean, var, cov = O.cagp_predict(family, st, p, None, full_cov=True)
    jit = 1e-6
    Kp = O.cross_covariance(family, x, x, p.lengthscale, p.variance) + jit * torch.eye(n, dtype=DT)
    Kq = cov + jit * torch.eye(n, dtype=DT)
    mp = p.mean_constant * torch.ones(n, dtype=DT)
    Lp = torch.linalg.cholesky(Kp)
    diff = mp - mean
    kl_explicit = 0.5 * (torch.trace(torch.cholesky_solve(Kq, Lp)) + diff @ torch.cholesky_solve(diff[:, None], Lp)[:, 0]
                         - n + torch.linalg.slogdet(Kp)[1] - torch.linalg.slogdet(Kq)[1])
    assert torch.allclose(kl, kl_explicit, rtol=1e-3)


def test_predict_no_inputs_equals_predict_train_inputs():
    x, y, p = _problem("rbf", 30, 1, 5)
    st = O.cagp_init("rbf", x, y, p, 5)
    a = O.cagp_predict("rbf", st, p, None, full_cov=True)
    b = O.cagp_predict("rbf", st, p, x, full_cov=True)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, atol=1e-7)


def test_exact_gp_when_actions_are_complete():
    """n_actions = n makes the CaGP predictive equal to the exact GP (reference tests/test_models.py:206-235)."""
    n = 30
    x, y, p = _problem("rbf", n, 1, n)
    z = torch.linspace(-0.5, 4.5, 20, dtype=DT)[:, None]
    st = O.cagp_init("rbf", x, y, p, n)
    mean, var, cov = O.cagp_predict("rbf", st, p, z, full_cov=True)
    m_ex, cov_ex = O.exact_gp_predict("rbf", x, y, z, p.lengthscale, p.variance, p.obs_stddev, p.mean_constant)
    assert torch.allclose(mean, m_ex, atol=1e-4)
    assert torch.allclose(cov, cov_ex, atol=1e-5)


@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
def test_elbo_identity_and_structured_form(family):
    n, d, k = 437, 3, 10  # overhang: 43 x 9 + 50
    x, y, s, ls = make_case(family, n, d, k, True, False, 3)
    p = O.make_params(ls, 1.3, 0.2, 0.5, s)
    st = O.cagp_init(family, x, y, p, k)
    elbo = O.cagp_elbo(family, st, p)
    mean, var = O.cagp_predict(family, st, p)
    explicit = O.expected_log_likelihood(y, mean, var, p.obs_stddev).sum() - O.cagp_prior_kl(st)
    assert torch.allclose(elbo, explicit)
    structured = O.elbo_structured(family, x.numpy(), y.numpy(), ls.numpy(), 1.3, 0.2, 0.5, s.numpy(), k)
    assert abs(float(elbo) - structured) < 1e-11 * abs(structured)
    # S^T K^ S = var * A' + diag(dvec)
    stats = O.structured_stats(family, x.numpy(), y.numpy(), ls.numpy(), 0.5, s.numpy(), k)
    P = 1.3 * torch.from_numpy(stats["A"]) + torch.diag(0.04 * torch.from_numpy(stats["q"]))
    assert torch.allclose(P, st.cov_prior_proj, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("family", O.KERNEL_FAMILIES)
def test_gradients_against_finite_differences(family):
    """Mirrors jax.test_util.check_grads(order=1, modes=['rev']) of the reference tests."""
    n, d, k = 60, 2, 6
    x, y, s, ls = make_case(family, n, d, k, True, False, 5)
    p = O.make_params(ls, 1.3, 0.3, 0.5, s)
    val, grads = O.elbo_and_grad_literal(family, x, y, p, k)
    h = 1e-6

    def f(**kw):
        q = O.make_params(kw.get("lengthscale", ls), kw.get("variance", 1.3), kw.get("obs_stddev", 0.3),
                          kw.get("mean_constant", 0.5), kw.get("actions", s))
        return float(O.elbo_literal(family, x, y, q, k))

    for name, base in [("variance", 1.3), ("obs_stddev", 0.3), ("mean_constant", 0.5)]:
        fd = (f(**{name: base + h}) - f(**{name: base - h})) / (2 * h)
        assert math.isclose(fd, float(grads[name]), rel_tol=2e-6, abs_tol=1e-5), name
    for t in range(d):
        e = torch.zeros(d, dtype=DT); e[t] = h
        fd = (f(lengthscale=ls + e) - f(lengthscale=ls - e)) / (2 * h)
        assert math.isclose(fd, float(grads["lengthscale"][t]), rel_tol=2e-6, abs_tol=1e-5)
    for j in [0, 7, n - 1]:
        e = torch.zeros(n, dtype=DT); e[j] = h
        fd = (f(actions=s + e) - f(actions=s - e)) / (2 * h)
        assert math.isclose(fd, float(grads["actions"][j]), rel_tol=2e-6, abs_tol=1e-5)


def test_row_shard_partials_sum_to_unsharded_statistics():
    """every n-sized quantity enters through row sums, so row shards add up exactly."""
    n, d, k = 437, 3, 10
    x, y, s, ls = make_case("rbf", n, d, k, False, True, 2)
    full = O.structured_stats("rbf", x.numpy(), y.numpy(), ls.numpy(), 0.5, s.numpy(), k)
    M = full["M"]
    blk = O.block_index(n, k).numpy()
    A = np.zeros((k, k)); B = np.zeros((k, k)); c = np.zeros(k)
    yt = y.numpy() - 0.5
    for lo, hi in [(0, 150), (150, 301), (301, n)]:
        Ml = M[lo:hi]
        np.add.at(A, blk[lo:hi], s.numpy()[lo:hi, None] * Ml)
        B += Ml.T @ Ml
        c += Ml.T @ yt[lo:hi]
    assert np.allclose(A, full["A"], rtol=1e-13) and np.allclose(B, full["B"], rtol=1e-13) and np.allclose(c, full["c"], rtol=1e-13)


# ---- golden fixtures ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_golden_vectors(case):
    name, family, n, d, k, ard, normalize = case
    x, y, s, ls = make_case(family, n, d, k, ard, normalize)
    assert np.array_equal(x.numpy(), GOLDEN[f"{name}/x"]) and np.array_equal(s.numpy(), GOLDEN[f"{name}/s"])
    p = O.make_params(ls, 1.3, 0.2, 0.5, s)
    val, grads = O.elbo_and_grad_literal(family, x, y, p, k)
    assert np.allclose(val.numpy(), GOLDEN[f"{name}/elbo"], rtol=1e-12)
    for gname, g in grads.items():
        ref = GOLDEN[f"{name}/grad_{gname}"]
        assert np.max(np.abs(g.numpy() - ref)) <= 1e-9 * (np.max(np.abs(ref)) + 1e-300), gname
