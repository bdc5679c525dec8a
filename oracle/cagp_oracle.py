"""CPU oracle for the CaGP hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, operation by operation, the arithmetic that the reference
(sethaxen/CAGPJax, pure Python/JAX) performs on the path
``ComputationAwareGP.init -> elbo / predict`` with a ``BlockSparsePolicy`` (or a dense
action matrix), a lazily evaluated stationary kernel and the ``Cholesky(jitter)`` solver.
It is written in torch float64 on the CPU so that ``torch.autograd`` plays the role that
JAX reverse-mode plays in the reference (the reference has no custom VJP on this path).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module - always as the checker or the timed
CPU baseline, never as part of the shipped product path.

PARITY: the reference cannot be imported here as is (jax, jaxlib, gpjax, cola-ml, lineax, equinox,
paramax are absent and there is no network), and its tests hold no golden vectors (every
assertion is relational).  Since round 2 its UNMODIFIED SOURCE is executed under numpy stand-ins
for those dependencies (oracle/refshim, oracle/run_reference.py) and this oracle is pinned against
what it computes: tests/golden/reference_golden.npz (generator tests/golden/make_reference_golden.py),
checked in tests/test_reference_golden.py at 1e-12.  NOT pinned: the third-party packages
themselves - the GPJax kernel formulas (``gpjax>=0.14.0rc1``, not vendored in /root/reference) are
restated from their published definitions in both the stand-ins and this file; they are anchored on
scikit-learn's kernels and exact GP regression (tests/test_oracle_third_party_anchors.py).  Further
pins: the relational identities the reference's own tests assert (tests/test_oracle.py), finite
differences for every gradient, and agreement of the two independent formulations below.

Reference citations are relative to /root/reference/src/cagpjax unless noted.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

DT = torch.float64

KERNEL_FAMILIES = ("rbf", "matern12", "matern32", "matern52")


# --------------------------------------------------------------------------------------
# Block layout (operators/block_diagonal_sparse.py:49-52, linalg/congruence.py:29-32,
# policies/block_sparse.py:22-25)
# --------------------------------------------------------------------------------------
def block_layout(n: int, n_blocks: int):
    """Returns (block_size, n_blocks_main, n_main); the last block absorbs n % n_blocks."""
    block_size = n // n_blocks
    n_blocks_main = n_blocks if n % n_blocks == 0 else n_blocks - 1
    n_main = n_blocks_main * block_size
    return block_size, n_blocks_main, n_main


def block_starts(n: int, n_blocks: int) -> np.ndarray:
    """Start offsets of every block plus a final n (length n_blocks + 1)."""
    block_size, n_blocks_main, n_main = block_layout(n, n_blocks)
    starts = [b * block_size for b in range(n_blocks_main)]
    if n > n_main:
        starts.append(n_main)
    # n_blocks > n with n % n_blocks != 0 gives block_size 0: blocks are empty but the
    # reference's reshape arithmetic still "works"; keep the same degenerate layout.
    while len(starts) < n_blocks:
        starts.append(n)
    starts.append(n)
    return np.asarray(starts, dtype=np.int64)


def block_index(n: int, n_blocks: int) -> torch.Tensor:
    starts = block_starts(n, n_blocks)
    idx = np.searchsorted(starts[1:], np.arange(n), side="right")
    return torch.from_numpy(np.minimum(idx, n_blocks - 1))


# --------------------------------------------------------------------------------------
# BlockSparsePolicy (policies/block_sparse.py:18-40, 68-108)
# --------------------------------------------------------------------------------------
def normalize_by_blocks(values: torch.Tensor, n_actions: int) -> torch.Tensor:
    """policies/block_sparse.py:18-40 (zero-norm guard ``where(norm > 0, norm, 1)``)."""
    n = values.shape[0]
    block_size, n_blocks_main, n_main = block_layout(n, n_actions)
    chunks = []
    if n_main > 0:
        main = values[:n_main].reshape(n_blocks_main, block_size)
        norms = torch.linalg.vector_norm(main, dim=1, keepdim=True)
        main = main / torch.where(norms > 0, norms, torch.ones_like(norms))
        chunks.append(main.reshape(n_main))
    if n > n_main:
        overhang = values[n_main:]
        norm = torch.linalg.vector_norm(overhang)
        overhang = overhang / torch.where(norm > 0, norm, torch.ones_like(norm))
        chunks.append(overhang)
    if not chunks:
        return values
    return torch.cat(chunks)


# --------------------------------------------------------------------------------------
# BlockDiagonalSparse (operators/block_diagonal_sparse.py:43-95)
# --------------------------------------------------------------------------------------
def bds_matmat(nz_values: torch.Tensor, n_blocks: int, X: torch.Tensor) -> torch.Tensor:
    """S @ X with X (n_blocks, m) -> (n, m).  block_diagonal_sparse.py:48-70."""
    n = nz_values.shape[0]
    block_size, n_blocks_main, n_main = block_layout(n, n_blocks)
    m = X.shape[1]
    blocks_main = nz_values[:n_main].reshape(n_blocks_main, block_size)
    X_main = X[:n_blocks_main, :]
    res_main = (blocks_main[..., None] * X_main[:, None, :]).reshape(n_main, m)
    if n > n_main:
        X_overhang = X[n_blocks_main, :]
        block_overhang = nz_values[n_main:]
        res_overhang = torch.outer(block_overhang, X_overhang)
        return torch.cat([res_main, res_overhang], dim=0)
    return res_main


def bds_rmatmat(nz_values: torch.Tensor, n_blocks: int, X: torch.Tensor) -> torch.Tensor:
    """X @ S with X (m, n) -> (m, n_blocks).  block_diagonal_sparse.py:72-95."""
    n = nz_values.shape[0]
    block_size, n_blocks_main, n_main = block_layout(n, n_blocks)
    m = X.shape[0]
    blocks_main = nz_values[:n_main].reshape(n_blocks_main, block_size)
    X_main = X[:, :n_main].reshape(m, n_blocks_main, block_size)
    res_main = torch.einsum("ik,jik->ji", blocks_main, X_main)
    if n > n_main:
        X_overhang = X[:, n_main:]
        block_overhang = nz_values[n_main:]
        res_overhang = (X_overhang @ block_overhang)[:, None]
        return torch.cat([res_main, res_overhang], dim=1)
    return res_main


def bds_dense(nz_values: torch.Tensor, n_blocks: int) -> torch.Tensor:
    """``op @ eye`` - the reference's definition of to_dense (tests/test_operators.py:24-28)."""
    return bds_matmat(nz_values, n_blocks, torch.eye(n_blocks, dtype=nz_values.dtype))


def congruence_bds_diag(nz_values: torch.Tensor, n_blocks: int, diag) -> torch.Tensor:
    """Diagonal of S^T D S for diagonal/scalar D.  linalg/congruence.py:25-43."""
    vals = diag * nz_values**2
    n = nz_values.shape[0]
    block_size, n_blocks_main, n_main = block_layout(n, n_blocks)
    if n_blocks_main > 0:
        out = vals[:n_main].reshape(n_blocks_main, block_size).sum(dim=1)
    else:
        out = vals.new_zeros((0,))
    if n > n_main:
        out = torch.cat([out, vals[n_main:].sum(dim=0, keepdim=True)])
    return out


# --------------------------------------------------------------------------------------
# Kernel formulas (GPJax gpjax.kernels.stationary.{rbf,matern12,matern32,matern52};
# third-party, restated; call sites operators/lazy_kernel.py:49,83-85,96-98)
# --------------------------------------------------------------------------------------
def cross_covariance(family: str, x1, x2, lengthscale, variance) -> torch.Tensor:
    """Dense K(x1, x2), shape (m, n).  ``lengthscale`` scalar or (d,).

    GPJax: x/=l; y/=l; RBF  K = var * exp(-0.5 * |x-y|^2)
           tau = sqrt(max(|x-y|^2, 1e-36))
           Matern12 K = var * exp(-tau)
           Matern32 K = var * (1 + sqrt3 tau) exp(-sqrt3 tau)
           Matern52 K = var * (1 + sqrt5 tau + 5/3 tau^2) exp(-sqrt5 tau)
    """
    a = x1 / lengthscale
    b = x2 / lengthscale
    diff = a[:, None, :] - b[None, :, :]
    sq = (diff * diff).sum(dim=-1)
    if family == "rbf":
        return variance * torch.exp(-0.5 * sq)
    tau = torch.sqrt(torch.clamp(sq, min=1e-36))
    if family == "matern12":
        return variance * torch.exp(-tau)
    if family == "matern32":
        r = math.sqrt(3.0) * tau
        return variance * (1.0 + r) * torch.exp(-r)
    if family == "matern52":
        r = math.sqrt(5.0) * tau
        return variance * (1.0 + r + (5.0 / 3.0) * tau * tau) * torch.exp(-r)
    raise ValueError(f"unknown kernel family {family!r}")


# --------------------------------------------------------------------------------------
# LazyKernel (operators/lazy_kernel.py:59-103)
# --------------------------------------------------------------------------------------
def lazy_batch_sizes(shape, itemsize: int, max_memory_mb=2**10, batch_size=None):
    """(batch_size_row, batch_size_col).  lazy_kernel.py:59-77."""
    if batch_size is not None:
        return batch_size, batch_size
    max_elements = int(max_memory_mb * 2**20) // itemsize
    return max(1, max_elements // shape[1]), max(1, max_elements // shape[0])


def lazy_kernel_matmat(family, x1, x2, lengthscale, variance, X, *, max_memory_mb=2**10,
                       batch_size=None) -> torch.Tensor:
    """K(x1,x2) @ X by sequential row batches.  lazy_kernel.py:79-90."""
    vec = X.dim() == 1
    Xm = X[:, None] if vec else X
    b_row, _ = lazy_batch_sizes((x1.shape[0], x2.shape[0]), x1.element_size(), max_memory_mb, batch_size)
    out = []
    for i0 in range(0, x1.shape[0], b_row):
        k_chunk = cross_covariance(family, x1[i0:i0 + b_row], x2, lengthscale, variance)
        out.append(k_chunk @ Xm)
    res = torch.cat(out, dim=0)
    return res[:, 0] if vec else res


def lazy_kernel_rmatmat(family, x1, x2, lengthscale, variance, X, *, max_memory_mb=2**10,
                        batch_size=None) -> torch.Tensor:
    """X @ K(x1,x2) by sequential column batches.  lazy_kernel.py:92-103."""
    vec = X.dim() == 1
    Xm = X[None, :] if vec else X
    _, b_col = lazy_batch_sizes((x1.shape[0], x2.shape[0]), x1.element_size(), max_memory_mb, batch_size)
    out = []
    for j0 in range(0, x2.shape[0], b_col):
        k_chunk = cross_covariance(family, x1, x2[j0:j0 + b_col], lengthscale, variance)
        out.append(Xm @ k_chunk)
    res = torch.cat(out, dim=1)
    return res[0] if vec else res


# --------------------------------------------------------------------------------------
# Cholesky solver (linalg/lower_cholesky.py:14-43, linalg/utils.py:15-36, solvers/cholesky.py:36-70,
# solvers/base.py:84-93)
# --------------------------------------------------------------------------------------
def lower_cholesky_dense(A: torch.Tensor, jitter: Optional[float]) -> torch.Tensor:
    """chol(A + jitter*I).  lower_cholesky.py:26-38 with utils.py:15-19."""
    if jitter is not None:
        A = A + jitter * torch.eye(A.shape[0], dtype=A.dtype)
    return torch.linalg.cholesky(A)


def lower_cholesky_diag(diag: torch.Tensor, jitter: Optional[float]) -> torch.Tensor:
    """sqrt(diag + jitter) - structure-preserving overloads.  lower_cholesky.py:41-43, utils.py:32-36."""
    if jitter is not None:
        diag = diag + jitter
    return torch.sqrt(diag)


def chol_solve(L: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """L^-T (L^-1 b).  solvers/cholesky.py:46-51."""
    bm = b[:, None] if b.dim() == 1 else b
    z = torch.linalg.solve_triangular(L, bm, upper=False)
    x = torch.linalg.solve_triangular(L.T, z, upper=True)
    return x[:, 0] if b.dim() == 1 else x


def chol_logdet(L: torch.Tensor) -> torch.Tensor:
    """solvers/cholesky.py:53-55."""
    return 2 * torch.sum(torch.log(torch.diagonal(L)))


def chol_inv_congruence(L: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """(L^-1 B)^T (L^-1 B).  solvers/cholesky.py:57-63."""
    right = torch.linalg.solve_triangular(L, B, upper=False)
    return right.T @ right


def chol_inv_quad(L: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """solvers/base.py:84-93."""
    return chol_inv_congruence(L, b[:, None]).squeeze()


def chol_trace_solve(L: torch.Tensor, L_other_dense: torch.Tensor) -> torch.Tensor:
    """|| L^-1 L_other ||_F^2.  solvers/cholesky.py:65-70."""
    return torch.sum(torch.square(torch.linalg.solve_triangular(L, L_other_dense, upper=False)))


# --------------------------------------------------------------------------------------
# ComputationAwareGP (models/cagp.py:73-271) - LITERAL variant: dense S, dense n x k products
# --------------------------------------------------------------------------------------
@dataclass
class Params:
    """Differentiable leaves of the model (what ``value_and_grad(cagp)`` differentiates)."""

    lengthscale: torch.Tensor  # scalar or (d,)
    variance: torch.Tensor  # scalar
    obs_stddev: torch.Tensor  # scalar
    mean_constant: torch.Tensor  # scalar (Zero mean == constant 0, not differentiated)
    actions: torch.Tensor  # nz_values (n,) for block-sparse, or dense (n, k)

    def leaves(self):
        return [self.lengthscale, self.variance, self.obs_stddev, self.mean_constant, self.actions]

    def requires_grad_(self, flag=True):
        for t in self.leaves():
            t.requires_grad_(flag)
        return self


def make_params(lengthscale, variance, obs_stddev, mean_constant, actions) -> Params:
    f = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64)).clone() if not torch.is_tensor(v) else v.detach().to(DT).clone()
    return Params(f(lengthscale), f(variance), f(obs_stddev), f(mean_constant), f(actions))


@dataclass
class State:
    """models/cagp.py:27-46 plus the tensors the oracle needs to replay predict/elbo."""

    x: torch.Tensor
    y: torch.Tensor
    S: torch.Tensor  # dense (n, k) actions (literal to_dense of the action operator)
    obs_cov_proj: torch.Tensor  # diagonal (k,) [block-sparse] or dense (k, k)
    obs_cov_proj_is_diag: bool
    L: torch.Tensor  # cov_prior_proj_state
    residual_proj: torch.Tensor
    repr_weights_proj: torch.Tensor
    cov_prior_proj: torch.Tensor  # S^T (K + sigma^2 I) S before jitter (for inspection)


def cagp_init(family: str, x, y, p: Params, n_actions: Optional[int], jitter: Optional[float] = 1e-6,
              *, max_memory_mb=2**10, batch_size=None) -> State:
    """models/cagp.py:73-124.  ``n_actions`` None => ``p.actions`` is a dense (n, k) matrix."""
    x = torch.atleast_2d(x)
    y = torch.atleast_1d(y).squeeze()
    if y.dim() == 0:
        y = y[None]
    n = x.shape[0]
    mean_prior = p.mean_constant * torch.ones(n, dtype=DT)  # cagp.py:99-103
    sigma2 = p.obs_stddev**2  # cagp.py:105
    if n_actions is not None:
        S = bds_dense(p.actions, n_actions)  # to_dense of BlockDiagonalSparse
        obs_cov_proj = congruence_bds_diag(p.actions, n_actions, sigma2)  # cagp.py:110
        is_diag = True
    else:
        S = p.actions
        obs_cov_proj = S.T @ (sigma2 * S)  # congruence fallback, congruence.py:15-17
        is_diag = False
    # cov_prior_proj = S^T ((K + sigma^2 I) S)  (cagp.py:111 -> lower_cholesky.py:36-38 to_dense)
    KS = lazy_kernel_matmat(family, x, x, p.lengthscale, p.variance, S,
                            max_memory_mb=max_memory_mb, batch_size=batch_size)
    cov_prior_S = KS + sigma2 * S
    cov_prior_proj = S.T @ cov_prior_S
    L = lower_cholesky_dense(cov_prior_proj, jitter)  # cagp.py:112
    residual_proj = S.T @ (y - mean_prior)  # cagp.py:114
    w = chol_solve(L, residual_proj)  # cagp.py:115
    return State(x, y, S, obs_cov_proj, is_diag, L, residual_proj, w, cov_prior_proj)


def cagp_predict(family: str, st: State, p: Params, z: Optional[torch.Tensor] = None, *, full_cov=False,
                 max_memory_mb=2**10, batch_size=None):
    """models/cagp.py:126-169 + distributions.py:48-56.  Returns (mean, variance[, covariance])."""
    x = st.x
    zz = x if z is None else torch.atleast_2d(z)
    m = zz.shape[0]
    mean_z = p.mean_constant * torch.ones(m, dtype=DT)
    # cov_zx_proj = K(z, x) @ S   (noise-free kernel; cagp.py:154-160)
    cov_zx_proj = lazy_kernel_matmat(family, zz, x, p.lengthscale, p.variance, st.S,
                                     max_memory_mb=max_memory_mb, batch_size=batch_size)
    mean_pred = mean_z + cov_zx_proj @ st.repr_weights_proj  # cagp.py:163
    R = torch.linalg.solve_triangular(st.L, cov_zx_proj.T, upper=False)  # L^-1 (K S)^T
    # diag of K(z,z) for the stationary families is `variance` (kernel at zero distance; matern uses
    # tau = sqrt(1e-36) = 1e-18, indistinguishable from 0 in float64)
    kzz_diag = cross_covariance(family, zz[:1], zz[:1], p.lengthscale, p.variance)[0, 0] * torch.ones(m, dtype=DT)
    var_pred = kzz_diag - (R * R).sum(dim=0)  # cola.diag(K_zz - R^T R)
    if full_cov:
        cov = cross_covariance(family, zz, zz, p.lengthscale, p.variance) - R.T @ R
        return mean_pred, var_pred, cov
    return mean_pred, var_pred


def cagp_prior_kl(st: State, jitter: Optional[float] = 1e-6) -> torch.Tensor:
    """models/cagp.py:171-201 and 252-271."""
    k = st.residual_proj.shape[0]
    if st.obs_cov_proj_is_diag:
        Lq_diag = lower_cholesky_diag(st.obs_cov_proj, jitter)  # solver.init(Diagonal)
        Lq = torch.diag(Lq_diag)
        logdet_q = 2 * torch.sum(torch.log(Lq_diag))
        wDw = torch.sum(st.repr_weights_proj**2 * st.obs_cov_proj)
    else:
        Lq = lower_cholesky_dense(st.obs_cov_proj, jitter)
        logdet_q = chol_logdet(Lq)
        wDw = st.repr_weights_proj @ (st.obs_cov_proj @ st.repr_weights_proj)
    diff = torch.zeros_like(st.residual_proj) - st.residual_proj  # mean_p - mean_q
    inner = chol_trace_solve(st.L, Lq)
    mahalanobis = chol_inv_quad(st.L, diff)
    logdet_ratio = chol_logdet(st.L) - logdet_q
    kl = 0.5 * (inner + logdet_ratio + mahalanobis - k)
    return kl - 0.5 * wDw


def expected_log_likelihood(y, mean, variance, obs_stddev) -> torch.Tensor:
    """GPJax ``Gaussian.expected_log_likelihood`` (analytical): per point
    -0.5 * (log(2 pi) + log(sigma^2) + ((y - m)^2 + v) / sigma^2).  Call site cagp.py:227-229."""
    sq = obs_stddev**2
    return -0.5 * (math.log(2.0 * math.pi) + torch.log(sq) + ((y - mean) ** 2 + variance) / sq)


def cagp_variational_expectation(family, st: State, p: Params, **kw) -> torch.Tensor:
    """models/cagp.py:203-231."""
    mean, var = cagp_predict(family, st, p, None, **kw)
    return expected_log_likelihood(st.y, mean, var, p.obs_stddev)


def cagp_elbo(family, st: State, p: Params, jitter=1e-6, **kw) -> torch.Tensor:
    """models/cagp.py:233-249."""
    return torch.sum(cagp_variational_expectation(family, st, p, **kw)) - cagp_prior_kl(st, jitter)


def elbo_literal(family, x, y, p: Params, n_actions, jitter=1e-6, **kw) -> torch.Tensor:
    """``cagp.elbo(cagp.init(train_data))`` - the objective of README.md:57-59 (sign not flipped)."""
    st = cagp_init(family, x, y, p, n_actions, jitter, **kw)
    return cagp_elbo(family, st, p, jitter, **kw)


def elbo_and_grad_literal(family, x, y, p: Params, n_actions, jitter=1e-6, **kw):
    """value_and_grad of the ELBO w.r.t. every leaf of ``Params`` via torch reverse mode."""
    q = Params(*[t.detach().clone().requires_grad_(True) for t in p.leaves()])
    val = elbo_literal(family, x, y, q, n_actions, jitter, **kw)
    grads = torch.autograd.grad(val, q.leaves(), allow_unused=True)
    grads = [torch.zeros_like(t) if g is None else g for g, t in zip(grads, q.leaves())]
    names = ("lengthscale", "variance", "obs_stddev", "mean_constant", "actions")
    return val.detach(), dict(zip(names, grads))


# --------------------------------------------------------------------------------------
# STRUCTURED variant (independent formulation; numpy, chunked; O(n^2) work, O(n k) memory).
# Never densifies S.  SURVEY.md section 3.2.  Used for mid-size parity (n ~ 1e4..1e5).
# --------------------------------------------------------------------------------------
def _np_kernel_unit(family, xa, xb):
    """Unit-variance kernel and (optionally needed) squared scaled distance, numpy."""
    sq = ((xa[:, None, :] - xb[None, :, :]) ** 2).sum(-1)
    if family == "rbf":
        return np.exp(-0.5 * sq), sq
    tau = np.sqrt(np.maximum(sq, 1e-36))
    if family == "matern12":
        return np.exp(-tau), sq
    if family == "matern32":
        r = math.sqrt(3.0) * tau
        return (1.0 + r) * np.exp(-r), sq
    if family == "matern52":
        r = math.sqrt(5.0) * tau
        return (1.0 + r + (5.0 / 3.0) * tau * tau) * np.exp(-r), sq
    raise ValueError(family)


def projected_kernel_blocksparse(family, x1, x2, lengthscale, s, n_blocks, *, chunk=1024) -> np.ndarray:
    """M' = K'(x1,x2) S with unit variance, shape (m, k): segment sums over the column blocks."""
    x1 = np.asarray(x1, dtype=np.float64) / np.asarray(lengthscale, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64) / np.asarray(lengthscale, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    starts = block_starts(x2.shape[0], n_blocks)
    out = np.empty((x1.shape[0], n_blocks))
    nonempty = starts[:-1] < starts[1:]
    for i0 in range(0, x1.shape[0], chunk):
        Kc, _ = _np_kernel_unit(family, x1[i0:i0 + chunk], x2)
        Kc *= s[None, :]
        seg = np.zeros((Kc.shape[0], n_blocks))
        seg[:, nonempty] = np.add.reduceat(Kc, starts[:-1][nonempty], axis=1)
        out[i0:i0 + chunk] = seg
    return out


def structured_stats(family, x, y, lengthscale, mean_constant, s, n_blocks, *, chunk=1024):
    """Row-sum-reducible statistics of SURVEY 3.2 (unit variance): A', B', c', r, q, yy and M'."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    s = np.asarray(s, dtype=np.float64)
    yt = np.asarray(y, dtype=np.float64).reshape(-1) - float(mean_constant)
    starts = block_starts(n, n_blocks)
    ne = starts[:-1] < starts[1:]
    M = projected_kernel_blocksparse(family, x, x, lengthscale, s, n_blocks, chunk=chunk)
    A = np.zeros((n_blocks, n_blocks))
    A[ne] = np.add.reduceat(s[:, None] * M, starts[:-1][ne], axis=0)
    B = M.T @ M
    c = M.T @ yt
    r = np.zeros(n_blocks); q = np.zeros(n_blocks)
    r[ne] = np.add.reduceat(s * yt, starts[:-1][ne])
    q[ne] = np.add.reduceat(s * s, starts[:-1][ne])
    return dict(A=A, B=B, c=c, r=r, q=q, yy=float(yt @ yt), n=n, M=M)


def elbo_from_stats(st: dict, variance, obs_stddev, jitter=1e-6, kdiag_unit=1.0):
    """Closed form of SURVEY 3.2 on k-sized tensors (torch, differentiable w.r.t. every input)."""
    T = lambda v: v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v, dtype=np.float64))
    A, B, c, r, q = T(st["A"]), T(st["B"]), T(st["c"]), T(st["r"]), T(st["q"])
    yy, n = T(st["yy"]), st["n"]
    variance, obs_stddev = T(variance), T(obs_stddev)
    k = A.shape[0]
    sig2 = obs_stddev**2
    dvec = sig2 * q
    eps = 0.0 if jitter is None else jitter
    P = variance * A + torch.diag(dvec) + eps * torch.eye(k, dtype=DT)
    L = torch.linalg.cholesky(P)
    w = chol_solve(L, r)
    Bv = variance**2 * B
    cv = variance * c
    tr_PinvB = torch.sum(torch.linalg.solve_triangular(L, Bv, upper=False) * torch.linalg.solve_triangular(L, torch.eye(k, dtype=DT), upper=False))
    sum_var = n * variance * kdiag_unit - tr_PinvB
    resid2 = yy - 2.0 * (w @ cv) + w @ (Bv @ w)
    var_exp = -0.5 * n * (math.log(2.0 * math.pi) + torch.log(sig2)) - 0.5 * (resid2 + sum_var) / sig2
    dj = dvec + eps
    inner = torch.sum(torch.square(torch.linalg.solve_triangular(L, torch.diag(torch.sqrt(dj)), upper=False)))
    kl = 0.5 * (inner + chol_logdet(L) - torch.sum(torch.log(dj)) + r @ w - k) - 0.5 * torch.sum(w * w * dvec)
    return var_exp - kl


def elbo_structured(family, x, y, lengthscale, variance, obs_stddev, mean_constant, s, n_blocks,
                    jitter=1e-6, *, chunk=1024) -> float:
    st = structured_stats(family, x, y, lengthscale, mean_constant, s, n_blocks, chunk=chunk)
    return float(elbo_from_stats(st, variance, obs_stddev, jitter))


# --------------------------------------------------------------------------------------
# Textbook exact GP (independent anchor for tests/test_models.py:206-235 of the reference)
# --------------------------------------------------------------------------------------
def exact_gp_predict(family, x, y, z, lengthscale, variance, obs_stddev, mean_constant):
    x = torch.atleast_2d(x); z = torch.atleast_2d(z)
    Kxx = cross_covariance(family, x, x, lengthscale, variance) + obs_stddev**2 * torch.eye(x.shape[0], dtype=DT)
    Kzx = cross_covariance(family, z, x, lengthscale, variance)
    Kzz = cross_covariance(family, z, z, lengthscale, variance)
    Lc = torch.linalg.cholesky(Kxx)
    alpha = torch.cholesky_solve((y.reshape(-1) - mean_constant)[:, None], Lc)[:, 0]
    V = torch.linalg.solve_triangular(Lc, Kzx.T, upper=False)
    return mean_constant + Kzx @ alpha, Kzz - V.T @ V
