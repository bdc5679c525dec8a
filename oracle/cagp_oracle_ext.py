"""CPU oracle for the rows SURVEY 8(f) ranks 3-4 add around the hot path.  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch float64 on the CPU (scalar / column loops, no shared code with the product):
  * ``PseudoInputPolicy`` actions (policies/pseudoinput.py:68-74),
  * ``orthogonalize`` by Householder QR / classical / modified Gram-Schmidt (linalg/orthogonalize.py:113-196),
  * the Lanczos eigendecomposition behind ``LanczosPolicy`` (policies/lanczos.py:31-46, linalg/eigh.py:113-138),
  * the ``PseudoInverse`` solver (solvers/pseudoinverse.py:62-143),
  * the literal ``ComputationAwareGP`` init -> predict -> ELBO chain (models/cagp.py:73-271) for a DENSE action
    matrix with either solver.

Only ``tests/`` may import this module.

PARITY: same situation as ``cagp_oracle.py`` - the reference cannot be imported here as is; since round 2 its
unmodified source runs under the numpy stand-ins of oracle/refshim and this file is pinned against what it computes
(tests/golden/reference_golden_ext.npz from tests/golden/make_reference_golden_ext.py, checked in
tests/test_reference_golden.py): ``orthogonalize`` (QR / CGS / MGS), pseudo-input actions, and the dense-action model
chain with either solver.  NOT pinned: the Lanczos recurrence (owned by matfree, not vendored) and the third-party
packages themselves.  Further pins: the relational identities of tests/test_solvers.py, tests/test_linalg.py and
tests/test_policies.py of the reference (re-expressed in tests/test_oracle_ext.py), dense ``eigh`` / ``lstsq``
cross-checks and finite
differences.  Reference citations are relative to /root/reference/src/cagpjax.
"""

from __future__ import annotations

import math
from typing import Optional

import torch

from . import cagp_oracle as base

DT = torch.float64


# --------------------------------------------------------------------------------------
# PseudoInputPolicy (policies/pseudoinput.py:68-74)
# --------------------------------------------------------------------------------------
def pseudo_input_actions(family: str, x, z, lengthscale, variance) -> torch.Tensor:
    """S = kernel.cross_covariance(train_inputs, pseudo_inputs), dense (n, M)."""
    return base.cross_covariance(family, torch.atleast_2d(x), torch.atleast_2d(z), lengthscale, variance)


# --------------------------------------------------------------------------------------
# orthogonalize (linalg/orthogonalize.py:113-196)
# --------------------------------------------------------------------------------------
def _l2_normalize(x: torch.Tensor, rtol: float = 0.0) -> torch.Tensor:
    r = torch.sqrt(torch.sum(x * x))
    if float(r.detach()) <= rtol * x.numel():
        return x.detach()
    return x / r


def orthogonalize_qr(A: torch.Tensor, n_reortho: int = 0) -> torch.Tensor:
    Q = A
    for _ in range(n_reortho + 1):
        Q = torch.linalg.qr(Q)[0]
    return Q


def orthogonalize_cgs(A: torch.Tensor, n_reortho: int = 0) -> torch.Tensor:
    """linalg/orthogonalize.py:113-141: column k minus its projections on the finalised columns (all computed from
    the same vector per sweep), then normalised with cutoff rtol = 10 eps."""
    rtol = float(torch.finfo(A.dtype).eps * 10)
    cols = [A[:, j] for j in range(A.shape[1])]
    for k in range(len(cols)):
        q = cols[k]
        for _ in range(n_reortho + 1):
            coeffs = [torch.dot(cols[j], q) for j in range(k)]
            for j in range(k):
                q = q - coeffs[j] * cols[j]
        cols[k] = _l2_normalize(q, rtol)
    return torch.stack(cols, dim=1)


def orthogonalize_mgs(A: torch.Tensor, n_reortho: int = 0) -> torch.Tensor:
    """linalg/orthogonalize.py:144-182: normalise column k (after ``n_reortho`` sequential sweeps against the
    finalised columns), then remove its component from every later column."""
    cols = [A[:, j] for j in range(A.shape[1])]
    for k in range(len(cols)):
        q = cols[k]
        for _ in range(n_reortho):
            for j in range(k):
                q = q - torch.dot(cols[j], q) * cols[j]
        q = _l2_normalize(q)
        cols[k] = q
        for j in range(k + 1, len(cols)):
            cols[j] = cols[j] - torch.dot(cols[j], q) * q
    return torch.stack(cols, dim=1)


# --------------------------------------------------------------------------------------
# Lanczos eigendecomposition (linalg/eigh.py:113-138; recurrence of matfree.decomp.tridiag_sym with
# reortho="full", third-party, restated from the textbook algorithm)
# --------------------------------------------------------------------------------------
def lanczos_eigh(A: torch.Tensor, num_matvecs: int, v0: torch.Tensor):
    """Returns (eigenvalues (m,), eigenvectors (n, m)) of the Lanczos approximation after ``num_matvecs`` steps."""
    n = A.shape[0]
    q = v0 / torch.sqrt(torch.sum(v0 * v0))
    Q = [q]
    alpha, beta = [], []
    for i in range(num_matvecs):
        w = A @ Q[i]
        a = torch.dot(Q[i], w)
        w = w - a * Q[i]
        if i > 0:
            w = w - beta[i - 1] * Q[i - 1]
        for qj in list(Q):  # full re-orthogonalisation, one vector at a time
            w = w - torch.dot(qj, w) * qj
        alpha.append(a)
        if i + 1 < num_matvecs:
            b = torch.sqrt(torch.sum(w * w))
            beta.append(b)
            Q.append(w / b if float(b.detach()) > 0 else torch.zeros_like(w))
    m = num_matvecs
    H = torch.zeros((m, m), dtype=A.dtype)
    for i in range(m):
        H[i, i] = alpha[i]
    for i in range(m - 1):
        H[i, i + 1] = beta[i]
        H[i + 1, i] = beta[i]
    vals, vecs = torch.linalg.eigh(H)
    return vals, torch.stack(Q, dim=1) @ vecs


# --------------------------------------------------------------------------------------
# PseudoInverse solver (solvers/pseudoinverse.py:62-143)
# --------------------------------------------------------------------------------------
class PinvState:
    def __init__(self, A: torch.Tensor, rtol: Optional[float] = None):
        n = A.shape[0]
        rtol_val = rtol if rtol is not None else float(torch.finfo(A.dtype).eps) * n
        self.A = A
        self.vals, self.vecs = torch.linalg.eigh(0.5 * (A + A.T))
        cutoff = rtol_val * torch.max(torch.abs(self.vals))
        self.mask = self.vals >= cutoff
        self.safe = torch.where(self.mask, self.vals, torch.ones_like(self.vals))
        self.inv = torch.where(self.mask, 1.0 / self.safe, torch.zeros_like(self.vals))


def pinv_solve(st: PinvState, b: torch.Tensor) -> torch.Tensor:
    return st.vecs @ ((st.vecs.T @ b).T * st.inv).T if b.dim() == 2 else st.vecs @ ((st.vecs.T @ b) * st.inv)


def pinv_logdet(st: PinvState) -> torch.Tensor:
    return torch.sum(torch.log(st.safe))


def pinv_inv_quad(st: PinvState, b: torch.Tensor) -> torch.Tensor:
    z = st.vecs.T @ b
    return torch.sum(z * z * st.inv)


def pinv_inv_congruence(st: PinvState, B: torch.Tensor) -> torch.Tensor:
    z = st.vecs.T @ B
    return z.T @ (st.inv[:, None] * z)


def pinv_trace_solve(st: PinvState, other: PinvState) -> torch.Tensor:
    """trace(A^+ B) with B the operator behind ``other`` (solvers/pseudoinverse.py:124-143, dense branch)."""
    return torch.trace(st.vecs @ (st.inv[:, None] * (st.vecs.T @ other.A)))


def pinv_unwhiten(st: PinvState, z: torch.Tensor) -> torch.Tensor:
    sq = torch.where(st.mask, torch.sqrt(st.safe), torch.zeros_like(st.safe))
    return st.vecs @ (sq[:, None] * z) if z.dim() == 2 else st.vecs @ (sq * z)


# --------------------------------------------------------------------------------------
# ComputationAwareGP with a dense action matrix and a choice of solver (models/cagp.py:73-271)
# --------------------------------------------------------------------------------------
def elbo_dense_actions(family: str, x, y, lengthscale, variance, obs_stddev, mean_constant, S: torch.Tensor, *,
                       solver: str = "cholesky", jitter: Optional[float] = 1e-6, rtol: Optional[float] = None,
                       z: Optional[torch.Tensor] = None):
    """init -> predict(train inputs) -> ELBO, literally, for dense actions ``S`` (n, k).

    Returns (elbo, mean_pred, var_pred) where the predictions are at ``z`` (training inputs if None)."""
    x = torch.atleast_2d(x)
    y = torch.atleast_1d(y).squeeze()
    n, k = S.shape
    mean_prior = mean_constant * torch.ones(n, dtype=DT)
    sigma2 = obs_stddev**2
    K = base.cross_covariance(family, x, x, lengthscale, variance)
    obs_cov_proj = S.T @ (sigma2 * S)                       # cagp.py:110 (congruence fallback)
    cov_prior_proj = S.T @ (K @ S + sigma2 * S)             # cagp.py:111
    residual_proj = S.T @ (y - mean_prior)                  # cagp.py:114
    if solver == "cholesky":
        Lp = base.lower_cholesky_dense(cov_prior_proj, jitter)
        Lq = base.lower_cholesky_dense(obs_cov_proj, jitter)
        solve_p = lambda b: base.chol_solve(Lp, b)
        inv_cong_p = lambda B: base.chol_inv_congruence(Lp, B)
        logdet_p, logdet_q = base.chol_logdet(Lp), base.chol_logdet(Lq)
        trace_pq = base.chol_trace_solve(Lp, Lq)
    elif solver == "pinv":
        sp, sq = PinvState(cov_prior_proj, rtol), PinvState(obs_cov_proj, rtol)
        solve_p = lambda b: pinv_solve(sp, b)
        inv_cong_p = lambda B: pinv_inv_congruence(sp, B)
        logdet_p, logdet_q = pinv_logdet(sp), pinv_logdet(sq)
        trace_pq = pinv_trace_solve(sp, sq)
    else:
        raise ValueError(solver)
    w = solve_p(residual_proj)                              # cagp.py:115

    def predict(zz):
        m = zz.shape[0]
        KzS = base.cross_covariance(family, zz, x, lengthscale, variance) @ S   # cagp.py:154-160 (noise-free)
        mean = mean_constant * torch.ones(m, dtype=DT) + KzS @ w                # cagp.py:163
        kdiag = base.cross_covariance(family, zz[:1], zz[:1], lengthscale, variance)[0, 0]
        var = kdiag - torch.diagonal(inv_cong_p(KzS.T))                         # cagp.py:164-167
        return mean, var

    mean_x, var_x = predict(x)
    ve = base.expected_log_likelihood(y, mean_x, var_x, obs_stddev)             # cagp.py:203-231
    diff = -residual_proj
    mahalanobis = torch.dot(diff, solve_p(diff))
    kl = 0.5 * (trace_pq + logdet_p - logdet_q + mahalanobis - k)               # cagp.py:252-271
    kl = kl - 0.5 * torch.dot(w, obs_cov_proj @ w)                              # cagp.py:199-201
    elbo = torch.sum(ve) - kl
    if z is not None:
        mean_z, var_z = predict(torch.atleast_2d(z))
        return elbo, mean_z, var_z
    return elbo, mean_x, var_x
