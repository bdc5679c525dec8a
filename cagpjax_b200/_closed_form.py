"""k-sized closed form of the CaGP ELBO and its hand-derived reverse mode.

Given the fused statistics of the forward pass (A' = S^T K' S, B' = M'^T M', c' = M'^T yt with the
unit-variance kernel K'), r = S^T yt, q_b = sum_{i in b} s_i^2, yy = |yt|^2 and the scalars
(variance v, obs_stddev sigma, jitter eps), the ELBO of models/cagp.py:233-271 is

    P   = v A' + diag(sigma^2 q + eps),  L = chol(P),  Q = P^-1,  w = Q r
    ELBO = -n/2 log(2 pi sigma^2) - [yy - 2 v w.c' + v^2 w.B'w + n v - v^2 tr(Q B')] / (2 sigma^2)
           - 1/2 [tr(Q Dj) + logdet P - sum log dj + r.w - k] + 1/2 w.D w,     D = sigma^2 diag(q), Dj = D + eps

(SURVEY 3.2).  Differentiating this with generic autograd launches ~1000 tiny kernels (Cholesky and
triangular-solve backward passes); the k x k work is replicated on every rank, so it is the serial
fraction that limits multi-GPU scaling.  Here the forward is one potrf + one potri + a few GEMV/GEMMs and
the backward is four k x k GEMMs, all on the device, with the cotangents written in closed form:

    Pbar = -(w h^T + h w^T)/(2 s2) + (w z^T + z w^T)/(2 s2) - Q Bv Q/(2 s2) + Q Dj Q/2 - Q/2 + w w^T/2
           - (w u^T + u w^T)/2,        h = Q cv, z = Q Bv w, u = Q D w, Bv = v^2 B', cv = v c', s2 = sigma^2
    Bvbar = (Q - w w^T)/(2 s2),  cvbar = w/s2,  rbar = (h - z)/s2 - w + u,
    dvecbar_b = Pbar_bb - Q_bb/2 + 1/(2 dj_b) + w_b^2/2,   yybar = -1/(2 s2)

``elbo_from_stats_autograd`` is the plain-autograd version of the same formula; tests check the two
against each other and against the oracle.
"""

from __future__ import annotations

import math

import torch


def elbo_from_stats_autograd(A, B, c, r, q, yy, n, variance, sigma, jitter):
    """Reference implementation with generic autograd (used by tests and as documentation)."""
    k = A.shape[0]
    s2 = sigma**2
    dvec = s2 * q
    dj = dvec + jitter
    P = variance * A + torch.diag(dj)
    L = torch.linalg.cholesky(P)
    w = torch.cholesky_solve(r[:, None], L)[:, 0]
    Bv = variance**2 * B
    cv = variance * c
    tr_QB = torch.diagonal(torch.cholesky_solve(Bv, L)).sum()
    resid2 = yy - 2.0 * (w @ cv) + w @ (Bv @ w)
    var_exp = -0.5 * n * (math.log(2.0 * math.pi) + torch.log(s2)) - 0.5 * (resid2 + n * variance - tr_QB) / s2
    Linv_sqrtD = torch.linalg.solve_triangular(L, torch.diag(torch.sqrt(dj)), upper=False)
    kl = 0.5 * (torch.sum(Linv_sqrtD**2) + 2.0 * torch.sum(torch.log(torch.diagonal(L))) - torch.sum(torch.log(dj))
                + r @ w - k) - 0.5 * torch.sum(w * w * dvec)
    return var_exp - kl


class ElboFromStats(torch.autograd.Function):
    """ELBO(A', B', c', r, q, yy, variance, sigma) with the closed-form reverse mode above."""

    @staticmethod
    def forward(ctx, A, B, c, r, q, yy, variance, sigma, n, jitter, chol=None):
        k = A.shape[0]
        s2 = sigma * sigma
        dvec = s2 * q
        dj = dvec + jitter
        if chol is not None:
            # the factor of exactly this P that ``init`` already computed (solvers/cholesky.py: init): the k^3
            # factorisation is replicated on every rank, so it is paid once per step, not twice
            L = chol
        else:
            P = variance * A
            P.diagonal().add_(dj)
            # cholesky_ex without error checking: no host synchronisation (jnp.linalg.cholesky, which the
            # reference uses, also returns NaNs instead of raising when P is not positive definite)
            L = torch.linalg.cholesky_ex(P, check_errors=False).L
        Q = torch.cholesky_inverse(L)
        w = Q @ r
        Bv = (variance * variance) * B
        cv = variance * c
        Bw = Bv @ w
        T1 = w @ cv
        T2 = w @ Bw
        T3 = torch.sum(Q * Bv)
        Qdiag = torch.diagonal(Q)
        T4 = torch.sum(Qdiag * dj)
        T5 = 2.0 * torch.sum(torch.log(torch.diagonal(L)))
        T6 = r @ w
        T7 = torch.sum(w * w * dvec)
        resid = yy - 2.0 * T1 + T2 + n * variance - T3
        elbo = -0.5 * n * (math.log(2.0 * math.pi) + torch.log(s2)) - 0.5 * resid / s2 \
            - 0.5 * (T4 + T5 - torch.sum(torch.log(dj)) + T6 - k) + 0.5 * T7
        ctx.save_for_backward(A, B, c, q, Q, w, Bv, cv, Bw, dvec, dj, variance, sigma, resid)
        ctx.n = n
        return elbo

    @staticmethod
    def backward(ctx, g):
        A, B, c, q, Q, w, Bv, cv, Bw, dvec, dj, variance, sigma, resid = ctx.saved_tensors
        n = ctx.n
        s2 = sigma * sigma
        h = Q @ cv
        z = Q @ Bw
        u = Q @ (dvec * w)
        QB = Q @ Bv
        # Pbar (symmetric)
        a = (z - h) / (2.0 * s2) - 0.5 * u + 0.25 * w            # coefficient vector paired with w (rank-2 part)
        Pbar = torch.outer(w, a)
        Pbar = Pbar + Pbar.T
        Pbar -= (QB @ Q) / (2.0 * s2)
        Pbar += 0.5 * (Q * dj[None, :]) @ Q
        Pbar -= 0.5 * Q
        Bvbar = (Q - torch.outer(w, w)) / (2.0 * s2)
        cvbar = w / s2
        rbar = (h - z) / s2 - w + u
        dvecbar = torch.diagonal(Pbar) - 0.5 * torch.diagonal(Q) + 0.5 / dj + 0.5 * w * w
        yybar = -0.5 / s2
        var_bar = torch.sum(Pbar * A) + 2.0 * variance * torch.sum(Bvbar * B) + cvbar @ c - 0.5 * n / s2
        s2_bar = -0.5 * n / s2 + 0.5 * resid / (s2 * s2) + dvecbar @ q
        sigma_bar = 2.0 * sigma * s2_bar
        return (g * variance * Pbar, g * (variance * variance) * Bvbar, g * variance * cvbar, g * rbar,
                g * s2 * dvecbar, g * yybar, g * var_bar, g * sigma_bar, None, None, None)


def elbo_from_stats(A, B, c, r, q, yy, n, variance, sigma, jitter, chol=None):
    """``chol``: optional lower Cholesky factor of P = v A' + diag(sigma^2 q + jitter) (detached); its dependence
    on the inputs is part of the closed-form reverse mode, so it carries no autograd edge."""
    return ElboFromStats.apply(A, B, c, r, q, yy, variance, sigma, n, 0.0 if jitter is None else float(jitter),
                               None if chol is None else chol.detach())


def elbo_from_stats_dense(A, B, c, r, Qs, yy, n, variance, sigma, jitter):
    """Same closed form for a DENSE action matrix: the projected noise covariance D = sigma^2 S^T S is a
    dense k x k matrix (``Qs`` = S^T S), cf. models/cagp.py:185-199 with a dense ``obs_cov_proj``.  Plain
    autograd on k-sized tensors (k is a few hundred on this path)."""
    k = A.shape[0]
    s2 = sigma**2
    eye = torch.eye(k, dtype=A.dtype, device=A.device)
    D = s2 * Qs
    Dj = D + jitter * eye
    P = variance * A + Dj
    L = torch.linalg.cholesky_ex(P, check_errors=False).L
    w = torch.cholesky_solve(r[:, None], L)[:, 0]
    Bv = variance**2 * B
    cv = variance * c
    tr_QB = torch.diagonal(torch.cholesky_solve(Bv, L)).sum()
    resid2 = yy - 2.0 * (w @ cv) + w @ (Bv @ w)
    var_exp = -0.5 * n * (math.log(2.0 * math.pi) + torch.log(s2)) - 0.5 * (resid2 + n * variance - tr_QB) / s2
    Lq = torch.linalg.cholesky_ex(Dj, check_errors=False).L
    inner = torch.sum(torch.linalg.solve_triangular(L, Lq, upper=False) ** 2)
    kl = 0.5 * (inner + 2.0 * torch.sum(torch.log(torch.diagonal(L))) - 2.0 * torch.sum(torch.log(torch.diagonal(Lq)))
                + r @ w - k) - 0.5 * (w @ (D @ w))
    return var_exp - kl
