"""Computation-aware Gaussian process model (reference: models/cagp.py:27-271).

Same public surface as the reference - ``ComputationAwareGP(posterior, policy, solver)`` with
``init / predict / prior_kl / variational_expectation / elbo`` - and the same sequence of operator
calls inside ``init`` and ``predict``.  What changes is what those operators execute: with a
``BlockSparsePolicy`` and a ``LazyKernelComputation`` kernel, ONE fused n^2 pass on the GPU produces
every statistic ``init``, ``predict()`` at the training inputs and ``elbo`` need, and
its reverse mode is ONE fused backward pass.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

from .. import _native
from ..distributions import GaussianDistribution
from ..gpx import ConjugatePosterior, Constant, Dataset, unwrap
from ..interop import lazify
from ..linalg import ProjectedKernelGram, congruence_transform
from ..operators import BlockDiagonalSparse, LazyKernel, diag_like
from ..operators.lazy_kernel import KernelActions, KernelDenseActions, _kernel_params
from ..ops import PSD, Dense, LinearOperator, Triangular, diag
from ..policies import AbstractBatchLinearSolverPolicy
from ..solvers import AbstractLinearSolver, Cholesky


@dataclass
class ComputationAwareGPState:
    """Projected quantities for computation-aware GP inference (models/cagp.py:27-46)."""

    train_data: Dataset
    actions: LinearOperator
    obs_cov_proj: LinearOperator
    cov_prior_proj_state: Any
    residual_proj: torch.Tensor
    repr_weights_proj: torch.Tensor
    # B200 additions (not part of the reference state): the fused projection and the residual
    _projection: Any = field(default=None, repr=False)
    _residual: Optional[torch.Tensor] = field(default=None, repr=False)
    _cov_xx: Any = field(default=None, repr=False)  # the Gram operator of init (memoises K S products)


class PredictiveCovariance(LinearOperator):
    """Lazy ``K_zz - (K_zx S) P^-1 (K_zx S)^T`` (models/cagp.py:164-167).  The m x m matrix is only
    formed by ``to_dense``; the diagonal comes from the k-major projection with one GEMM."""

    def __init__(self, cov_zz, cov_zx_proj, solver, solver_state):
        super().__init__(cov_zz.dtype, cov_zz.shape, device=cov_zz.device)
        self.cov_zz = cov_zz
        self.cov_zx_proj = cov_zx_proj
        self.solver = solver
        self.solver_state = solver_state

    def _matmat(self, X):
        KS = self.cov_zx_proj
        return self.cov_zz @ X - KS @ self.solver.solve(self.solver_state, KS.T @ X)

    def _rmatmat(self, X):
        return self._matmat(X.T).T

    @property
    def T(self):
        return self

    def to_dense(self):
        KSt = self.cov_zx_proj.to_dense().T
        return self.cov_zz.to_dense() - self.solver.inv_congruence_transform(self.solver_state, KSt)

    def _diag(self):
        KS, st = self.cov_zx_proj, self.solver_state
        if isinstance(KS, KernelActions) and isinstance(st, Triangular):
            Mt = KS.Mt()
            var = KS._variance()
            if not (torch.is_grad_enabled() and (Mt.requires_grad or st.A.requires_grad or var.requires_grad)):
                k = Mt.shape[0]
                Pinv = torch.cholesky_inverse(st.A)
                zero = torch.zeros(k, dtype=Mt.dtype, device=Mt.device)
                _, v = _native.predict_moments(Mt.contiguous(), zero, Pinv.contiguous(), var, zero[0])
                return v
            R = torch.linalg.solve_triangular(st.A, var * Mt, upper=False)
            return diag(self.cov_zz) - (R * R).sum(dim=0)
        KSt = KS.to_dense().T  # (k, m); never form the m x m product
        if isinstance(st, Triangular):
            R = torch.linalg.solve_triangular(st.A, KSt, upper=not st.lower)
            return diag(self.cov_zz) - (R * R).sum(dim=0)
        return diag(self.cov_zz) - (KSt * self.solver.solve(st, KSt)).sum(dim=0)


class ComputationAwareGP:
    """Computation-aware Gaussian process (single output).

    Attributes:
        posterior: the exact (conjugate) posterior: prior (mean function + kernel) x Gaussian likelihood.
        policy: batch linear solver policy defining the subspace the data is projected to.
        solver: linear solver for the projected k x k systems; default ``Cholesky(1e-6)``.
    """

    def __init__(self, posterior: ConjugatePosterior, policy: AbstractBatchLinearSolverPolicy,
                 solver: Optional[AbstractLinearSolver] = None):
        self.posterior = posterior
        self.policy = policy
        self.solver = solver if solver is not None else Cholesky(1e-6)

    # ------------------------------------------------------------------------------------------
    def _mean(self, x: torch.Tensor) -> torch.Tensor:
        prior = self.posterior.prior
        mean = prior.mean_function(x).squeeze(-1)
        if isinstance(prior.mean_function, Constant):
            constant = unwrap(prior.mean_function.constant)
            mean = mean.to(constant.dtype)  # cagp.py:100-103
        return mean.to(x.device)

    def init(self, train_data: Dataset, *, key=None) -> ComputationAwareGPState:
        """Condition on the training data (models/cagp.py:73-124)."""
        if train_data.X is None or train_data.y is None:
            raise ValueError("Training data must be supervised.")
        x = torch.atleast_2d(train_data.X)
        y = torch.atleast_1d(train_data.y).squeeze()
        if y.dim() == 0:
            y = y[None]
        prior = self.posterior.prior
        likelihood = self.posterior.likelihood

        mean_prior = self._mean(x)
        cov_xx = lazify(prior.kernel.gram(x))
        obs_var = unwrap(likelihood.obs_stddev).to(x.device) ** 2
        obs_cov = diag_like(cov_xx, obs_var)
        cov_prior = cov_xx + obs_cov

        actions = self.policy.to_actions(cov_prior, key=key)
        residual = y - mean_prior
        # B200: register the residual so the single fused pass also yields c' = (K'S)^T residual
        projection = None
        if isinstance(actions, BlockDiagonalSparse) and isinstance(cov_xx, LazyKernel) and cov_xx.is_gram:
            projection = cov_xx.project(actions, residual)
        elif isinstance(actions, Dense) and isinstance(cov_xx, LazyKernel) and cov_xx.is_gram \
                and isinstance(self.solver, Cholesky) and _fused_dense_statistics_ok(self.policy, actions, cov_xx):
            projection = cov_xx.project_dense(actions, residual)
        obs_cov_proj = congruence_transform(actions, obs_cov)
        cov_prior_proj = congruence_transform(actions, cov_prior)
        cov_prior_proj_state = self.solver.init(cov_prior_proj)

        residual_proj = actions.T @ residual
        repr_weights_proj = self.solver.solve(cov_prior_proj_state, residual_proj)
        return ComputationAwareGPState(train_data=train_data, actions=actions, obs_cov_proj=obs_cov_proj,
                                       cov_prior_proj_state=cov_prior_proj_state, residual_proj=residual_proj,
                                       repr_weights_proj=repr_weights_proj, _projection=projection, _residual=residual,
                                       _cov_xx=cov_xx)

    def predict(self, state: ComputationAwareGPState, test_inputs: Optional[torch.Tensor] = None) -> GaussianDistribution:
        """Predictive distribution at ``test_inputs`` (training inputs if None) (models/cagp.py:126-169)."""
        x = state.train_data.X
        x = torch.atleast_2d(x)
        z = test_inputs if test_inputs is not None else x
        z = torch.atleast_2d(z)
        prior = self.posterior.prior
        mean_z = self._mean(z)
        if test_inputs is None:
            # same x, same kernel: reuse init's Gram operator so that its memoised K S product
            # (fused projection / dense product) is not recomputed
            cov_zz = state._cov_xx if state._cov_xx is not None else lazify(prior.kernel.gram(x))
            cov_zx = cov_zz
        else:
            cov_zz = lazify(prior.kernel.gram(z))
            cov_zx = lazify(prior.kernel.cross_covariance(z, x))
        cov_zx_proj = cov_zx @ state.actions

        mean_pred = torch.atleast_1d(mean_z + cov_zx_proj @ state.repr_weights_proj)
        cov_pred = PSD(PredictiveCovariance(cov_zz, cov_zx_proj, self.solver, state.cov_prior_proj_state))
        return GaussianDistribution(mean_pred, cov_pred, solver=self.solver)

    def prior_kl(self, state: ComputationAwareGPState) -> torch.Tensor:
        """KL[q(f) || p(f)] (models/cagp.py:171-201)."""
        obs_cov_proj_solver_state = self.solver.init(state.obs_cov_proj)
        kl = _kl_divergence_from_solvers(self.solver, state.residual_proj, obs_cov_proj_solver_state,
                                         torch.zeros_like(state.residual_proj), state.cov_prior_proj_state)
        w = state.repr_weights_proj
        return kl - 0.5 * congruence_transform(w, state.obs_cov_proj).squeeze()

    def variational_expectation(self, state: ComputationAwareGPState) -> torch.Tensor:
        """Pointwise expected log-likelihood under q (models/cagp.py:203-231)."""
        y = torch.atleast_1d(state.train_data.y).reshape(-1, 1)
        qpred = self.predict(state)
        mean, variance = qpred.mean, qpred.variance
        return self.posterior.likelihood.expected_log_likelihood(y, mean[:, None], variance[:, None])

    def elbo(self, state: ComputationAwareGPState) -> torch.Tensor:
        """Evidence lower bound: sum of variational expectations minus the prior KL (models/cagp.py:233-249)."""
        if state._projection is not None and isinstance(self.solver, Cholesky) \
                and isinstance(state.cov_prior_proj_state, Triangular):
            if isinstance(state.actions, Dense):
                return self._elbo_fused_dense(state)
            return self._elbo_fused(state)
        return torch.sum(self.variational_expectation(state)) - self.prior_kl(state)

    # ------------------------------------------------------------------------------------------
    def _elbo_fused_dense(self, state: ComputationAwareGPState) -> torch.Tensor:
        """Dense action matrix: same k-sized statistics, dense projected noise covariance."""
        return _elbo_fused_dense_impl(self, state)

    def _elbo_fused(self, state: ComputationAwareGPState) -> torch.Tensor:
        """ELBO from the k-sized statistics of the fused pass:
        sum_i var_i = n var - var^2 tr(P^-1 B'),  |y - mean|^2 = |r~|^2 - 2 var w.c' + var^2 w.B'w,
        minus the KL of ``prior_kl``; evaluated and differentiated in closed form (``_closed_form``)."""
        from .._closed_form import elbo_from_stats
        from .._fused import segment_sum

        proj = state._projection
        x = torch.atleast_2d(state.train_data.X)
        _, var = _kernel_params(self.posterior.prior.kernel, x)
        sigma = unwrap(self.posterior.likelihood.obs_stddev).to(device=x.device, dtype=var.dtype)
        nz = state.actions.nz_values
        q = segment_sum(nz * nz, state.actions.n_blocks)
        yt = state._residual
        L = state.cov_prior_proj_state  # lower factor of exactly P = v A' + diag(sigma^2 q) + jitter I (init)
        chol = L.A if (isinstance(L, Triangular) and L.lower and L.A.dtype == torch.float64) else None
        # The k-sized algebra always runs in float64: it costs O(k^3) against the O(n^2) pair kernels, and in
        # float32 it is what loses the digits (cancellation in w.B'w - tr(P^-1 B'), a Cholesky of a matrix with
        # condition number ~1/jitter).  Mixed-dtype callers (float32 data, float64 library defaults) land here too.
        f64 = lambda t: t.to(torch.float64)
        out = elbo_from_stats(f64(proj.A), f64(proj.B), f64(proj.c), f64(state.residual_proj), f64(q), f64(yt @ yt),
                              x.shape[0], f64(var), f64(sigma), self.solver.jitter, chol=chol)
        return out.to(proj.A.dtype)


def _fused_dense_statistics_ok(policy, actions, cov_xx) -> bool:
    """Whether the ELBO of a dense action matrix may be evaluated from the k-sized statistics A', B', c'.

    The closed form contains w^T B' w and tr(P^-1 B'), which cancel catastrophically when S^T K^ S is badly
    conditioned (raw pseudo-input actions S = K(X, Z): |w| ~ 1 / jitter).  It is therefore used only for actions
    that are (scaled-)orthogonal by construction or declared well-conditioned by their policy, and for row-sharded
    runs (where the statistics are what gets all-reduced).  Everything else takes the literal per-point route,
    which shares the single memoised K S product between ``init`` and ``predict`` and costs the same kernel passes."""
    from ..operators.annotations import ScaledOrthogonal
    from ..ops import StiefelAnnotation, UnitaryAnnotation

    orthogonal = any(actions.isa(a) for a in (StiefelAnnotation, UnitaryAnnotation, ScaledOrthogonal))
    if cov_xx.process_group is not None:
        if not (orthogonal or getattr(policy, "fused_statistics", False)):
            import warnings

            warnings.warn("row-sharded run with dense actions that are not orthogonal by construction: the all-reduced "
                          "k-sized statistics lose accuracy when S^T K S is ill-conditioned; wrap the policy in "
                          "OrthogonalizationPolicy (as the reference recommends for pseudo-input actions)")
        return True
    return orthogonal or getattr(policy, "fused_statistics", False)


def _elbo_fused_dense_impl(model, state):
    from .._closed_form import elbo_from_stats_dense

    proj = state._projection
    x = torch.atleast_2d(state.train_data.X)
    _, var = _kernel_params(model.posterior.prior.kernel, x)
    sigma = unwrap(model.posterior.likelihood.obs_stddev).to(device=x.device, dtype=var.dtype)
    S = state.actions.A
    yt = state._residual
    return elbo_from_stats_dense(proj.A, proj.B, proj.c, state.residual_proj, S.T @ S, yt @ yt, x.shape[0], var, sigma,
                                 0.0 if model.solver.jitter is None else model.solver.jitter)


def _kl_divergence_from_solvers(solver, mean_q, cov_q_state, mean_p, cov_p_state) -> torch.Tensor:
    """KL between two Gaussians given solver states (models/cagp.py:252-271)."""
    n = mean_q.shape[0]
    diff = mean_p - mean_q
    inner = solver.trace_solve(cov_p_state, cov_q_state)
    mahalanobis = solver.inv_quad(cov_p_state, diff)
    logdet_ratio = solver.logdet(cov_p_state) - solver.logdet(cov_q_state)
    return 0.5 * (inner + logdet_ratio + mahalanobis - n)
