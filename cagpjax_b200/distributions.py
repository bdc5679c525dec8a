"""Gaussian distribution with an implicit covariance (reference: distributions.py:15-106)."""

from __future__ import annotations

import math

import torch

from .ops import LinearOperator, diag
from .solvers import AbstractLinearSolver, Cholesky


class GaussianDistribution:
    """N(loc, scale) where ``scale`` is a (possibly lazy) covariance operator."""

    def __init__(self, loc: torch.Tensor, scale: LinearOperator, solver: AbstractLinearSolver = None, **kwargs):
        self.loc = loc
        self.scale = scale
        self.solver = solver if solver is not None else Cholesky(1e-6)
        self.batch_shape = ()
        self.event_shape = tuple(loc.shape)
        self.event_dim = 1
        self.support = "real_vector"  # numpyro.distributions.constraints.real_vector in the reference

    def shape(self, sample_shape=()):
        return tuple(sample_shape) + self.batch_shape + self.event_shape

    @property
    def mean(self) -> torch.Tensor:
        return self.loc

    @property
    def variance(self) -> torch.Tensor:
        """Marginal variances, ``cola.diag(scale)`` (distributions.py:53-56)."""
        return diag(self.scale)

    @property
    def stddev(self) -> torch.Tensor:
        return torch.sqrt(self.variance)

    def covariance(self) -> LinearOperator:
        return self.scale

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        mu, sigma = self.loc, self.scale
        n = mu.shape[-1]
        state = self.solver.init(sigma)
        return (n * math.log(2 * math.pi) + self.solver.logdet(state) + self.solver.inv_quad(state, value - mu)) / -2

    def sample(self, key, sample_shape=()) -> torch.Tensor:
        mu, sigma = self.loc, self.scale
        n = mu.shape[-1]
        state = self.solver.init(sigma)
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        z = torch.randn((n, math.prod(sample_shape)), generator=gen, dtype=mu.dtype).to(mu.device)
        x = self.solver.unwhiten(state, z)
        return x.T.reshape(tuple(sample_shape) + (n,)) + mu
