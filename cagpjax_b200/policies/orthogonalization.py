"""Orthogonalising policy wrapper (reference: policies/orthogonalization.py:11-45).

Orthogonalises (if necessary) the action operator of a base policy.  Operators that are already
(scaled-)orthogonal pass through untouched (linalg/orthogonalize.py:82-93, pinned by
tests/test_policies.py:455-488): a ``BlockDiagonalSparse`` carries the ``ScaledOrthogonal`` annotation and Lanczos
vectors the ``Stiefel`` one, so on the block-sparse hot path this wrapper is the identity."""

from __future__ import annotations

from ..linalg.orthogonalize import OrthogonalizationMethod, orthogonalize
from ..ops import LinearOperator, lazify
from .base import AbstractBatchLinearSolverPolicy


class OrthogonalizationPolicy(AbstractBatchLinearSolverPolicy):
    """Args:
        base_policy: policy producing the action operator to orthogonalise.
        method: ``OrthogonalizationMethod`` (default Householder QR).
        n_reortho: extra re-orthogonalisation sweeps (one is generally sufficient for Gram-Schmidt variants)."""

    def __init__(self, base_policy: AbstractBatchLinearSolverPolicy,
                 method: OrthogonalizationMethod = OrthogonalizationMethod.QR, n_reortho: int = 0):
        self.base_policy = base_policy
        self.method = method
        self.n_reortho = n_reortho
        self.n_actions = base_policy.n_actions
        if n_reortho < 0:
            raise ValueError("n_reortho must be non-negative")
        self._check_init()

    def to_actions(self, A, *, key=None) -> LinearOperator:
        op = self.base_policy.to_actions(A, key=key)
        return lazify(orthogonalize(op, method=self.method, n_reortho=self.n_reortho))
