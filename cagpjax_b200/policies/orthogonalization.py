"""Orthogonalising policy wrapper (reference: policies/orthogonalization.py:11-45).

``to_actions`` asks the wrapped policy for its action operator and hands it to
``cagpjax_b200.linalg.orthogonalize``.  That function returns operators that are already (scaled-)orthogonal
unchanged (linalg/orthogonalize.py:82-93 of the reference, pinned by its tests/test_policies.py:455-488): a
``BlockDiagonalSparse`` carries the ``ScaledOrthogonal`` annotation and Lanczos vectors the ``Stiefel`` one, so on
the block-sparse hot path this wrapper costs nothing.  Dense actions (pseudo-input cross-covariances) are
orthogonalised on the device; the result keeps its orthogonality annotation, which is what lets
``ComputationAwareGP`` evaluate their ELBO from the k-sized statistics of one fused pass."""

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
        if n_reortho < 0:
            raise ValueError("n_reortho must be non-negative")
        self.base_policy, self.method, self.n_reortho = base_policy, method, n_reortho
        self.n_actions = base_policy.n_actions
        self._check_init()

    def to_actions(self, A, *, key=None) -> LinearOperator:
        raw_actions = self.base_policy.to_actions(A, key=key)
        return lazify(orthogonalize(raw_actions, method=self.method, n_reortho=self.n_reortho))
