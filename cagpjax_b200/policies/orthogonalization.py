"""Orthogonalising policy wrapper (reference: policies/orthogonalization.py:11-45).

The reference wraps a base policy and orthogonalises its action matrix, but passes operators that are already
(scaled-)orthogonal through untouched (policies/orthogonalization.py:38-45, pinned by tests/test_policies.py:
473-488): a ``BlockDiagonalSparse`` carries the ``ScaledOrthogonal`` annotation, so on the hot path this wrapper
is the identity.  Dense base actions are orthonormalised with a thin QR on the device."""

from __future__ import annotations

import torch

from ..operators.annotations import ScaledOrthogonal
from ..ops import Dense, LinearOperator
from .base import AbstractBatchLinearSolverPolicy


class OrthogonalizationPolicy(AbstractBatchLinearSolverPolicy):
    def __init__(self, base_policy: AbstractBatchLinearSolverPolicy):
        self.base_policy = base_policy
        self.n_actions = base_policy.n_actions
        self._check_init()

    def to_actions(self, A, *, key=None) -> LinearOperator:
        actions = self.base_policy.to_actions(A, key=key)
        if actions.isa(ScaledOrthogonal):
            return actions
        Q, _ = torch.linalg.qr(actions.to_dense(), mode="reduced")
        return Dense(Q)
