"""Lanczos-based policy (reference: policies/lanczos.py:13-46).

Actions are approximate eigenvectors of the operator ``A`` passed to ``to_actions`` (the prior covariance
K + sigma^2 I in ``ComputationAwareGP.init``): ``n_actions`` matrix-free Lanczos iterations, each one matvec
``A @ v`` - for a ``LazyKernel`` one pass of the sm_100a pair kernels - followed by a small dense ``eigh``."""

from __future__ import annotations

from typing import Optional

from ..linalg.eigh import Lanczos, eigh
from .base import AbstractBatchLinearSolverPolicy


class LanczosPolicy(AbstractBatchLinearSolverPolicy):
    """Attributes:
        n_actions: number of Lanczos vectors / actions (None: as many as rows of the operator).
        grad_rtol: cutoff for treating eigenvalues as equal in the reverse mode (default 0.0; None or negative:
            all distinct), see ``cagpjax_b200.linalg.eigh``."""

    def __init__(self, n_actions: Optional[int], grad_rtol: Optional[float] = 0.0):
        self.n_actions = n_actions
        self.grad_rtol = grad_rtol
        if n_actions is not None:
            self._check_init()

    def to_actions(self, A, *, key=None):
        return eigh(A, alg=Lanczos(self.n_actions, key=key), grad_rtol=self.grad_rtol).eigenvectors
