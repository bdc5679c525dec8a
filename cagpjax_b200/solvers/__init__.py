"""Solvers for the k x k projected systems of the CaGP model.

Counterpart of ``cagpjax.solvers``.  ``Cholesky`` is the solver of the measured hot path (its ``init`` densifies
the lazy projected Gram operator, which is the fused CUDA pass); ``PseudoInverse`` works from a device
eigendecomposition and is the choice for rank-deficient action sets.
"""

from .base import AbstractLinearSolver
from .cholesky import Cholesky
from .pseudoinverse import PseudoInverse, PseudoInverseState

__all__ = ["AbstractLinearSolver", "Cholesky", "PseudoInverse", "PseudoInverseState"]
