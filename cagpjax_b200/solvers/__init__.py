"""Linear solvers (reference: solvers/__init__.py)."""

from .base import AbstractLinearSolver
from .cholesky import Cholesky
from .pseudoinverse import PseudoInverse

__all__ = ["AbstractLinearSolver", "Cholesky", "PseudoInverse"]
