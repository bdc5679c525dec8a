"""Linear solvers (reference: solvers/__init__.py)."""

from .base import AbstractLinearSolver
from .cholesky import Cholesky

__all__ = ["AbstractLinearSolver", "Cholesky"]
