"""Base class for linear solvers (reference: solvers/base.py:15-106)."""

from abc import ABC, abstractmethod


class AbstractLinearSolver(ABC):
    """Solvers for ``A x = b`` with a positive (semi-)definite operator ``A``."""

    @abstractmethod
    def init(self, A):
        """Construct a solver state for ``A``."""

    @abstractmethod
    def solve(self, state, b):
        """Solution of ``A x = b``."""

    @abstractmethod
    def unwhiten(self, state, z):
        """Given IID standard normal ``z`` return ``x`` with covariance ``A``."""

    @abstractmethod
    def logdet(self, state):
        """log |A|."""

    @abstractmethod
    def inv_congruence_transform(self, state, B):
        """``B^T A^-1 B``."""

    def inv_quad(self, state, b):
        """``b^T A^-1 b`` (solvers/base.py:84-93)."""
        return self.inv_congruence_transform(state, b[:, None]).squeeze()

    @abstractmethod
    def trace_solve(self, state, state_other):
        """``trace(A^-1 B)`` for the operators behind the two states."""
