"""Policy base classes (reference: policies/base.py:8-43)."""

from abc import ABC, abstractmethod


class AbstractLinearSolverPolicy(ABC):
    """Policies define actions used to solve ``A x = b``."""


class AbstractBatchLinearSolverPolicy(AbstractLinearSolverPolicy):
    """Policies that produce an (n, n_actions) action matrix."""

    n_actions: int

    def _check_init(self):
        if self.n_actions < 1:
            raise ValueError("n_actions must be at least 1")

    @abstractmethod
    def to_actions(self, A, *, key=None):
        """Action operator of shape (n, n_actions) for the (n, n) operator ``A``."""
