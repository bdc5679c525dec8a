"""Linear solver policies (reference: policies/__init__.py)."""

from .base import AbstractBatchLinearSolverPolicy, AbstractLinearSolverPolicy
from .block_sparse import BlockSparsePolicy
from .dense import DensePolicy
from .lanczos import LanczosPolicy
from .orthogonalization import OrthogonalizationPolicy
from .pseudoinput import PseudoInputPolicy

__all__ = ["AbstractLinearSolverPolicy", "AbstractBatchLinearSolverPolicy", "BlockSparsePolicy", "DensePolicy",
           "LanczosPolicy", "OrthogonalizationPolicy", "PseudoInputPolicy"]
