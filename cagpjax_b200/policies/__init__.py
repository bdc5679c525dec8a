"""Linear solver policies (reference: policies/__init__.py)."""

from .base import AbstractBatchLinearSolverPolicy, AbstractLinearSolverPolicy
from .block_sparse import BlockSparsePolicy
from .dense import DensePolicy
from .orthogonalization import OrthogonalizationPolicy

__all__ = ["AbstractLinearSolverPolicy", "AbstractBatchLinearSolverPolicy", "BlockSparsePolicy", "DensePolicy", "OrthogonalizationPolicy"]
