"""Action policies: which n x k subspace the data is projected onto.

Counterpart of ``cagpjax.policies``.  ``BlockSparsePolicy`` is the policy of the measured hot path (one learnable
non-zero per data point, fused symmetric kernels); ``DensePolicy`` holds a learnable dense matrix (BASELINE config
5); ``PseudoInputPolicy`` materialises K(X, Z) with the native cross-covariance kernels; ``LanczosPolicy`` runs
matrix-free Lanczos on the lazy kernel operator; ``OrthogonalizationPolicy`` wraps any of them.
"""

from .base import AbstractBatchLinearSolverPolicy, AbstractLinearSolverPolicy
from .block_sparse import BlockSparsePolicy
from .dense import DensePolicy
from .lanczos import LanczosPolicy
from .orthogonalization import OrthogonalizationPolicy
from .pseudoinput import PseudoInputPolicy

__all__ = [
    "AbstractBatchLinearSolverPolicy",
    "AbstractLinearSolverPolicy",
    "BlockSparsePolicy",
    "DensePolicy",
    "LanczosPolicy",
    "OrthogonalizationPolicy",
    "PseudoInputPolicy",
]
