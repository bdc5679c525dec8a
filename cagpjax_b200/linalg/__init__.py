"""Linear algebra on operators: congruence transforms, Cholesky factors, eigendecompositions, orthogonalisation.

Counterpart of ``cagpjax.linalg``.  ``congruence_transform`` carries the dispatch that routes S^T (K + D) S with a
``LazyKernel`` to the fused CUDA projections (``ProjectedKernelGram`` / ``ProjectedDenseGram``).
"""

from .congruence import ProjectedDenseGram, ProjectedKernelGram, congruence_transform
from .eigh import Eigh, EighResult, Lanczos, eigh
from .lower_cholesky import lower_cholesky
from .orthogonalize import OrthogonalizationMethod, orthogonalize

__all__ = [
    "Eigh",
    "EighResult",
    "Lanczos",
    "OrthogonalizationMethod",
    "ProjectedDenseGram",
    "ProjectedKernelGram",
    "congruence_transform",
    "eigh",
    "lower_cholesky",
    "orthogonalize",
]
