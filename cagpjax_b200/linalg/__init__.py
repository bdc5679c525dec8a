"""Linear algebra functions (reference: linalg/__init__.py)."""

from .congruence import ProjectedDenseGram, ProjectedKernelGram, congruence_transform
from .eigh import Eigh, EighResult, Lanczos, eigh
from .lower_cholesky import lower_cholesky
from .orthogonalize import OrthogonalizationMethod, orthogonalize

__all__ = ["congruence_transform", "lower_cholesky", "ProjectedKernelGram", "ProjectedDenseGram", "eigh", "Eigh",
           "EighResult", "Lanczos", "orthogonalize", "OrthogonalizationMethod"]
