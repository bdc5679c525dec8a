"""Linear algebra functions (reference: linalg/__init__.py)."""

from .congruence import ProjectedDenseGram, ProjectedKernelGram, congruence_transform
from .lower_cholesky import lower_cholesky

__all__ = ["congruence_transform", "lower_cholesky", "ProjectedKernelGram", "ProjectedDenseGram"]
