"""Linear algebra utilities (reference: linalg/utils.py:15-41)."""

from ..operators import diag_like
from ..ops import Diagonal, Identity, LinearOperator, ScalarMul, diag


def _add_jitter(A: LinearOperator, jitter):
    """A + jitter * I, preserving structure for ScalarMul / Identity / Diagonal operands."""
    import torch

    vector_jitter = torch.is_tensor(jitter) and jitter.dim() > 0
    if isinstance(A, ScalarMul) and not vector_jitter:
        return ScalarMul(A.c + jitter, A.shape, A.dtype, A.device)
    if isinstance(A, Identity) and not vector_jitter:
        return ScalarMul(1.0 + jitter, A.shape, A.dtype, A.device)
    if isinstance(A, (ScalarMul, Diagonal, Identity)):
        return Diagonal(diag(A) + jitter)
    return A + diag_like(A, jitter)
