"""Lower Cholesky factor of a PSD operator (reference: linalg/lower_cholesky.py:14-58)."""

import torch

from ..ops import PSD, Diagonal, Identity, LinearOperator, ScalarMul, Triangular
from .utils import _add_jitter


def lower_cholesky(A: LinearOperator, jitter=None) -> LinearOperator:
    """Lower Cholesky factor of ``A`` (+ ``jitter`` * I).  Structured operators keep their
    structure; a general (possibly lazy) operator is densified - for the projected kernel Gram
    operator that ``to_dense`` is the fused sm_100a pass - and factorised on the device."""
    if jitter is None:
        return _lower_cholesky(PSD(A))
    return _lower_cholesky(PSD(_add_jitter(A, jitter)))


def _lower_cholesky(A: LinearOperator) -> LinearOperator:
    if isinstance(A, Diagonal):
        return Diagonal(torch.sqrt(A.diag))
    if isinstance(A, ScalarMul):
        c = A.c if torch.is_tensor(A.c) else torch.as_tensor(A.c, dtype=A.dtype, device=A.device)
        return ScalarMul(torch.sqrt(c), A.shape, A.dtype, A.device)
    if isinstance(A, Identity):
        return A
    # no error check => no host synchronisation; like jnp.linalg.cholesky (lower_cholesky.py:38 of the
    # reference) a non-PD input yields NaNs rather than an exception
    return Triangular(torch.linalg.cholesky_ex(A.to_dense(), check_errors=False).L, lower=True)
