"""``diag_like`` (reference: operators/diag_like.py:8-26)."""

import torch

from ..ops import Diagonal, LinearOperator, ScalarMul


def diag_like(operator: LinearOperator, values):
    """Diagonal operator with the shape, dtype and device of ``operator``: a scalar gives a
    ``ScalarMul``, a vector a ``Diagonal``."""
    if not torch.is_tensor(values) or values.dim() == 0:
        if torch.is_tensor(values):
            values = values.to(device=operator.device, dtype=operator.dtype)
        return ScalarMul(values, operator.shape, dtype=operator.dtype, device=operator.device)
    return Diagonal(values.to(device=operator.device, dtype=operator.dtype))
