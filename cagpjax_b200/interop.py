"""Operator conversion helpers (reference: interop.py:47-81).

The reference converts between lineax and cola operators; this package has a single operator
family (``cagpjax_b200.ops``), so ``lazify`` only wraps raw tensors."""

from .ops import LinearOperator, lazify as _lazify


def lazify(A) -> LinearOperator:
    return _lazify(A)
