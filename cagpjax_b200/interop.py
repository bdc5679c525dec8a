"""Operator conversion helpers (reference: interop.py:13-81).

The reference keeps two operator families apart - GPJax wants lineax operators, the CaGP code works on cola operators -
and converts at the seams: ``ColaLinearOperator`` presents a cola operator through the lineax interface (``mv``,
``as_matrix``, ``transpose``, ``in_structure`` / ``out_structure``, symmetric / PSD tags), ``lazify`` goes back, and
``to_lineax`` converts only where GPJax requires it.  This package has ONE operator family (``cagpjax_b200.ops``), so
the "lineax side" is the thin wrapper below with the same method names; a compute engine may return either form and
``lazify`` unwraps it exactly like the reference's (interop.py:47-66)."""

from __future__ import annotations

from typing import Any, Tuple

import torch

from .ops import Dense, Diagonal, Identity, LinearOperator, PSDAnnotation, lazify as _lazify


class ColaLinearOperator:
    """Lineax-shaped view of a ``LinearOperator`` (reference interop.py:13-34)."""

    def __init__(self, operator: LinearOperator):
        self.operator = operator

    def mv(self, vector: torch.Tensor) -> torch.Tensor:
        return self.operator @ vector

    def as_matrix(self) -> torch.Tensor:
        return self.operator.to_dense()

    def transpose(self) -> "ColaLinearOperator":
        return ColaLinearOperator(self.operator.T)

    @property
    def T(self) -> "ColaLinearOperator":
        return self.transpose()

    def in_structure(self) -> Tuple[Tuple[int, ...], torch.dtype]:
        """(shape, dtype) of the operand - the torch counterpart of ``jax.ShapeDtypeStruct``."""
        return (self.operator.shape[1],), self.operator.dtype

    def out_structure(self) -> Tuple[Tuple[int, ...], torch.dtype]:
        return (self.operator.shape[0],), self.operator.dtype

    def in_size(self) -> int:
        return self.operator.shape[1]

    def out_size(self) -> int:
        return self.operator.shape[0]


def is_symmetric(operator: Any) -> bool:
    """``lx.is_symmetric`` for the wrapper (interop.py:37-39): true iff the wrapped operator is tagged PSD."""
    op = operator.operator if isinstance(operator, ColaLinearOperator) else operator
    return isinstance(op, LinearOperator) and op.isa(PSDAnnotation)


def is_positive_semidefinite(operator: Any) -> bool:
    """``lx.is_positive_semidefinite`` for the wrapper (interop.py:42-44)."""
    return is_symmetric(operator)


def lazify(A: Any) -> LinearOperator:
    """Convert wrapper / operator / tensor inputs into a ``LinearOperator`` (interop.py:47-66)."""
    if isinstance(A, ColaLinearOperator):
        return A.operator
    return _lazify(A)


def to_lineax(A: Any):
    """Convert into the lineax-shaped form only where a GPJax-style consumer requires it (interop.py:69-81):
    wrappers pass through, structured operators stay structured inside the wrapper, raw tensors become ``Dense``."""
    if isinstance(A, ColaLinearOperator):
        return A
    if isinstance(A, (Diagonal, Identity)):
        return ColaLinearOperator(A)
    if isinstance(A, LinearOperator):
        return ColaLinearOperator(A)
    return ColaLinearOperator(Dense(torch.as_tensor(A)))
