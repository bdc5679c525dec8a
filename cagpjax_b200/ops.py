"""Minimal linear-operator algebra on torch tensors.

Stands in for the part of CoLA (``cola.ops``) the reference's hot path touches: ``LinearOperator``
with ``_matmat`` / ``_rmatmat``, ``.T``, ``@``, ``to_dense()``, lazy ``Sum`` / ``Product`` /
``Transpose``, the structured ``Dense`` / ``Diagonal`` / ``ScalarMul`` / ``Identity`` /
``Triangular`` operators and annotations (``PSD``).  CoLA is a third-party dependency that is
not vendored in /root/reference (pyproject.toml:33); only its protocol is mirrored here
(SURVEY 8b(ii)): ``to_dense`` of a lazy operator is ``op @ eye`` (tests/test_operators.py:24-28).
"""

from __future__ import annotations

from typing import Iterable, Sequence

import torch


class Annotation:
    """Base class for operator annotations (cola.annotations.Annotation)."""


class PSDAnnotation(Annotation):
    pass


class SelfAdjointAnnotation(Annotation):
    pass


class UnitaryAnnotation(Annotation):
    """Square operator with orthonormal columns (cola.Unitary)."""


class StiefelAnnotation(Annotation):
    """Tall operator with orthonormal columns (cola.Stiefel)."""


class LinearOperator:
    """Abstract (m, n) linear operator; subclasses implement ``_matmat`` (and ``_rmatmat``)."""

    def __init__(self, dtype: torch.dtype, shape: Sequence[int], annotations: Iterable[type] = (), device=None):
        self.dtype = dtype
        self.shape = tuple(int(v) for v in shape)
        self.annotations = set(annotations)
        self.device = torch.device(device) if device is not None else torch.device("cpu")

    # -- protocol -------------------------------------------------------------------------
    def _matmat(self, X: torch.Tensor) -> torch.Tensor:  # op @ X, X (n, p)
        raise NotImplementedError

    def _rmatmat(self, X: torch.Tensor) -> torch.Tensor:  # X @ op, X (p, m)
        return (self.T._matmat(X.T)).T

    # -- algebra --------------------------------------------------------------------------
    @property
    def T(self) -> "LinearOperator":
        return Transpose(self)

    def isa(self, annotation: type) -> bool:
        return annotation in self.annotations

    def to_dense(self) -> torch.Tensor:
        return self._matmat(torch.eye(self.shape[1], dtype=self.dtype, device=self.device))

    def __matmul__(self, other):
        if isinstance(other, LinearOperator):
            return Product(self, other)
        other = torch.as_tensor(other)
        if other.dim() == 1:
            return self._matmat(other[:, None])[:, 0]
        return self._matmat(other)

    def __rmatmul__(self, other):
        other = torch.as_tensor(other)
        if other.dim() == 1:
            return self._rmatmat(other[None, :])[0]
        return self._rmatmat(other)

    def __add__(self, other):
        if not isinstance(other, LinearOperator):
            other = Dense(torch.as_tensor(other))
        return Sum(self, other)

    def __radd__(self, other):
        return Dense(torch.as_tensor(other)) + self

    def __neg__(self):
        return Product(ScalarMul(-1.0, (self.shape[0], self.shape[0]), self.dtype, self.device), self)

    def __sub__(self, other):
        if not isinstance(other, LinearOperator):
            other = Dense(torch.as_tensor(other))
        return Sum(self, -other)

    def __mul__(self, c):
        return Product(ScalarMul(c, (self.shape[0], self.shape[0]), self.dtype, self.device), self)

    __rmul__ = __mul__


class Dense(LinearOperator):
    def __init__(self, A: torch.Tensor):
        super().__init__(A.dtype, A.shape, device=A.device)
        self.A = A

    def _matmat(self, X):
        return self.A @ X

    def _rmatmat(self, X):
        return X @ self.A

    def to_dense(self):
        return self.A

    @property
    def T(self):
        return Dense(self.A.T)


class Diagonal(LinearOperator):
    def __init__(self, diag: torch.Tensor):
        n = diag.shape[0]
        super().__init__(diag.dtype, (n, n), device=diag.device)
        self.diag = diag

    def _matmat(self, X):
        return self.diag[:, None] * X

    def _rmatmat(self, X):
        return X * self.diag[None, :]

    def to_dense(self):
        return torch.diag(self.diag)

    @property
    def T(self):
        return self


class ScalarMul(LinearOperator):
    def __init__(self, c, shape, dtype=None, device=None):
        if torch.is_tensor(c):
            dtype = dtype or c.dtype
            device = device if device is not None else c.device
        super().__init__(dtype or torch.get_default_dtype(), shape, device=device)
        self.c = c

    def _matmat(self, X):
        return self.c * X

    def _rmatmat(self, X):
        return self.c * X

    def to_dense(self):
        return self.c * torch.eye(self.shape[0], dtype=self.dtype, device=self.device)

    @property
    def T(self):
        return self


class Identity(LinearOperator):
    def __init__(self, shape, dtype=None, device=None):
        super().__init__(dtype or torch.get_default_dtype(), shape, device=device)

    def _matmat(self, X):
        return X

    def _rmatmat(self, X):
        return X

    def to_dense(self):
        return torch.eye(self.shape[0], dtype=self.dtype, device=self.device)

    @property
    def T(self):
        return self


class Triangular(LinearOperator):
    def __init__(self, A: torch.Tensor, lower: bool = True):
        super().__init__(A.dtype, A.shape, device=A.device)
        self.A = A
        self.lower = lower

    def _matmat(self, X):
        return self.A @ X

    def _rmatmat(self, X):
        return X @ self.A

    def to_dense(self):
        return self.A

    @property
    def T(self):
        return Triangular(self.A.T, lower=not self.lower)


class Transpose(LinearOperator):
    def __init__(self, op: LinearOperator):
        super().__init__(op.dtype, op.shape[::-1], device=op.device)
        self.op = op

    def _matmat(self, X):
        return self.op._rmatmat(X.T).T

    def _rmatmat(self, X):
        return self.op._matmat(X.T).T

    @property
    def T(self):
        return self.op


class Sum(LinearOperator):
    def __init__(self, *ops: LinearOperator):
        shape = ops[0].shape
        for o in ops:
            if o.shape != shape:
                raise ValueError(f"shape mismatch in Sum: {o.shape} vs {shape}")
        super().__init__(ops[0].dtype, shape, device=ops[0].device)
        self.ops = tuple(ops)

    def _matmat(self, X):
        out = self.ops[0]._matmat(X)
        for o in self.ops[1:]:
            out = out + o._matmat(X)
        return out

    def _rmatmat(self, X):
        out = self.ops[0]._rmatmat(X)
        for o in self.ops[1:]:
            out = out + o._rmatmat(X)
        return out

    @property
    def T(self):
        return Sum(*[o.T for o in self.ops])


class Product(LinearOperator):
    def __init__(self, *ops: LinearOperator):
        for a, b in zip(ops[:-1], ops[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError(f"shape mismatch in Product: {a.shape} @ {b.shape}")
        super().__init__(ops[0].dtype, (ops[0].shape[0], ops[-1].shape[1]), device=ops[0].device)
        self.ops = tuple(ops)

    def _matmat(self, X):
        for o in reversed(self.ops):
            X = o._matmat(X)
        return X

    def _rmatmat(self, X):
        for o in self.ops:
            X = o._rmatmat(X)
        return X

    @property
    def T(self):
        return Product(*[o.T for o in reversed(self.ops)])


def PSD(op: LinearOperator) -> LinearOperator:
    """Annotate ``op`` as positive semi-definite (cola.PSD)."""
    op.annotations.add(PSDAnnotation)
    return op


def SelfAdjoint(op: LinearOperator) -> LinearOperator:
    op.annotations.add(SelfAdjointAnnotation)
    return op


def Unitary(op: LinearOperator) -> LinearOperator:
    op.annotations.add(UnitaryAnnotation)
    return op


def Stiefel(op: LinearOperator) -> LinearOperator:
    op.annotations.add(StiefelAnnotation)
    return op


def I_like(op: LinearOperator) -> "Identity":
    return Identity(op.shape, op.dtype, op.device)


def lazify(A) -> LinearOperator:
    if isinstance(A, LinearOperator):
        return A
    return Dense(torch.as_tensor(A))


def densify(A) -> torch.Tensor:
    return A.to_dense() if isinstance(A, LinearOperator) else torch.as_tensor(A)


def diag(op) -> torch.Tensor:
    """Diagonal of an operator (cola.linalg.diag / cola.diag).  Structured cases are exact and
    cheap; the generic fallback densifies (only used for small operators)."""
    if isinstance(op, Diagonal):
        return op.diag
    if isinstance(op, ScalarMul):
        return op.c * torch.ones(op.shape[0], dtype=op.dtype, device=op.device)
    if isinstance(op, Identity):
        return torch.ones(op.shape[0], dtype=op.dtype, device=op.device)
    if isinstance(op, (Dense, Triangular)):
        return torch.diagonal(op.A)
    if hasattr(op, "_diag"):
        return op._diag()
    if isinstance(op, Sum):
        out = diag(op.ops[0])
        for o in op.ops[1:]:
            out = out + diag(o)
        return out
    return torch.diagonal(op.to_dense())


def inv_triangular_matmat(L: Triangular, B: torch.Tensor) -> torch.Tensor:
    """cola.linalg.inv(Triangular) @ B."""
    return torch.linalg.solve_triangular(L.A, B, upper=not L.lower)
