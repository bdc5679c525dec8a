"""Orthogonalization methods (reference: linalg/orthogonalize.py:16-196).

Operands here are tall action matrices (n, k) living on the device; every variant is expressed in GEMV / GEMM
shaped torch calls so that it runs on the GPU and is differentiable by reverse mode (the reference test
``tests/test_linalg.py:487-500`` checks first-order gradients of all three methods).
"""

from __future__ import annotations

from enum import Enum

import torch

from ..operators.annotations import ScaledOrthogonal
from ..ops import Dense, Diagonal, I_like, Identity, LinearOperator, ScalarMul, StiefelAnnotation, Unitary, densify


class OrthogonalizationMethod(Enum):
    """Methods for orthogonalizing a matrix."""

    QR = "qr"    # Householder QR decomposition
    CGS = "cgs"  # classical Gram-Schmidt
    MGS = "mgs"  # modified Gram-Schmidt


def orthogonalize(A, /, method: OrthogonalizationMethod = OrthogonalizationMethod.QR, n_reortho: int = 0):
    """Orthogonalise the columns of ``A`` (tensor or operator): the result spans (a superset of) the column space
    of ``A`` with mutually orthogonal columns; Gram-Schmidt variants may return zero columns for rank-deficient
    input.  ``n_reortho`` extra re-orthogonalisation sweeps (one is generally enough for Gram-Schmidt).  Operators
    that are already (scaled-)orthogonal pass through unchanged (linalg/orthogonalize.py:82-93)."""
    if isinstance(A, LinearOperator):
        if isinstance(A, (Identity, Diagonal, ScalarMul)):
            return Unitary(I_like(A))
        if A.isa(StiefelAnnotation) or A.isa(ScaledOrthogonal):
            return A
        Q = Dense(orthogonalize(densify(A), method=method, n_reortho=n_reortho))
        Q.annotations.add(StiefelAnnotation if method is OrthogonalizationMethod.QR else ScaledOrthogonal)
        return Q
    if n_reortho < 0:
        raise ValueError("n_reortho must be non-negative")
    A = torch.as_tensor(A)
    if method is OrthogonalizationMethod.QR:
        return _qr_q(A, n_reortho)
    if method is OrthogonalizationMethod.CGS:
        return _classical_gram_schmidt(A, n_reortho)
    if method is OrthogonalizationMethod.MGS:
        return _modified_gram_schmidt(A, n_reortho)
    raise ValueError(f"unknown orthogonalization method {method!r}")


def _qr_q(A: torch.Tensor, n_reortho: int) -> torch.Tensor:
    Q = A
    for _ in range(n_reortho + 1):
        Q = torch.linalg.qr(Q, mode="reduced")[0]
    return Q


def _l2_normalize(x: torch.Tensor, *, rtol: float = 0.0) -> torch.Tensor:
    """x / |x|, or x itself (with zero differential) when |x| <= rtol * len(x) (linalg/orthogonalize.py:185-196)."""
    r = torch.linalg.vector_norm(x)
    small = r <= rtol * x.numel()
    return torch.where(small, x.detach(), x / torch.where(small, torch.ones_like(r), r))


def _classical_gram_schmidt(A: torch.Tensor, n_reortho: int) -> torch.Tensor:
    """Column k is orthogonalised against all finalised columns at once (two GEMVs per sweep).  As in the
    reference a small relative cutoff keeps ill-conditioned inputs stable."""
    rtol = float(torch.finfo(A.dtype).eps * 10)
    cols = list(A.unbind(dim=1))
    for k in range(len(cols)):
        q = cols[k]
        if k > 0:
            Qf = torch.stack(cols[:k], dim=1)
            for _ in range(n_reortho + 1):
                q = q - Qf @ (Qf.T @ q)
        cols[k] = _l2_normalize(q, rtol=rtol)
    return torch.stack(cols, dim=1)


def _modified_gram_schmidt(A: torch.Tensor, n_reortho: int) -> torch.Tensor:
    """Right-looking modified Gram-Schmidt: once column k is final, every later column is orthogonalised against
    it with one rank-1 update; ``n_reortho`` left-looking sweeps re-orthogonalise column k first."""
    n_cols = A.shape[1]
    done = []
    rest = A
    for k in range(n_cols):
        q = rest[:, 0]
        for _ in range(n_reortho):
            for qj in done:
                q = q - torch.dot(qj, q) * qj
        q = _l2_normalize(q)
        done.append(q)
        rest = rest[:, 1:]
        if rest.shape[1]:
            rest = rest - torch.outer(q, rest.T @ q)
    return torch.stack(done, dim=1)
