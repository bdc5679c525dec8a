"""Cholesky-based linear solver (reference: solvers/cholesky.py:17-70)."""

import torch

from ..linalg import lower_cholesky
from ..ops import Dense, Diagonal, LinearOperator, Product, ScalarMul, Triangular, densify
from .base import AbstractLinearSolver


def _linv_matmat(state: LinearOperator, B: torch.Tensor) -> torch.Tensor:
    """``inv(state) @ B`` for a triangular / diagonal / scalar Cholesky factor."""
    if isinstance(state, Triangular):
        return torch.linalg.solve_triangular(state.A, B, upper=not state.lower)
    if isinstance(state, Diagonal):
        return B / state.diag[:, None]
    if isinstance(state, ScalarMul):
        return B / state.c
    return B  # Identity


class Cholesky(AbstractLinearSolver):
    """Solve with the Cholesky factor of ``A + jitter * I``; the k x k factorisation and the
    triangular solves stay on the device (cuSOLVER / cuBLAS through torch)."""

    def __init__(self, jitter=None):
        if jitter is not None and jitter < 0:
            raise ValueError("jitter must be non-negative")
        self.jitter = jitter

    def init(self, A):
        return lower_cholesky(A, jitter=self.jitter)

    def unwhiten(self, state, z):
        return state @ z

    def solve(self, state, b):
        bm = b[:, None] if b.dim() == 1 else b
        z = _linv_matmat(state, bm)
        x = _linv_matmat(state.T, z)
        return x[:, 0] if b.dim() == 1 else x

    def logdet(self, state):
        from ..ops import diag

        return 2 * torch.sum(torch.log(diag(state)))

    def inv_congruence_transform(self, state, B):
        if isinstance(B, LinearOperator):
            B = B.to_dense()
        right = _linv_matmat(state, B)
        return right.T @ right

    def trace_solve(self, state, state_other):
        L = _linv_matmat(state, densify(state_other))
        return torch.sum(torch.square(L))
