"""Moore-Penrose pseudo-inverse linear solver (reference: solvers/pseudoinverse.py:17-143).

Works on the k x k projected operators of the CaGP model; the eigendecomposition comes from
``cagpjax_b200.linalg.eigh`` (dense on the device, or matrix-free Lanczos), so its reverse mode supports
(almost-)degenerate operators through ``grad_rtol``."""

from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from ..linalg.eigh import Algorithm, Eigh, EighResult, eigh
from ..ops import Dense, Diagonal, LinearOperator, lazify
from .base import AbstractLinearSolver


class PseudoInverseState(NamedTuple):
    A: LinearOperator
    eigh_result: EighResult
    eigvals_mask: torch.Tensor
    eigvals_safe: torch.Tensor
    eigvals_inv: torch.Tensor


class PseudoInverse(AbstractLinearSolver):
    """Least-squares solution ``x = A^+ b``: exact for non-singular ``A``, well defined for singular ``A``.

    Attributes:
        rtol: eigenvalues below ``rtol * max |eigenvalue|`` are treated as zero (default: eps * n, the
            ``jax.numpy.linalg.lstsq`` heuristic the reference uses).
        grad_rtol: cutoff for treating eigenvalues as equal in the reverse mode (None: all distinct).
        alg: eigendecomposition algorithm passed to ``eigh``.
    """

    def __init__(self, rtol: Optional[float] = None, grad_rtol: Optional[float] = None, alg: Optional[Algorithm] = None):
        if rtol is not None and rtol < 0:
            raise ValueError("rtol must be non-negative")
        if grad_rtol is not None and grad_rtol < 0:
            raise ValueError("grad_rtol must be non-negative")
        self.rtol = rtol
        self.grad_rtol = grad_rtol
        self.alg = alg if alg is not None else Eigh()

    def init(self, A: LinearOperator) -> PseudoInverseState:
        A = lazify(A)
        n = A.shape[0]
        rtol_val = self.rtol if self.rtol is not None else float(torch.finfo(A.dtype).eps) * n
        result = eigh(A, alg=self.alg, grad_rtol=self.grad_rtol)
        vals = result.eigenvalues
        cutoff = rtol_val * torch.max(torch.abs(vals))
        mask = vals >= cutoff
        safe = torch.where(mask, vals, torch.ones_like(vals))
        inv = torch.where(mask, torch.reciprocal(safe), torch.zeros_like(vals))
        return PseudoInverseState(A, result, mask, safe, inv)

    def unwhiten(self, state: PseudoInverseState, z: torch.Tensor) -> torch.Tensor:
        sqrt = torch.where(state.eigvals_mask, torch.sqrt(state.eigvals_safe), torch.zeros_like(state.eigvals_safe))
        x = sqrt[:, None] * z if z.dim() == 2 else sqrt * z
        return state.eigh_result.eigenvectors @ x

    def solve(self, state: PseudoInverseState, b: torch.Tensor) -> torch.Tensor:
        vec = b.dim() == 1
        bm = b[:, None] if vec else b
        V = state.eigh_result.eigenvectors
        x = V @ ((V.T @ bm) * state.eigvals_inv[:, None])
        return x[:, 0] if vec else x

    def logdet(self, state: PseudoInverseState) -> torch.Tensor:
        return torch.sum(torch.log(state.eigvals_safe))

    def inv_quad(self, state: PseudoInverseState, b: torch.Tensor) -> torch.Tensor:
        z = state.eigh_result.eigenvectors.T @ b
        return torch.dot(torch.square(z).reshape(-1), state.eigvals_inv).squeeze()

    def inv_congruence_transform(self, state: PseudoInverseState, B):
        z = state.eigh_result.eigenvectors.T @ B
        if isinstance(z, LinearOperator):
            return z.T @ Diagonal(state.eigvals_inv) @ z
        return z.T @ (state.eigvals_inv[:, None] * z)

    def trace_solve(self, state: PseudoInverseState, state_other: PseudoInverseState) -> torch.Tensor:
        V = state.eigh_result.eigenvectors.to_dense()
        if isinstance(state_other.eigh_result.eigenvectors, Dense):
            # tr(V diag(inv) V^T A_other)
            return torch.einsum("ij,j,kj,ik->", V, state.eigvals_inv, V, state_other.A.to_dense())
        W = state_other.eigh_result.eigenvectors.T @ V
        return torch.einsum("ij,j,ij,i->", W, state.eigvals_inv, W, state_other.eigvals_safe)
