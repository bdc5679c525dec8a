"""Moore-Penrose pseudo-inverse linear solver (reference: solvers/pseudoinverse.py:17-143).

Every operation of the solver is a spectral filter: with A = V diag(lam) V^T and a weight vector g(lam),
``V diag(g) V^T b``.  ``solve`` uses g = 1/lam on the retained eigenvalues, ``unwhiten`` g = sqrt(lam), the
quadratic forms contract V^T b with g directly.  The eigendecomposition comes from ``cagpjax_b200.linalg.eigh``
(dense on the device, or matrix-free Lanczos), whose reverse mode tolerates (almost-)degenerate operators through
``grad_rtol``.  In the CaGP model the operators are the k x k projected covariances, so all of this is k-sized
device work next to the n^2 kernel passes."""

from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from ..linalg.eigh import Algorithm, Eigh, EighResult, eigh
from ..ops import Dense, Diagonal, LinearOperator, lazify
from .base import AbstractLinearSolver


class PseudoInverseState(NamedTuple):
    """Spectral data of ``A``: retained-eigenvalue mask, eigenvalues with the discarded ones replaced by 1
    (``eigvals_safe``) and the filtered reciprocals (``eigvals_inv``, 0 where discarded)."""

    A: LinearOperator
    eigh_result: EighResult
    eigvals_mask: torch.Tensor
    eigvals_safe: torch.Tensor
    eigvals_inv: torch.Tensor


def _filter_spectrum(vals: torch.Tensor, rtol: float):
    """Keep eigenvalues >= rtol * max |eigenvalue| (no host synchronisation: everything stays a device tensor)."""
    keep = vals >= rtol * vals.abs().max()
    one, zero = torch.ones_like(vals), torch.zeros_like(vals)
    safe = torch.where(keep, vals, one)
    return keep, safe, torch.where(keep, safe.reciprocal(), zero)


class PseudoInverse(AbstractLinearSolver):
    """Least-squares solution ``x = A^+ b``: exact for non-singular ``A``, well defined for singular ``A``, more
    stable for almost-singular ``A``.  (If the rank of ``A`` depends on hyper-parameters being optimised the
    objective is discontinuous there - the pseudo-inverse is.)

    Attributes:
        rtol: eigenvalues below ``rtol * max |eigenvalue|`` count as zero; default ``eps(dtype) * n``, the
            ``lstsq`` heuristic the reference takes from JAX.
        grad_rtol: cutoff for treating eigenvalues as equal in the reverse mode (None: all distinct - not finite
            for repeated eigenvalues, e.g. sigma^2 S^T S with orthonormal actions).
        alg: eigendecomposition algorithm passed to ``eigh`` (default dense ``Eigh``).
    """

    def __init__(self, rtol: Optional[float] = None, grad_rtol: Optional[float] = None, alg: Optional[Algorithm] = None):
        for name, value in (("rtol", rtol), ("grad_rtol", grad_rtol)):
            if value is not None and value < 0:
                raise ValueError(f"{name} must be non-negative")
        self.rtol, self.grad_rtol = rtol, grad_rtol
        self.alg = Eigh() if alg is None else alg

    # -- state -----------------------------------------------------------------------------------
    def init(self, A: LinearOperator) -> PseudoInverseState:
        A = lazify(A)
        rtol = float(torch.finfo(A.dtype).eps) * A.shape[0] if self.rtol is None else self.rtol
        decomposition = eigh(A, alg=self.alg, grad_rtol=self.grad_rtol)
        return PseudoInverseState(A, decomposition, *_filter_spectrum(decomposition.eigenvalues, rtol))

    @staticmethod
    def _filtered(state: PseudoInverseState, weights: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """V diag(weights) V^T b for a vector or a matrix b."""
        V = state.eigh_result.eigenvectors
        cols = b[:, None] if b.dim() == 1 else b
        out = V @ (weights[:, None] * (V.T @ cols))
        return out[:, 0] if b.dim() == 1 else out

    # -- solver interface ------------------------------------------------------------------------
    def solve(self, state: PseudoInverseState, b: torch.Tensor) -> torch.Tensor:
        return self._filtered(state, state.eigvals_inv, b)

    def unwhiten(self, state: PseudoInverseState, z: torch.Tensor) -> torch.Tensor:
        """V diag(sqrt(lam)) z: maps IID standard normal z to a sample with covariance A."""
        root = torch.where(state.eigvals_mask, state.eigvals_safe.sqrt(), torch.zeros_like(state.eigvals_safe))
        scaled = root[:, None] * z if z.dim() == 2 else root * z
        return state.eigh_result.eigenvectors @ scaled

    def logdet(self, state: PseudoInverseState) -> torch.Tensor:
        """Pseudo-determinant: sum of the logs of the retained eigenvalues."""
        return state.eigvals_safe.log().sum()

    def inv_quad(self, state: PseudoInverseState, b: torch.Tensor) -> torch.Tensor:
        coords = (state.eigh_result.eigenvectors.T @ b).reshape(-1)
        return torch.dot(coords * coords, state.eigvals_inv)

    def inv_congruence_transform(self, state: PseudoInverseState, B):
        """``B^T A^+ B``; lazy if ``B`` is an operator, dense otherwise."""
        coords = state.eigh_result.eigenvectors.T @ B
        if isinstance(coords, LinearOperator):
            return coords.T @ Diagonal(state.eigvals_inv) @ coords
        return coords.T @ (state.eigvals_inv[:, None] * coords)

    def trace_solve(self, state: PseudoInverseState, state_other: PseudoInverseState) -> torch.Tensor:
        """``trace(A^+ B)`` with ``B`` the operator behind ``state_other``."""
        V = state.eigh_result.eigenvectors.to_dense()
        other_vecs = state_other.eigh_result.eigenvectors
        if isinstance(other_vecs, Dense):
            # trace(V diag(inv) V^T B) = sum_j inv_j v_j^T B v_j
            BV = state_other.A.to_dense() @ V
            return torch.sum(state.eigvals_inv * torch.sum(V * BV, dim=0))
        # structured B (diagonal-like: its eigenvectors are the identity): sum_ij W_ij^2 inv_j mu_i
        W = other_vecs.T @ V
        return torch.sum((W * W) * state.eigvals_inv[None, :] * state_other.eigvals_safe[:, None])
