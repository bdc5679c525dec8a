"""Congruence transformations A^T B A (reference: linalg/congruence.py:15-54)."""

from __future__ import annotations

import torch

from .. import _fused
from ..operators import BlockDiagonalSparse, LazyKernel
from ..ops import Dense, Diagonal, LinearOperator, ScalarMul, Sum, diag


class ProjectedKernelGram(LinearOperator):
    """S^T (K + D) S for block-sparse S, a ``LazyKernel`` Gram operator K and diagonal-like D.

    Lazy (k, k) operator, like the reference's ``A.T @ (B @ A)`` fallback (congruence.py:15-17);
    its ``to_dense`` - which the Cholesky solver triggers (lower_cholesky.py:36-38) - is the fused
    sm_100a projection instead of a dense n x k pass through ``LazyKernel._matmat``."""

    def __init__(self, actions: BlockDiagonalSparse, lazy: LazyKernel, diag_terms):
        k = actions.shape[1]
        super().__init__(lazy.dtype, (k, k), device=lazy.device)
        self.actions = actions
        self.lazy = lazy
        self.diag_terms = tuple(diag_terms)

    def projection(self):
        return self.lazy.project(self.actions)

    def to_dense(self) -> torch.Tensor:
        from ..operators.lazy_kernel import _kernel_params

        _, var = _kernel_params(self.lazy.kernel, self.lazy.x1)
        out = var * self.projection().A
        for D in self.diag_terms:
            out = out + torch.diag(congruence_transform(self.actions, D).diag)
        return out

    def _matmat(self, X):
        return self.to_dense() @ X

    def _rmatmat(self, X):
        return X @ self.to_dense()

    @property
    def T(self):
        return self


class ProjectedDenseGram(LinearOperator):
    """S^T (K + D) S for a DENSE action matrix S; ``to_dense`` reuses the memoised dense kernel product
    K S (one 2 n^2 k pass on the GPU) instead of the generic chain of lazy products."""

    def __init__(self, actions: Dense, lazy: LazyKernel, diag_terms):
        k = actions.shape[1]
        super().__init__(lazy.dtype, (k, k), device=lazy.device)
        self.actions = actions
        self.lazy = lazy
        self.diag_terms = tuple(diag_terms)

    def to_dense(self) -> torch.Tensor:
        from ..operators.lazy_kernel import _kernel_params

        from ..operators.lazy_kernel import _tensor_key

        S = self.actions.A
        proj = self.lazy._projections.get(("dense",) + _tensor_key(S))
        if proj is not None:  # statistics of the (possibly row-sharded) fused pass registered by init
            out = _kernel_params(self.lazy.kernel, self.lazy.x1)[1] * proj.A
        else:
            out = S.T @ self.lazy.dense_actions(self.actions).to_dense()
        for D in self.diag_terms:
            out = out + S.T @ (D @ S)
        return out

    def _matmat(self, X):
        return self.to_dense() @ X

    def _rmatmat(self, X):
        return X @ self.to_dense()

    @property
    def T(self):
        return self


def _split_kernel_sum(B):
    """If B = LazyKernel(gram) [+ diagonal-like terms] return (lazy, [diag terms]) else None."""
    if isinstance(B, LazyKernel):
        return (B, []) if B.is_gram else None
    if isinstance(B, Sum):
        lazies = [o for o in B.ops if isinstance(o, LazyKernel)]
        others = [o for o in B.ops if not isinstance(o, LazyKernel)]
        if len(lazies) == 1 and lazies[0].is_gram and all(isinstance(o, (Diagonal, ScalarMul)) for o in others):
            return lazies[0], others
    return None


def congruence_transform(A, B):
    """Congruence transformation ``A.T @ B @ A``.

    Overloads (same dispatch table as the reference, plus the fused kernel case):
      * ``Diagonal, Diagonal`` -> ``Diagonal`` (congruence.py:20-22)
      * ``BlockDiagonalSparse, Diagonal | ScalarMul`` -> ``Diagonal`` by per-block segment sums (:25-43)
      * ``BlockDiagonalSparse, LazyKernel [+ diagonal]`` -> lazy ``ProjectedKernelGram`` (fused CUDA)
      * ``Dense, LazyKernel [+ diagonal]`` -> lazy ``ProjectedDenseGram`` (dense CUDA kernel product)
      * anything else -> ``A.T @ (B @ A)`` (:15-17)
    """
    if isinstance(A, Diagonal) and isinstance(B, Diagonal):
        return Diagonal(diag(A) ** 2 * diag(B))
    if isinstance(A, BlockDiagonalSparse) and isinstance(B, (Diagonal, ScalarMul)):
        vals = B @ A.nz_values**2
        return Diagonal(_fused.segment_sum(vals, A.shape[1]))
    if isinstance(A, BlockDiagonalSparse):
        split = _split_kernel_sum(B)
        if split is not None:
            return ProjectedKernelGram(A, split[0], split[1])
    if isinstance(A, Dense):
        split = _split_kernel_sum(B)
        if split is not None:
            return ProjectedDenseGram(A, split[0], split[1])
    if torch.is_tensor(A) and A.dim() == 1:
        return A @ (B @ A)
    return A.T @ (B @ A)
