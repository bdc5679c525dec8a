"""Block-diagonal sparse action operator (reference: operators/block_diagonal_sparse.py:10-95)."""

from __future__ import annotations

import torch

from .. import _fused
from ..ops import LinearOperator
from .annotations import ScaledOrthogonal


class BlockDiagonalSparse(LinearOperator):
    """(n, n_blocks) matrix with exactly one non-zero per row, in contiguous blocks.

    ``block_size = n // n_blocks``; if ``n % n_blocks != 0`` the last block absorbs the remainder
    (block_diagonal_sparse.py:49-52).  The dense matrix is never formed on the hot path: products
    with a ``LazyKernel`` go to the fused sm_100a kernels (see ``LazyKernel.__matmul__``).

    Example: ``BlockDiagonalSparse(torch.arange(1., 7.), n_blocks=3) @ torch.eye(3)`` is
    ``[[1,0,0],[2,0,0],[0,3,0],[0,4,0],[0,0,5],[0,0,6]]``.
    """

    def __init__(self, nz_values: torch.Tensor, n_blocks: int):
        n = nz_values.shape[0]
        super().__init__(nz_values.dtype, (n, n_blocks), annotations={ScaledOrthogonal}, device=nz_values.device)
        self.nz_values = nz_values

    @property
    def n_blocks(self) -> int:
        return self.shape[1]

    def _matmat(self, X: torch.Tensor) -> torch.Tensor:
        """S @ X, X (n_blocks, m) -> (n, m): row i is nz_i * X[blk(i)]  (block_diagonal_sparse.py:48-70)."""
        n, n_blocks = self.shape
        block_size, n_blocks_main, n_main = _fused.block_layout(n, n_blocks)
        m = X.shape[1]
        blocks_main = self.nz_values[:n_main].reshape(n_blocks_main, block_size)
        res = (blocks_main[..., None] * X[:n_blocks_main, None, :]).reshape(n_main, m)
        if n > n_main:
            res = torch.cat([res, torch.outer(self.nz_values[n_main:], X[n_blocks_main, :])], dim=0)
        return res

    def _rmatmat(self, X: torch.Tensor) -> torch.Tensor:
        """X @ S, X (m, n) -> (m, n_blocks): per-block weighted sums  (block_diagonal_sparse.py:72-95)."""
        n, n_blocks = self.shape
        block_size, n_blocks_main, n_main = _fused.block_layout(n, n_blocks)
        m = X.shape[0]
        blocks_main = self.nz_values[:n_main].reshape(n_blocks_main, block_size)
        res = torch.einsum("ik,jik->ji", blocks_main, X[:, :n_main].reshape(m, n_blocks_main, block_size))
        if n > n_main:
            res = torch.cat([res, (X[:, n_main:] @ self.nz_values[n_main:])[:, None]], dim=1)
        return res
