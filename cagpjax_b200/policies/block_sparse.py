"""Block-sparse policy (reference: policies/block_sparse.py:18-108)."""

from __future__ import annotations

from typing import Callable, Optional

import torch

from .. import _fused
from ..gpx import Parameter, unwrap
from ..operators import BlockDiagonalSparse
from .base import AbstractBatchLinearSolverPolicy


def _normalize_by_blocks(values: torch.Tensor, n_actions: int) -> torch.Tensor:
    """Scale every action block to unit L2 norm; zero blocks are left alone
    (``where(norm > 0, norm, 1)``, block_sparse.py:18-40)."""
    n = values.shape[0]
    block_size, n_blocks_main, n_main = _fused.block_layout(n, n_actions)
    chunks = []
    if n_main > 0:
        main = values[:n_main].reshape(n_blocks_main, block_size)
        norms = torch.linalg.vector_norm(main, dim=1, keepdim=True)
        chunks.append((main / torch.where(norms > 0, norms, torch.ones_like(norms))).reshape(n_main))
    if n > n_main:
        overhang = values[n_main:]
        norm = torch.linalg.vector_norm(overhang)
        chunks.append(overhang / torch.where(norm > 0, norm, torch.ones_like(norm)))
    if not chunks:
        return values
    return torch.cat(chunks)


def normal_sampler(key, shape, dtype=None):
    """Default sampler: standard normal from a ``torch.Generator`` (or an int seed)."""
    gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
    return torch.randn(shape, generator=gen, dtype=dtype or torch.float64)


class BlockSparsePolicy(AbstractBatchLinearSolverPolicy):
    """Fixed block-diagonal sparse action structure with learnable non-zeros ``nz_values``."""

    def __init__(self, n_actions: int, nz_values):
        self.n_actions = n_actions
        self.nz_values = nz_values
        self._check_init()

    @classmethod
    def from_random(cls, key, num_datapoints: int, n_actions: int, *,
                    sampler: Callable = normal_sampler, dtype=None, device=None) -> "BlockSparsePolicy":
        """Initialise from block-normalised random samples (block_sparse.py:68-93)."""
        if num_datapoints < 1:
            raise ValueError("num_datapoints must be at least 1")
        nz_values = sampler(key, (num_datapoints,), dtype)
        nz_values = _normalize_by_blocks(nz_values, n_actions)
        if device is not None:
            nz_values = nz_values.to(device)
        return cls(n_actions=n_actions, nz_values=nz_values)

    def to_actions(self, A, *, key=None) -> BlockDiagonalSparse:
        """``A`` and ``key`` are unused; no re-normalisation happens here (block_sparse.py:95-108)."""
        nz = unwrap(self.nz_values)
        if A is not None and hasattr(A, "device") and nz.device != A.device:
            nz = nz.to(A.device)
        return BlockDiagonalSparse(nz, self.n_actions)
