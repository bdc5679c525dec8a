"""Lazy kernel operator backed by the sm_100a kernels (reference: operators/lazy_kernel.py:11-103).

The reference evaluates K(x1, x2) @ X by sequential row batches of a dense kernel chunk under a
memory budget (``jax.lax.map`` + optional ``jax.checkpoint``).  Here the same operator never
materialises any chunk: products go to tiled CUDA kernels that evaluate kernel entries on the fly.
``max_memory_mb`` / ``batch_size`` / ``checkpoint`` are kept (and the derived ``batch_size_row`` /
``batch_size_col`` properties behave as in the reference) but they do not influence the result
or the memory footprint: the native kernels are recompute-by-construction.
"""

from __future__ import annotations

from typing import Optional

import torch

from .. import _fused
from ..gpx import AbstractKernel, unwrap
from ..ops import Dense, LinearOperator
from .block_diagonal_sparse import BlockDiagonalSparse


def _tensor_key(t: torch.Tensor):
    """Memoisation key of a tensor's CONTENT as far as torch can tell without reading it: storage address, layout
    and the in-place version counter.  ``id(tensor)`` (round 1) returned a stale projection after an optimiser's
    in-place update and can be reused after a free; the memo holds a strong reference to the tensor it was computed
    from, so the address cannot be recycled while the entry lives."""
    return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype, str(t.device))


_MEMO_LIMIT = 4  # entries kept per operator (a training loop creates one per step when it updates in place)


def _memo_put(memo: dict, key, value):
    memo[key] = value
    while len(memo) > _MEMO_LIMIT:
        memo.pop(next(iter(memo)))
    return value


def _kernel_params(kernel: AbstractKernel, like: torch.Tensor):
    ls = unwrap(kernel.lengthscale).to(device=like.device, dtype=like.dtype)
    var = unwrap(kernel.variance).to(device=like.device, dtype=like.dtype)
    return ls, var


class KernelProjection:
    """Row-sum-reducible statistics of K'(X, X) S for a Gram operator and block-sparse actions:
    A' = S^T K' S, B' = (K' S)^T (K' S), c' = (K' S)^T residual (unit-variance kernel K').
    One fused forward pass computes all three; their reverse mode is one fused backward pass."""

    def __init__(self, lazy: "LazyKernel", actions: BlockDiagonalSparse, residual: Optional[torch.Tensor]):
        self.lazy = lazy
        self.actions = actions
        self.residual = residual
        self._stats = None
        self._holder = {}

    def _compute(self):
        if self._stats is None:
            lz = self.lazy
            ls, _ = _kernel_params(lz.kernel, lz.x1)
            s = self.actions.nz_values
            yt = self.residual if self.residual is not None else torch.zeros_like(s)
            self._stats = _fused.project_stats(lz.kernel.name, lz.x1, ls, s.contiguous(), yt.contiguous(),
                                               self.actions.n_blocks, lz.row_shard_for(self.actions.n_blocks),
                                               self._holder)
        return self._stats

    @property
    def A(self):
        return self._compute()[0]

    @property
    def B(self):
        B = self._compute()[1]
        ready = self._holder.pop("B_event", None)  # B' was produced on a side stream (_fused.ProjectStats.forward)
        if ready is not None:
            torch.cuda.current_stream(B.device).wait_event(ready)
        return B

    @property
    def c(self):
        return self._compute()[2]

    @property
    def Mt(self) -> torch.Tensor:
        """Saved (k, m_local) projection M'^T of the forward pass (detached)."""
        self._compute()
        return self._holder["Mt"]


class DenseKernelProjection:
    """Statistics A' = S^T K' S, B' = (K'S)^T (K'S), c' = (K'S)^T residual for a DENSE action matrix S,
    computed by one dense kernel pass over this rank's rows (+ one all-reduce when sharded)."""

    def __init__(self, lazy: "LazyKernel", actions, residual: torch.Tensor):
        self.lazy, self.actions, self.residual = lazy, actions, residual
        self._stats = None
        self._holder = {}

    def _compute(self):
        if self._stats is None:
            lz = self.lazy
            ls, _ = _kernel_params(lz.kernel, lz.x1)
            n = lz.shape[0]
            shard = _fused.RowShard.full(n)
            if lz.process_group is not None:
                import torch.distributed as dist

                shard = _fused.RowShard.for_rank(n, dist.get_rank(lz.process_group),
                                                 dist.get_world_size(lz.process_group), lz.process_group)
            self._stats = _fused.dense_project_stats(lz.kernel.name, lz.x1, ls, self.actions.A.to(lz.dtype),
                                                     self.residual.contiguous(), shard, self._holder)
        return self._stats

    A = property(lambda self: self._compute()[0])
    B = property(lambda self: self._compute()[1])
    c = property(lambda self: self._compute()[2])


class KernelActions(LinearOperator):
    """Lazy (m, k) operator K(x1, x2) @ S for block-sparse S (the reference's ``cov_zx @ actions``,
    models/cagp.py:160).  Backed by the k-major projection ``Mt`` computed by one kernel pass."""

    def __init__(self, lazy: "LazyKernel", actions: BlockDiagonalSparse):
        super().__init__(lazy.dtype, (lazy.shape[0], actions.shape[1]), device=lazy.device)
        self.lazy = lazy
        self.actions = actions
        self._Mt = None

    def Mt(self) -> torch.Tensor:
        """(k, m) unit-variance projection, differentiable w.r.t. lengthscale and nz_values."""
        if self._Mt is None:
            lz = self.lazy
            proj = lz._projections.get(_tensor_key(self.actions.nz_values))
            if proj is not None and lz.is_gram and not torch.is_grad_enabled() and lz.process_group is None:
                self._Mt = proj.Mt
            else:
                ls, _ = _kernel_params(lz.kernel, lz.x1)
                self._Mt = _fused.ks_matrix(lz.kernel.name, lz.x1, lz.x2, ls, self.actions.nz_values.contiguous(),
                                            self.actions.n_blocks)
        return self._Mt

    def _variance(self):
        return _kernel_params(self.lazy.kernel, self.lazy.x1)[1]

    def _matmat(self, X):
        return self._variance() * (self.Mt().T @ X)

    def _rmatmat(self, X):
        return self._variance() * (X @ self.Mt().T)

    def to_dense(self):
        return self._variance() * self.Mt().T


class KernelDenseActions(LinearOperator):
    """Lazy (m, k) operator K(x1, x2) @ S for a DENSE action matrix S (n, k) (cola.ops.Dense actions,
    SURVEY 8a row a20).  The product M' = K'(x1, x2) S is computed once by the dense sm_100a kernel
    (2 m n k flops) and shared by init, predict and the ELBO; its reverse mode is two more passes."""

    def __init__(self, lazy: "LazyKernel", actions):
        super().__init__(lazy.dtype, (lazy.shape[0], actions.shape[1]), device=lazy.device)
        self.lazy = lazy
        self.actions = actions
        self._M = None

    def M(self) -> torch.Tensor:
        """(m, k) unit-variance product, differentiable w.r.t. lengthscale and the action matrix."""
        if self._M is None:
            lz = self.lazy
            ls, _ = _kernel_params(lz.kernel, lz.x1)
            self._M = _fused.kernel_matmat(lz.kernel.name, lz.x1, lz.x2, ls, self.actions.A.to(lz.dtype))
        return self._M

    def _variance(self):
        return _kernel_params(self.lazy.kernel, self.lazy.x1)[1]

    def _matmat(self, X):
        return self._variance() * (self.M() @ X)

    def _rmatmat(self, X):
        return self._variance() * (X @ self.M())

    def to_dense(self):
        return self._variance() * self.M()


class LazyKernel(LinearOperator):
    """K(x1, x2) as a matrix-free operator.

    Args:
        kernel: stationary kernel (``gpx.RBF`` / ``Matern12`` / ``Matern32`` / ``Matern52``).
        x1, x2: input points (M, D) and (N, D) on a CUDA device.
        max_memory_mb, batch_size, checkpoint: accepted for API parity (see module docstring).
    """

    def __init__(self, kernel: AbstractKernel, x1: torch.Tensor, x2: torch.Tensor, /, *, max_memory_mb: int = 2**10,
                 batch_size: Optional[int] = None, checkpoint: bool = False, process_group=None):
        shape = (x1.shape[0], x2.shape[0])
        dtype = kernel(x1[0, ...], x2[0, ...]).dtype  # operators/lazy_kernel.py:49
        super().__init__(dtype=dtype, shape=shape, device=x1.device)
        self.kernel = kernel
        self.x1 = x1
        self.x2 = x2
        self.max_memory_mb = max_memory_mb
        self.batch_size = batch_size
        self.checkpoint = checkpoint
        self.process_group = process_group
        self._projections = {}
        self._dense_products = {}

    def row_shard_for(self, n_blocks: int) -> _fused.RowShard:
        """Row partition of this (Gram) operator over the ranks of ``process_group``:
        aligned to action-block boundaries when there are at least as many blocks as ranks."""
        n = self.shape[0]
        if self.process_group is None:
            return _fused.RowShard.full(n)
        import torch.distributed as dist

        rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        if n_blocks >= world:
            return _fused.RowShard.for_rank_blocks(n, n_blocks, rank, world, self.process_group)
        return _fused.RowShard.for_rank(n, rank, world, self.process_group)

    # -- reference batching knobs (operators/lazy_kernel.py:59-77) ---------------------------
    @property
    def max_elements(self) -> int:
        return int(self.max_memory_mb * 2**20) // torch.empty((), dtype=self.dtype).element_size()

    @property
    def batch_size_row(self) -> int:
        if self.batch_size is not None:
            return self.batch_size
        return max(1, self.max_elements // self.shape[1])

    @property
    def batch_size_col(self) -> int:
        if self.batch_size is not None:
            return self.batch_size
        return max(1, self.max_elements // self.shape[0])

    # -- structure -----------------------------------------------------------------------------
    @property
    def is_gram(self) -> bool:
        return self.x1 is self.x2

    @property
    def T(self) -> "LazyKernel":
        if self.is_gram:
            return self
        return LazyKernel(self.kernel, self.x2, self.x1, max_memory_mb=self.max_memory_mb, batch_size=self.batch_size,
                          checkpoint=self.checkpoint)

    def _diag(self) -> torch.Tensor:
        if self.shape[0] != self.shape[1]:
            raise ValueError("diagonal of a non-square kernel operator")
        if self.is_gram:
            _, var = _kernel_params(self.kernel, self.x1)
            return var * torch.ones(self.shape[0], dtype=self.dtype, device=self.device)
        ls, var = _kernel_params(self.kernel, self.x1)
        sq = (((self.x1 - self.x2) / ls) ** 2).sum(dim=1)
        return var * self.kernel._profile(sq)

    def project(self, actions: BlockDiagonalSparse, residual: Optional[torch.Tensor] = None) -> KernelProjection:
        """Fused statistics of K'(X, X) S (Gram operators only); memoised per action tensor."""
        if not self.is_gram:
            raise ValueError("project() needs a Gram operator (x1 is x2)")
        key = _tensor_key(actions.nz_values)
        proj = self._projections.get(key)
        if proj is None or (residual is not None and proj.residual is not residual):
            proj = _memo_put(self._projections, key, KernelProjection(self, actions, residual))
        return proj

    # -- products -----------------------------------------------------------------------------
    def __matmul__(self, other):
        if isinstance(other, BlockDiagonalSparse):
            return KernelActions(self, other)
        if isinstance(other, Dense):
            return self.dense_actions(other)
        return super().__matmul__(other)

    def project_dense(self, actions, residual: torch.Tensor) -> DenseKernelProjection:
        """Fused statistics of K'(X, X) S for a dense S (Gram operators only)."""
        if not self.is_gram:
            raise ValueError("project_dense() needs a Gram operator (x1 is x2)")
        key = ("dense",) + _tensor_key(actions.A)
        proj = self._projections.get(key)
        if proj is None or proj.residual is not residual:
            proj = _memo_put(self._projections, key, DenseKernelProjection(self, actions, residual))
        return proj

    def dense_actions(self, actions) -> KernelDenseActions:
        """K @ S for dense S, memoised per action tensor so that one kernel pass serves every consumer."""
        key = _tensor_key(actions.A)
        op = self._dense_products.get(key)
        if op is None:
            op = _memo_put(self._dense_products, key, KernelDenseActions(self, actions))
        return op

    def _matmat(self, X: torch.Tensor) -> torch.Tensor:
        """K(x1, x2) @ X for a dense X (n, p) (operators/lazy_kernel.py:79-90)."""
        ls, var = _kernel_params(self.kernel, self.x1)
        return var * _fused.kernel_matmat(self.kernel.name, self.x1, self.x2, ls, X.to(self.dtype))

    def _rmatmat(self, X: torch.Tensor) -> torch.Tensor:
        """X @ K(x1, x2) = (K(x2, x1) @ X^T)^T (operators/lazy_kernel.py:92-103)."""
        ls, var = _kernel_params(self.kernel, self.x1)
        return var * _fused.kernel_matmat(self.kernel.name, self.x2, self.x1, ls, X.T.to(self.dtype)).T
