"""Linear operators of the CaGP path, on torch tensors.

Counterpart of ``cagpjax.operators``: ``LazyKernel`` (a kernel matrix that is never materialised - products go to
the sm_100a kernels), ``BlockDiagonalSparse`` (the action matrix of ``BlockSparsePolicy``, stored as its n
non-zeros) and ``diag_like``.  The generic operator algebra they plug into lives in ``cagpjax_b200.ops``.
"""

from .block_diagonal_sparse import BlockDiagonalSparse
from .diag_like import diag_like
from .lazy_kernel import LazyKernel

__all__ = ["BlockDiagonalSparse", "LazyKernel", "diag_like"]
