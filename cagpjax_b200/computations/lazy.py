"""Lazy kernel compute engine (reference: computations/lazy.py:15-124).

This is the drop-in plug-in point: ``RBF(compute_engine=LazyKernelComputation(...))``.  ``gram``
returns the PSD-annotated ``LazyKernel(kernel, x, x)`` and ``cross_covariance`` the rectangular
``LazyKernel(kernel, x1, x2)``, both executed by the sm_100a kernels.  ``process_group`` (optional)
partitions the rows of the Gram operator across the ranks of a process group.
"""

from __future__ import annotations

from typing import Optional

from ..gpx import AbstractKernelComputation
from ..operators import LazyKernel
from ..ops import PSD


class LazyKernelComputation(AbstractKernelComputation):
    def __init__(self, *, batch_size: Optional[int] = None, max_memory_mb: int = 2**10, checkpoint: bool = True,
                 process_group=None):
        self.batch_size = batch_size
        self.max_memory_mb = max_memory_mb
        self.checkpoint = checkpoint
        self.process_group = process_group

    def gram(self, kernel, x):
        return PSD(LazyKernel(kernel, x, x, batch_size=self.batch_size, max_memory_mb=self.max_memory_mb,
                              checkpoint=self.checkpoint, process_group=self.process_group))

    def cross_covariance(self, kernel, x1, x2):
        return LazyKernel(kernel, x1, x2, batch_size=self.batch_size, max_memory_mb=self.max_memory_mb,
                          checkpoint=self.checkpoint)
