"""Kernel compute engines (counterpart of ``cagpjax.computations``).

``LazyKernelComputation`` is the drop-in plug-in point of this package: a kernel constructed with
``compute_engine=LazyKernelComputation(...)`` hands out ``LazyKernel`` operators whose products run on the GPU.
"""

from .lazy import LazyKernelComputation

__all__ = ["LazyKernelComputation"]
