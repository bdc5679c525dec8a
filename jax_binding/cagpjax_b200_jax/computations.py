"""B200KernelComputation: drop-in for cagpjax.computations.LazyKernelComputation (computations/lazy.py:15-124).
Never run on JAX (none in this image); run under numpy stand-ins with emulated FFI targets: tests/test_jax_binding_wiring.py."""
import cola
import lineax as lx
from cagpjax.interop import ColaLinearOperator
from gpjax.kernels.computations import AbstractKernelComputation

from .operators import B200LazyKernel


class B200KernelComputation(AbstractKernelComputation):
    """``RBF(compute_engine=B200KernelComputation())``: Gram and cross-covariance operators backed by sm_100a kernels.
    ``batch_size`` / ``max_memory_mb`` / ``checkpoint`` are accepted for API parity and ignored (the kernels are
    recompute-by-construction)."""

    def __init__(self, *, batch_size=None, max_memory_mb=2**10, checkpoint=True):
        self.batch_size, self.max_memory_mb, self.checkpoint = batch_size, max_memory_mb, checkpoint

    def gram(self, kernel, x):  # computations/lazy.py:88-106
        op = cola.PSD(B200LazyKernel(kernel, x, x, batch_size=self.batch_size, max_memory_mb=self.max_memory_mb,
                                     checkpoint=self.checkpoint))
        return lx.TaggedLinearOperator(ColaLinearOperator(op), lx.positive_semidefinite_tag)

    def cross_covariance(self, kernel, x1, x2):  # computations/lazy.py:108-124
        return ColaLinearOperator(B200LazyKernel(kernel, x1, x2, batch_size=self.batch_size,
                                                 max_memory_mb=self.max_memory_mb, checkpoint=self.checkpoint))
