"""B200LazyKernel: the reference's LazyKernel (operators/lazy_kernel.py:11-103) with its products on the sm_100a
kernels.  Never run on JAX (none in this image); run under numpy stand-ins with emulated FFI targets
(tests/test_jax_binding_wiring.py)."""
import jax.numpy as jnp
from cagpjax.operators import BlockDiagonalSparse, LazyKernel

from . import ops
from .ffi import FAMILIES


def _family(kernel) -> int:
    name = type(kernel).__name__
    if name not in FAMILIES:
        raise NotImplementedError(f"B200 kernels implement {sorted(FAMILIES)}; got {name}")
    return FAMILIES[name]


class B200LazyKernel(LazyKernel):
    """K(x1, x2) as a matrix-free operator whose products never materialise a kernel chunk."""

    def _scaled(self):
        fam = _family(self.kernel)
        ls = jnp.atleast_1d(self.kernel.lengthscale.value).astype(self.x2.dtype)
        origin = 0.5 * (self.x2.min(axis=0) + self.x2.max(axis=0))
        xs2 = ops.scale_inputs(self.x2, ls, fam, origin)
        xs1 = xs2 if self.x1 is self.x2 else ops.scale_inputs(self.x1, ls, fam, origin)
        return fam, ls, xs1, xs2

    def _matmat(self, X):  # operators/lazy_kernel.py:79-90
        fam, ls, xs1, xs2 = self._scaled()
        return self.kernel.variance.value * ops.kernel_matmat(xs1, xs2, X.astype(self.dtype), ls, fam, self.x2.shape[1])

    def _rmatmat(self, X):  # operators/lazy_kernel.py:92-103: X @ K(x1, x2) = (K(x2, x1) @ X^T)^T
        fam, ls, xs1, xs2 = self._scaled()
        return self.kernel.variance.value * ops.kernel_matmat(xs2, xs1, X.T.astype(self.dtype), ls, fam, self.x2.shape[1]).T

    def project(self, actions: BlockDiagonalSparse, residual):
        """(A', B', c') of SURVEY 3.2 for block-sparse actions (Gram operators only): one fused n^2 pass."""
        if self.x1 is not self.x2:
            raise ValueError("project() needs a Gram operator")
        return ops.project_stats(self.x1, self.kernel.lengthscale.value, actions.nz_values, residual,
                                 _family(self.kernel), actions.shape[1])
