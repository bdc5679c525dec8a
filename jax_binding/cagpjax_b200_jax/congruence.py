"""plum overload: S^T (K + sigma^2 I) S for block-sparse S and a B200LazyKernel goes to the fused projection instead
of the dense fallback ``A.T @ (B @ A)`` (linalg/congruence.py:15-17).  Never run on JAX (none in this
image); the overload is exercised under numpy stand-ins with emulated FFI targets: tests/test_jax_binding_wiring.py."""
import cola
import jax.numpy as jnp
from cagpjax.linalg.congruence import congruence_transform
from cagpjax.operators import BlockDiagonalSparse

from .operators import B200LazyKernel


def _split(B: cola.ops.Sum):
    lazy = [m for m in B.Ms if isinstance(getattr(m, "A", m), B200LazyKernel)]
    rest = [m for m in B.Ms if m not in lazy]
    return (getattr(lazy[0], "A", lazy[0]) if len(lazy) == 1 else None), rest


@congruence_transform.dispatch
def _congruence_b200(A: BlockDiagonalSparse, B: cola.ops.Sum):
    lazy, rest = _split(B)
    if lazy is None:
        return A.T @ (B @ A)
    k = A.shape[1]
    Ap, _, _ = lazy.project(A, jnp.zeros(A.shape[0], lazy.dtype))
    out = cola.ops.Dense(lazy.kernel.variance.value * Ap)
    for term in rest:  # ScalarMul / Diagonal noise terms keep their structured overload
        out = out + congruence_transform(A, term)
    assert out.shape == (k, k)
    return out
