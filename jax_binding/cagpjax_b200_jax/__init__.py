"""JAX binding of the B200-native CaGP hot path (see ../README.md).  Needs jax, gpjax, cola and cagpjax."""
try:
    import jax  # noqa: F401
except ImportError as exc:  # pragma: no cover - this image has no JAX
    raise ImportError("cagpjax_b200_jax needs jax (and gpjax, cola-ml, cagpjax); the torch host layer that runs in "
                      "this repository is the package `cagpjax_b200`") from exc

from .computations import B200KernelComputation  # noqa: E402
from .operators import B200LazyKernel  # noqa: E402
from . import congruence as _congruence  # noqa: E402,F401  (registers the plum overload)

__all__ = ["B200KernelComputation", "B200LazyKernel"]
