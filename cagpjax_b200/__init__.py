"""B200-native computation-aware Gaussian processes: the CAGPJax hot path on sm_100a.

Package map (same module names as ``cagpjax``, reference: src/cagpjax/__init__.py:3-13):

  models, policies, solvers, operators, linalg, computations, distributions
      the reference's public surface for the kernel-projection path, on torch CUDA tensors;
  gpx, ops
      stand-ins for the parts of GPJax (kernels, likelihood, prior / posterior, ``fit``) and CoLA (operator algebra)
      that path consumes - neither library can be installed next to this package;
  _native, _fused, _closed_form
      the ctypes binding of ``libcagp_b200.so`` (C ABI in include/cagp_b200.h), the custom reverse modes and row
      sharding built on it, and the k-sized closed-form ELBO.

There is no CPU fallback: operators raise ``_native.NativeLibraryError`` when the library or a CUDA device is missing.
"""

from . import computations, distributions, gpx, linalg, models, operators, policies, solvers
from .models import ComputationAwareGP, ComputationAwareGPState
from .operators import BlockDiagonalSparse, LazyKernel
from .policies import BlockSparsePolicy

__version__ = "0.1.0"
__all__ = ["BlockDiagonalSparse", "BlockSparsePolicy", "ComputationAwareGP", "ComputationAwareGPState", "LazyKernel"]
