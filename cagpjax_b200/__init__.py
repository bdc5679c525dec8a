"""B200-native computation-aware Gaussian processes: the CAGPJax hot path on sm_100a.

Mirrors the public surface of ``cagpjax`` (reference: src/cagpjax/__init__.py:3-13) for the
kernel-projection path; ``cagpjax_b200.gpx`` holds stand-ins for the GPJax objects it consumes.
"""

from . import computations, distributions, gpx, linalg, models, operators, policies, solvers
from .models import ComputationAwareGP
from .operators import BlockDiagonalSparse
from .policies import BlockSparsePolicy

__version__ = "0.1.0"
__all__ = ["BlockDiagonalSparse", "BlockSparsePolicy", "ComputationAwareGP"]
