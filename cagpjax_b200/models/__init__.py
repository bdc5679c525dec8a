"""Models (counterpart of ``cagpjax.models``).

``ComputationAwareGP`` keeps the reference's ``init`` / ``predict`` / ``prior_kl`` / ``variational_expectation`` /
``elbo`` surface; with block-sparse actions and a lazily evaluated kernel its ELBO and gradient are two fused n^2
passes on the GPU (``cagp.py``).  ``ComputationAwareGPState`` is the projected state ``init`` returns.
"""

from .cagp import ComputationAwareGP, ComputationAwareGPState

__all__ = ["ComputationAwareGP", "ComputationAwareGPState"]
