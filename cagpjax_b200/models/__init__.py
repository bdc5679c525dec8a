"""Gaussian process models (reference: models/__init__.py)."""

from .cagp import ComputationAwareGP, ComputationAwareGPState

__all__ = ["ComputationAwareGP", "ComputationAwareGPState"]
