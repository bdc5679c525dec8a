"""Operator annotations specific to this package (reference: operators/annotations.py:6-9).

Annotations are marker classes stored in ``LinearOperator.annotations``; dispatch code asks ``op.isa(marker)``.
``ScaledOrthogonal`` is what lets ``orthogonalize`` / ``OrthogonalizationPolicy`` pass block-sparse actions and
Gram-Schmidt results through untouched, and what tells ``ComputationAwareGP`` that the k-sized closed-form ELBO is
numerically safe for a dense action matrix.
"""

from ..ops import Annotation


class ScaledOrthogonal(Annotation):
    """Columns are mutually orthogonal but not necessarily of unit norm (S^T S is diagonal)."""
