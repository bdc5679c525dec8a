"""Annotations for operators (reference: operators/annotations.py:6-9)."""

from ..ops import Annotation


class ScaledOrthogonal(Annotation):
    """Operator whose columns are orthogonal (but not necessarily orthonormal)."""
