"""Stand-in for `cola` (CoLA): the operator algebra and plum-style dispatch the reference uses, on numpy."""
from . import annotations, linalg, ops  # noqa: F401
from ._dispatch import dispatch  # noqa: F401
from .annotations import PSD, SelfAdjoint, Stiefel, Unitary  # noqa: F401
from .linalg import diag  # noqa: F401
from .ops import LinearOperator, densify, lazify  # noqa: F401
