"""Stand-in for `numpyro` (only the Distribution base class the reference derives from)."""
from . import distributions  # noqa: F401
