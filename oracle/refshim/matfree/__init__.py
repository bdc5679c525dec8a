"""Stand-in for `matfree` (import only; the Lanczos recurrence is outside the pinned path)."""
from . import decomp  # noqa: F401
