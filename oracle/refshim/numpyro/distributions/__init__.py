from . import constraints, distribution  # noqa: F401
from .distribution import Distribution  # noqa: F401
