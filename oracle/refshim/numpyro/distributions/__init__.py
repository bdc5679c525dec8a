from . import constraints, distribution  # noqa: F401
from .distribution import Distribution, MultivariateNormal  # noqa: F401
