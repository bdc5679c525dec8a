from . import computations  # noqa: F401
from .base import AbstractKernel  # noqa: F401
from .stationary import RBF, Matern12, Matern32, Matern52  # noqa: F401
