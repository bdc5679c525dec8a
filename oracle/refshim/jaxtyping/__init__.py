"""Stand-in for `jaxtyping` (the real package imports jax on first use of `Array`).  Annotations only."""
import numpy as np


class _ArrayMeta(type):
    def __instancecheck__(cls, x):
        return isinstance(x, (np.ndarray, np.generic))

    def __subclasscheck__(cls, sub):
        return sub is cls or (isinstance(sub, type) and issubclass(sub, (np.ndarray, np.generic)))

    def __getitem__(cls, item):  # Float[Array, "N D"] -> the same array-like marker
        return cls


class ArrayLike(metaclass=_ArrayMeta):
    """Matches numpy arrays and numpy scalars (what a jax.Array is under the numpy stand-in)."""


Array = ArrayLike
Float = Int = Bool = Shaped = Num = Inexact = Integer = ArrayLike


class PRNGKeyArray:
    pass


class _KeyMeta(type):
    def __getitem__(cls, item):
        return cls


class Key(PRNGKeyArray, metaclass=_KeyMeta):
    pass


class _PyTreeMeta(type):
    def __getitem__(cls, item):
        return cls


class PyTree(metaclass=_PyTreeMeta):
    pass


Scalar = ArrayLike
