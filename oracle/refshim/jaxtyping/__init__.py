"""Stand-in for `jaxtyping` (the real package imports jax on first use of `Array`).

`Float[Array, "N D"]` is a class whose isinstance check accepts numpy arrays / numpy scalars with the annotated NUMBER
OF DIMENSIONS ("" = 0-d, "N" = 1-d, "N #K" = 2-d; a `*name` or `...` token makes the rank free).  The reference's
multiple dispatch relies on that: `ScalarFloat = float | Float[Array, ""]` must not match a vector
(linalg/utils.py `_add_jitter`)."""
import numpy as np


class _ArrayMeta(type):
    _cache = {}

    def __instancecheck__(cls, x):
        if not isinstance(x, (np.ndarray, np.generic)):
            return False
        nd = cls.__dict__.get("_ndim", None)
        return nd is None or np.ndim(x) == nd

    def __subclasscheck__(cls, sub):
        if sub is cls:
            return True
        if isinstance(sub, _ArrayMeta):
            return type.__subclasscheck__(cls, sub)
        return "_ndim" not in cls.__dict__ and isinstance(sub, type) and issubclass(sub, (np.ndarray, np.generic))

    def __getitem__(cls, item):
        spec = item[1] if isinstance(item, tuple) and len(item) > 1 and isinstance(item[1], str) else None
        if spec is None:
            return ArrayLike
        tokens = spec.split()
        ndim = None if any(t.startswith("*") or t == "..." for t in tokens) else len(tokens)
        key = ndim
        if key not in _ArrayMeta._cache:
            ns = {} if ndim is None else {"_ndim": ndim}
            _ArrayMeta._cache[key] = _ArrayMeta(f"ArrayLike[{'*' if ndim is None else ndim}d]", (ArrayLike,), ns)
        return _ArrayMeta._cache[key]


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
