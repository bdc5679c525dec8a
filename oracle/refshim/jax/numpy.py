"""Stand-in for `jax.numpy`: numpy with the behavioural differences the reference relies on.

* arrays are `JArray`s: jax arrays are immutable, so `q -= x` REBINDS `q` (it must not write through a view into the
  array `q` was sliced from) and updates are spelled `a.at[idx].set(v)`;
* `isscalar` is also True for zero-dimensional arrays;
* `linalg.eigh(a, symmetrize_input=True)`.
Every function of numpy is forwarded with its array results re-wrapped as `JArray`."""
import numpy as _np
from numpy import linalg as _linalg


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] = v
        return out.view(JArray)

    def add(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] += v
        return out.view(JArray)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class JArray(_np.ndarray):
    """ndarray with jax's value semantics for augmented assignment and the `.at[...]` update syntax."""

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, JArray):
        return x.view(JArray)
    if isinstance(x, tuple) and type(x) is not tuple:  # namedtuple results (qr, eigh, slogdet ...)
        return type(x)(*[_wrap(e) for e in x])
    if isinstance(x, (tuple, list)):
        return type(x)(_wrap(e) for e in x)
    return x


def _forward(fn):
    def call(*a, **k):
        k.pop("precision", None)
        return _wrap(fn(*a, **k))

    call.__name__ = getattr(fn, "__name__", "fn")
    return call


class _Linalg:
    def __getattr__(self, name):
        obj = getattr(_linalg, name)
        return _forward(obj) if callable(obj) and not isinstance(obj, type) else obj

    @staticmethod
    def eigh(a, UPLO=None, symmetrize_input=True):
        a = _np.asarray(a)
        if symmetrize_input:
            a = 0.5 * (a + a.T)
        return _wrap(_linalg.eigh(a))


linalg = _Linalg()


def isscalar(x):
    return _np.isscalar(x) or (isinstance(x, _np.ndarray) and x.ndim == 0)


def array(x, dtype=None, **kw):
    return _np.array(x, dtype=dtype).view(JArray)


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        return _forward(obj)
    return obj
