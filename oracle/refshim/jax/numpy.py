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
        out = fn(*a, **k)
        # jax: functions of Python scalars give WEAKLY typed results (they do not promote float32 operands later)
        if isinstance(out, _np.floating) and a and all(isinstance(x, (int, float)) and not isinstance(x, _np.generic) for x in a):
            return _WeakFloat(out)
        return _wrap(out)

    call.__name__ = getattr(fn, "__name__", "fn")
    return call


class _WeakFloat(float):
    """Result of a jnp function applied to Python scalars: weakly typed (a Python float to numpy's promotion rules),
    with the array methods the reference calls on it."""

    def astype(self, dtype):
        return _np.asarray(float(self), dtype=dtype)[()]

    @property
    def dtype(self):
        return _np.dtype(_np.float64)

    ndim = 0
    shape = ()


class _Linalg:
    def __getattr__(self, name):
        obj = getattr(_linalg, name)
        return _forward(obj) if callable(obj) and not isinstance(obj, type) else obj

    @staticmethod
    def lstsq(a, b, rcond=None):
        """jnp.linalg.lstsq: SVD solution with rcond = eps(dtype) * max(m, n) by default."""
        a, b = _np.asarray(a), _np.asarray(b)
        u, s, vt = _linalg.svd(a, full_matrices=False)
        if rcond is None:
            rcond = _np.finfo(a.dtype).eps * max(a.shape)
        keep = s > rcond * s[0]
        sinv = _np.where(keep, 1.0 / _np.where(keep, s, 1), 0).astype(a.dtype)
        x = vt.T @ (sinv[:, None] * (u.T @ b) if b.ndim == 2 else sinv * (u.T @ b))
        return _wrap(x), None, int(keep.sum()), _wrap(s)

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
