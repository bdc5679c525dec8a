"""Stand-in for `jax.numpy`: numpy with the few behavioural differences the reference relies on."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import linalg as _linalg


def isscalar(x):
    """jnp.isscalar is also True for zero-dimensional arrays (numpy's is not)."""
    return _np.isscalar(x) or (isinstance(x, _np.ndarray) and x.ndim == 0)


class _Linalg:
    def __getattr__(self, name):
        return getattr(_linalg, name)

    @staticmethod
    def eigh(a, UPLO=None, symmetrize_input=True):
        a = _np.asarray(a)
        if symmetrize_input:
            a = 0.5 * (a + a.T)
        return _linalg.eigh(a)


linalg = _Linalg()


def array(x, dtype=None, **kw):
    return _np.array(x, dtype=dtype)


def dot(a, b, precision=None):
    return _np.dot(a, b)
