"""Stand-in for `jax.random`: keys are immutable VALUES (the same key always gives the same draw), numpy bit streams
(not JAX's threefry: goldens store their inputs)."""
import numpy as _np


class _Key:
    def __init__(self, seed):
        self.seed = tuple(int(s) for s in (seed if isinstance(seed, (tuple, list)) else (seed,)))

    def rng(self):
        return _np.random.default_rng(self.seed)

    def __repr__(self):
        return f"Key{self.seed}"


def key(seed):
    return _Key(seed)


PRNGKey = key


def split(k, num=2):
    return [_Key(k.seed + (i,)) for i in range(num)]


def fold_in(k, data):
    return _Key(k.seed + (int(data),))


def _out(x, dtype):
    from .numpy import _wrap

    return _wrap(_np.asarray(x).astype(dtype or _np.float64))


def normal(k, shape=(), dtype=None):
    return _out(k.rng().standard_normal(shape), dtype)


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _out(minval + (maxval - minval) * k.rng().random(shape), dtype)


def rademacher(k, shape=(), dtype=None):
    return _out(2 * k.rng().integers(0, 2, size=shape) - 1, dtype or _np.int32)
