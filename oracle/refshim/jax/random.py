"""Stand-in for `jax.random` (different bit streams than JAX's threefry: goldens store their inputs)."""
import numpy as _np


def key(seed):
    return _np.random.default_rng(int(seed))


PRNGKey = key


def split(k, num=2):
    return [_np.random.default_rng(s) for s in k.bit_generator.seed_seq.spawn(num)] if hasattr(k.bit_generator, "seed_seq") else [k] * num


def normal(k, shape=(), dtype=None):
    return k.standard_normal(shape).astype(dtype or _np.float64)


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return (minval + (maxval - minval) * k.random(shape)).astype(dtype or _np.float64)
