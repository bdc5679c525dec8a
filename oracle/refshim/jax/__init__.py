"""Stand-in for `jax`: numpy execution, no tracing, no autodiff.  Only what the reference imports."""
import contextlib

import numpy as _np

import unittest as _unittest

from . import ffi, lax, random, scipy, test_util  # noqa: F401
from . import numpy  # noqa: F401  (jax.numpy)

Array = _np.ndarray


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = _np.dtype(dtype)


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def checkpoint(f=None, **kw):
    return f if f is not None else (lambda g: g)


remat = checkpoint


def jit(f=None, **kw):
    return f if f is not None else (lambda g: g)


def eval_shape(f, *args, **kw):
    r = f(*args, **kw)
    return ShapeDtypeStruct(r.shape, r.dtype)


def vmap(f, in_axes=0, out_axes=0):
    def mapped(*args):
        n = len(args[0])
        return _np.stack([f(*(a[i] for a in args)) for i in range(n)], axis=out_axes)

    return mapped


class custom_vjp:
    """No autodiff here: the decorated function runs as is; `defvjp` records the rules and nothing else."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        self.nondiff_argnums = nondiff_argnums

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


@contextlib.contextmanager
def default_matmul_precision(_):
    yield


@contextlib.contextmanager
def numpy_rank_promotion(_):
    yield


def _no_autodiff(*a, **k):
    """`jax.grad(f)` etc. return a callable that skips the calling test: the numpy stand-in has no autodiff (the golden
    generators differentiate the reference by central finite differences instead)."""

    def call(*aa, **kk):
        raise _unittest.SkipTest("no autodiff in the numpy stand-in of jax")

    return call


grad = value_and_grad = jacobian = jacfwd = jacrev = hessian = _no_autodiff


def jvp(*a, **k):
    raise _unittest.SkipTest("no autodiff in the numpy stand-in of jax")


vjp = linearize = jvp


def devices(backend=None):
    return ["cpu"]
