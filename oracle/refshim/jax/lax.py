"""Stand-in for `jax.lax` (eager Python control flow)."""
import numpy as _np


def map(f, xs, *, batch_size=None):  # noqa: A001  (jax.lax.map: batching only changes memory use, not results)
    from .numpy import _wrap

    return _wrap(_np.stack([f(_wrap(x)) for x in xs]))


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(lower, upper):
        val = body_fun(i, val)
    return val


def scan(f, init, xs, length=None):
    from .numpy import _wrap

    n = len(xs[0]) if isinstance(xs, (tuple, list)) else len(xs)
    carry, ys = init, []
    for i in range(n):
        x = tuple(_wrap(a[i]) for a in xs) if isinstance(xs, (tuple, list)) else _wrap(xs[i])
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (None if not ys or ys[0] is None else _wrap(_np.stack(ys)))


def stop_gradient(x):
    return x


class Precision:
    HIGHEST = "highest"
