"""Multiple dispatch in the style of plum (which CoLA re-exports as `cola.dispatch`).

Two registration styles occur in the reference:
  * `typing_extensions.overload`-decorated implementations followed by one `@cola.dispatch` definition whose body is
    `pass` (linalg/congruence.py, lower_cholesky.py, utils.py, orthogonalize.py): the overloads are the methods;
  * repeated `@cola.dispatch(precedence=p)` definitions of one name (linalg/eigh.py): each is a method.
Resolution: all methods whose annotations accept the positional arguments; the most specific one wins (a signature is
more specific when each of its parameter types is a subtype of the other's); among incomparable candidates the higher
precedence wins.
"""
import inspect
import types
import typing

import typing_extensions

_functions = {}


def _members(tp):
    if tp is typing.Any or tp is inspect.Parameter.empty:
        return None  # anything
    origin = typing.get_origin(tp)
    if origin is typing.Union or isinstance(tp, types.UnionType):
        out = []
        for a in typing.get_args(tp):
            m = _members(a)
            if m is None:
                return None
            out.extend(m)
        return out
    if tp is None:
        return [type(None)]
    if tp is float:
        return [float, int]
    return [tp]


def _accepts(tp, value):
    m = _members(tp)
    if m is None:
        return True
    return any(isinstance(value, c) for c in m)


def _subtype(a, b):
    ma, mb = _members(a), _members(b)
    if mb is None:
        return True
    if ma is None:
        return False
    return all(any(issubclass(x, y) for y in mb) for x in ma)


class Method:
    def __init__(self, fn, precedence):
        self.fn, self.precedence = fn, precedence
        hints = typing.get_type_hints(fn)
        self.types = [hints.get(p.name, typing.Any) for p in inspect.signature(fn).parameters.values()
                      if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]

    def accepts(self, args):
        return len(args) <= len(self.types) and all(_accepts(t, a) for t, a in zip(self.types, args))

    def more_specific_than(self, other):
        return all(_subtype(a, b) for a, b in zip(self.types, other.types))


class Function:
    def __init__(self, name):
        self.name, self.methods = name, []

    def dispatch(self, fn=None, *, precedence=0):
        """plum's `Function.dispatch`: register one more method from outside the defining module."""
        if fn is None:
            return lambda f: self.dispatch(f, precedence=precedence)
        self.methods.append(Method(fn, precedence))
        return self

    def __call__(self, *args, **kw):
        cands = [m for m in self.methods if m.accepts(args)]
        if not cands:
            raise TypeError(f"{self.name}: no method for {[type(a).__name__ for a in args]}")
        best = [m for m in cands if not any(o is not m and o.more_specific_than(m) and not m.more_specific_than(o) for o in cands)]
        if len(best) > 1:
            top = max(m.precedence for m in best)
            best = [m for m in best if m.precedence == top]
        if len(best) > 1:
            raise TypeError(f"{self.name}: ambiguous call for {[type(a).__name__ for a in args]}")
        return best[0].fn(*args, **kw)


def dispatch(fn=None, *, precedence=0):
    if fn is None:
        return lambda f: dispatch(f, precedence=precedence)
    key = (fn.__module__, fn.__qualname__)
    func = _functions.setdefault(key, Function(fn.__qualname__))
    overloads = typing_extensions.get_overloads(fn)
    if overloads:
        if not func.methods:
            func.methods = [Method(o, 0) for o in overloads]
    else:
        func.methods.append(Method(fn, precedence))
    func.__doc__ = fn.__doc__
    return func
