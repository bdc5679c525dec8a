"""Stand-in for `equinox`: `Module` = a class whose annotated attributes become dataclass fields (not frozen),
`field(static=..., default=..., default_factory=..., init=...)`, `__check_init__` run after construction."""
import abc
import dataclasses


def field(*, static=False, default=dataclasses.MISSING, default_factory=dataclasses.MISSING, init=True, converter=None, **kw):
    return dataclasses.field(default=default, default_factory=default_factory, init=init)


class _ModuleMeta(abc.ABCMeta):
    def __new__(mcs, name, bases, ns, **kw):
        own_init = "__init__" in ns
        cls = super().__new__(mcs, name, bases, ns, **kw)
        cls = dataclasses.dataclass(init=not own_init, repr=False, eq=False)(cls)
        user_init = cls.__init__

        def __init__(self, *a, **k):
            outer = not getattr(self, "_eqx_in_init", False)
            object.__setattr__(self, "_eqx_in_init", True)
            user_init(self, *a, **k)
            if outer:
                object.__setattr__(self, "_eqx_in_init", False)
                seen = set()
                for klass in type(self).__mro__:
                    chk = klass.__dict__.get("__check_init__")
                    if chk is not None and chk not in seen:
                        seen.add(chk)
                        chk(self)

        __init__.__wrapped__ = user_init
        cls.__init__ = __init__
        return cls


class Module(metaclass=_ModuleMeta):
    pass


def filter_jit(f=None, **kw):
    return f if f is not None else (lambda g: g)
