"""Stand-in for `jax.ffi`: targets are recorded; `ffi_call(name, out_spec)` runs a HOST EMULATION registered with
`emulate(name, fn)` (tests register numpy restatements of the C-ABI entry points' semantics) - there is no XLA here."""
_targets = {}
_emulations = {}


def pycapsule(fn):
    return fn


def register_ffi_target(name, capsule, platform="cpu", api_version=1, **kw):
    _targets[name] = (capsule, platform)


def emulate(name, fn):
    _emulations[name] = fn


def ffi_call(target_name, result_shape_dtypes, **call_kw):
    def call(*args, **attrs):
        if target_name not in _targets:
            raise RuntimeError(f"FFI target {target_name} was never registered")
        if target_name not in _emulations:
            raise NotImplementedError(f"no host emulation registered for FFI target {target_name}")
        from .numpy import _wrap

        out = _emulations[target_name](*args, **attrs)
        specs = result_shape_dtypes if isinstance(result_shape_dtypes, (tuple, list)) else (result_shape_dtypes,)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        if len(outs) != len(specs):
            raise ValueError(f"{target_name}: {len(outs)} results for {len(specs)} declared")
        import numpy as _np

        outs = [_np.zeros(sp.shape, sp.dtype) if o is None else o for o, sp in zip(outs, specs)]  # None: scratch buffer
        for o, sp in zip(outs, specs):
            if tuple(o.shape) != tuple(sp.shape) or o.dtype != sp.dtype:
                raise ValueError(f"{target_name}: result {o.shape} {o.dtype} does not match the declared {sp.shape} {sp.dtype}")
        outs = [_wrap(o) for o in outs]
        return tuple(outs) if isinstance(result_shape_dtypes, (tuple, list)) else outs[0]

    return call
