"""Import the reference's own `cagpjax` package under the dependency stand-ins of oracle/refshim.  TEST INFRASTRUCTURE.

Use in a dedicated process (it puts stand-ins for jax / cola / gpjax / ... at the front of sys.path):

    from oracle.run_reference import load_reference
    ref = load_reference()            # dict: cagpjax, gpjax, jnp, ...
"""
import importlib
import os
import sys

REFERENCE_SRC = os.environ.get("CAGP_REFERENCE_SRC", "/root/reference/src")
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_STANDINS = ("jax", "jaxtyping", "equinox", "paramax", "cola", "gpjax", "lineax", "numpyro", "matfree")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "cagpjax"))


def load_reference():
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for name in _STANDINS + ("cagpjax",):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith((SHIM_DIR, REFERENCE_SRC)):
            raise RuntimeError(f"{name} is already imported from elsewhere; run the reference in a fresh process")
    sys.path[:0] = [SHIM_DIR, REFERENCE_SRC]
    mods = {name: importlib.import_module(name) for name in _STANDINS + ("cagpjax",)}
    mods["jnp"] = importlib.import_module("jax.numpy")
    return mods
