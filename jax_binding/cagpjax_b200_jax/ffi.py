"""FFI target registration and host-only helpers.  Never run on JAX (none in this image; API: jax.ffi of
jax >= 0.4.38); registration and host helpers are exercised by tests/test_jax_binding_wiring.py under numpy stand-ins."""
import ctypes
import os

import jax

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.environ.get("CAGP_B200_LIBDIR", os.path.join(_HERE, "..", "..", "cagpjax_b200", "_lib"))
core = ctypes.CDLL(os.path.join(_LIBDIR, "libcagp_b200.so"))       # the C ABI itself (host-only helpers)
shim = ctypes.CDLL(os.path.join(_LIBDIR, "libcagp_b200_xla.so"))   # csrc/xla_ffi_shim.cc

# one XLA FFI handler per compute entry point of include/cagp_b200.h (handler symbol -> C symbol it wraps)
TARGETS = {
    "CagpScaleInputs": "cagp_scale_inputs",
    "CagpBoundingBox": "cagp_bounding_box",
    "CagpKsBlocksparseFwd": "cagp_ks_blocksparse_fwd",
    "CagpKsBlocksparseBwd": "cagp_ks_blocksparse_bwd",
    "CagpKsBlocksparseFwdSym": "cagp_ks_blocksparse_fwd_sym",
    "CagpKsBlocksparseFwdSymPeer": "cagp_ks_blocksparse_fwd_sym_peer",
    "CagpKsBlocksparseBwdSym": "cagp_ks_blocksparse_bwd_sym",
    "CagpProjectRows": "cagp_project_rows",
    "CagpGramMtm": "cagp_gram_mtm",
    "CagpAssembleCotangent": "cagp_assemble_cotangent",
    "CagpAssembleCotangentPeer": "cagp_assemble_cotangent_peer",
    "CagpProjectRowsBwd": "cagp_project_rows_bwd",
    "CagpPredictMoments": "cagp_predict_moments",
    "CagpKsDenseFwd": "cagp_ks_dense_fwd",
    "CagpKsDenseBwdLs": "cagp_ks_dense_bwd_ls",
    "CagpCrossCovFwd": "cagp_cross_cov_fwd",
    "CagpCrossCovBwd": "cagp_cross_cov_bwd",
}
for _name in TARGETS:
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(shim, _name)), platform="CUDA")

FAMILIES = {"RBF": 0, "Matern12": 1, "Matern32": 2, "Matern52": 3}
F32, F64 = 0, 1


class KernelDesc(ctypes.Structure):  # cagp_kernel_desc (only shape-like fields matter for the *_ws_bytes helpers)
    _fields_ = [("family", ctypes.c_int32), ("dtype", ctypes.c_int32), ("n_lengthscale", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("lengthscale", ctypes.c_void_p), ("bbox", ctypes.c_void_p)]


core.cagp_padded_dim.restype = ctypes.c_int
core.cagp_padded_dim.argtypes = [ctypes.c_int32]
for _f, _args in {"cagp_ks_bwd_ws_bytes": [ctypes.POINTER(KernelDesc), ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32],
                  "cagp_ks_bwd_sym_ws_bytes": [ctypes.POINTER(KernelDesc), ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32],
                  "cagp_ks_dense_bwd_ls_ws_bytes": [ctypes.POINTER(KernelDesc), ctypes.c_int64, ctypes.c_int64, ctypes.c_int32],
                  "cagp_cross_cov_bwd_ws_bytes": [ctypes.POINTER(KernelDesc), ctypes.c_int64, ctypes.c_int64, ctypes.c_int32],
                  "cagp_gram_ws_bytes": [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32],
                  "cagp_predict_ws_bytes": [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32]}.items():
    getattr(core, _f).restype = ctypes.c_size_t
    getattr(core, _f).argtypes = _args


def dtype_code(dtype) -> int:
    import jax.numpy as jnp

    return F64 if dtype == jnp.float64 else F32


def desc(family: int, dtype, n_lengthscale: int) -> KernelDesc:
    """Shape-only descriptor for the workspace-size queries (pointers are not dereferenced there, but must be non-null)."""
    return KernelDesc(family, dtype_code(dtype), n_lengthscale, 0, 1, None)
