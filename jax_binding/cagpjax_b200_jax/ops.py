"""jax.custom_vjp wrappers over the FFI targets: forward and backward kernels are paired by hand, as
cagpjax_b200/_fused.py does with torch.autograd.Function.  Never run on JAX (none in this image); the forward
wrappers run under numpy stand-ins with emulated FFI targets (tests/test_jax_binding_wiring.py) and the reverse-mode rule
of `project_stats` is checked there against central differences (called directly: no autodiff in the stand-ins).

Layouts follow include/cagp_b200.h: xs = scaled SoA inputs (dpad, n); Mt = (K'S)^T stored k-major (k, m)."""
import functools
from ctypes import byref

import jax
import jax.numpy as jnp
import numpy as np

from . import ffi as F

_u8 = lambda nbytes: jax.ShapeDtypeStruct((max(int(nbytes), 8),), jnp.uint8)   # workspace result buffer
_empty = lambda dtype: jnp.zeros((0,), dtype)                                  # "no bounding box" operand


def scale_inputs(X, lengthscale, family: int, origin=None):
    n, d = X.shape
    dp = F.core.cagp_padded_dim(d)
    origin = _empty(X.dtype) if origin is None else origin.astype(X.dtype)
    return jax.ffi.ffi_call("CagpScaleInputs", jax.ShapeDtypeStruct((dp, n), X.dtype))(
        X, jnp.atleast_1d(lengthscale).astype(X.dtype), origin, family=np.int64(family))


def bounding_box(xs):
    return jax.ffi.ffi_call("CagpBoundingBox", jax.ShapeDtypeStruct((2 * xs.shape[0],), xs.dtype))(xs)


# ---- rectangular projection M' = K'(X1, X2) S --------------------------------------------------------------------
@functools.partial(jax.custom_vjp, nondiff_argnums=(4, 5, 6))
def ks_projection(xs1, xs2, s, lengthscale, family, d, k):
    m = xs1.shape[1]
    return jax.ffi.ffi_call("CagpKsBlocksparseFwd", jax.ShapeDtypeStruct((k, m), xs1.dtype))(
        xs1, xs2, s, lengthscale, _empty(xs1.dtype), family=np.int64(family), d=np.int64(d), k=np.int64(k))


def _ks_fwd(xs1, xs2, s, lengthscale, family, d, k):
    return ks_projection(xs1, xs2, s, lengthscale, family, d, k), (xs1, xs2, s, lengthscale)


def _ks_bwd(family, d, k, res, Gt):
    xs1, xs2, s, ls = res
    kd = F.desc(family, xs1.dtype, ls.size)
    ws = F.core.cagp_ks_bwd_ws_bytes(byref(kd), xs1.shape[1], xs2.shape[1], d, k)
    s_bar, ls_bar, _ = jax.ffi.ffi_call(
        "CagpKsBlocksparseBwd", (jax.ShapeDtypeStruct(s.shape, s.dtype), jax.ShapeDtypeStruct(ls.shape, ls.dtype), _u8(ws)))(
        xs1, xs2, s, ls, Gt, family=np.int64(family), d=np.int64(d), k=np.int64(k))
    return None, None, s_bar, ls_bar  # inputs are data (non-trainable) in the reference path


ks_projection.defvjp(_ks_fwd, _ks_bwd)


# ---- symmetric (Gram) projection: every kernel value once per unordered block pair ------------------------------------
@functools.partial(jax.custom_vjp, nondiff_argnums=(4, 5, 6))
def ks_projection_sym(xs, s, lengthscale, bbox, family, d, k):
    n = xs.shape[1]
    return jax.ffi.ffi_call("CagpKsBlocksparseFwdSym", jax.ShapeDtypeStruct((k, n), xs.dtype))(
        xs, s, lengthscale, bbox, family=np.int64(family), d=np.int64(d), k=np.int64(k), a_begin=np.int64(0), a_end=np.int64(k))


def _sym_fwd(xs, s, lengthscale, bbox, family, d, k):
    return ks_projection_sym(xs, s, lengthscale, bbox, family, d, k), (xs, s, lengthscale, bbox)


def _sym_bwd(family, d, k, res, Gt):
    xs, s, ls, bbox = res
    kd = F.desc(family, xs.dtype, ls.size)
    ws = F.core.cagp_ks_bwd_sym_ws_bytes(byref(kd), xs.shape[1], d, k, 0, k)
    s_bar, ls_bar, _ = jax.ffi.ffi_call(
        "CagpKsBlocksparseBwdSym", (jax.ShapeDtypeStruct(s.shape, s.dtype), jax.ShapeDtypeStruct(ls.shape, ls.dtype), _u8(ws)))(
        xs, s, ls, bbox, Gt, family=np.int64(family), d=np.int64(d), k=np.int64(k), a_begin=np.int64(0), a_end=np.int64(k))
    return None, s_bar, ls_bar, None


ks_projection_sym.defvjp(_sym_fwd, _sym_bwd)


# ---- k-sized statistics of the fused pass and their reverse mode ----------------------------------------
@functools.partial(jax.custom_vjp, nondiff_argnums=(4, 5))
def project_stats(X, lengthscale, s, yt, family, k):
    """(A', B', c') = (S^T K' S, (K'S)^T (K'S), (K'S)^T yt) for the Gram operator of X."""
    return _project_stats_fwd(X, lengthscale, s, yt, family, k)[0]


def _project_stats_fwd(X, lengthscale, s, yt, family, k):
    n, d = X.shape
    origin = 0.5 * (X.min(axis=0) + X.max(axis=0))
    xs = scale_inputs(X, lengthscale, family, origin)
    bbox = bounding_box(xs)
    ls = jnp.atleast_1d(lengthscale).astype(X.dtype)
    Mt = ks_projection_sym(xs, s, ls, bbox, family, d, k)
    A, c = jax.ffi.ffi_call("CagpProjectRows", (jax.ShapeDtypeStruct((k, k), X.dtype), jax.ShapeDtypeStruct((k,), X.dtype)))(
        Mt, s, yt, row_offset=np.int64(0), n_total=np.int64(n))
    B, _ = jax.ffi.ffi_call("CagpGramMtm", (jax.ShapeDtypeStruct((k, k), X.dtype),
                                            _u8(F.core.cagp_gram_ws_bytes(F.dtype_code(X.dtype), n, k))))(Mt)
    return (A, B, c), (xs, bbox, ls, s, yt, Mt, lengthscale)


def _project_stats_bwd(family, k, res, cot):
    xs, bbox, ls, s, yt, Mt, lengthscale = res
    Abar, Bbar, cbar = cot
    n, d = xs.shape[1], xs.shape[0]
    Gt = jax.ffi.ffi_call("CagpAssembleCotangent", jax.ShapeDtypeStruct(Mt.shape, Mt.dtype))(
        Mt, Bbar + Bbar.T, Abar, cbar, s, yt, row_offset=np.int64(0), n_total=np.int64(n))
    kd = F.desc(family, xs.dtype, ls.size)
    ws = F.core.cagp_ks_bwd_sym_ws_bytes(byref(kd), n, d, k, 0, k)
    s_bar, ls_bar, _ = jax.ffi.ffi_call(
        "CagpKsBlocksparseBwdSym", (jax.ShapeDtypeStruct(s.shape, s.dtype), jax.ShapeDtypeStruct(ls.shape, ls.dtype), _u8(ws)))(
        xs, s, ls, bbox, Gt, family=np.int64(family), d=np.int64(d), k=np.int64(k), a_begin=np.int64(0), a_end=np.int64(k))
    s_dir, yt_bar = jax.ffi.ffi_call("CagpProjectRowsBwd", (jax.ShapeDtypeStruct(s.shape, s.dtype), jax.ShapeDtypeStruct(yt.shape, yt.dtype)))(
        Mt, Abar, cbar, row_offset=np.int64(0), n_total=np.int64(n))
    return None, ls_bar.reshape(jnp.shape(lengthscale)), s_bar + s_dir, yt_bar


project_stats.defvjp(_project_stats_fwd, _project_stats_bwd)


# ---- dense right-hand side: LazyKernel._matmat (operators/lazy_kernel.py:79-90) ---------------------------------------
@functools.partial(jax.custom_vjp, nondiff_argnums=(4, 5))
def kernel_matmat(xs1, xs2, V, lengthscale, family, d):
    m, p = xs1.shape[1], V.shape[1]
    return jax.ffi.ffi_call("CagpKsDenseFwd", jax.ShapeDtypeStruct((m, p), V.dtype))(
        xs1, xs2, V, lengthscale, family=np.int64(family), d=np.int64(d))


def _mm_fwd(xs1, xs2, V, lengthscale, family, d):
    return kernel_matmat(xs1, xs2, V, lengthscale, family, d), (xs1, xs2, V, lengthscale)


def _mm_bwd(family, d, res, Ybar):
    xs1, xs2, V, ls = res
    V_bar = kernel_matmat(xs2, xs1, Ybar, ls, family, d)  # K'^T Ybar: K'(x, y) = K'(y, x)
    kd = F.desc(family, xs1.dtype, ls.size)
    ws = F.core.cagp_ks_dense_bwd_ls_ws_bytes(byref(kd), xs1.shape[1], xs2.shape[1], d)
    ls_bar, _ = jax.ffi.ffi_call("CagpKsDenseBwdLs", (jax.ShapeDtypeStruct(ls.shape, ls.dtype), _u8(ws)))(
        xs1, xs2, V.T, Ybar.T, ls, family=np.int64(family), d=np.int64(d))
    return None, None, V_bar, ls_bar


kernel_matmat.defvjp(_mm_fwd, _mm_bwd)


def predict_moments(Mt, w, Pinv, variance, mean_const):
    k, m = Mt.shape
    ws = F.core.cagp_predict_ws_bytes(F.dtype_code(Mt.dtype), m, k)
    mean, var, _ = jax.ffi.ffi_call("CagpPredictMoments", (jax.ShapeDtypeStruct((m,), Mt.dtype), jax.ShapeDtypeStruct((m,), Mt.dtype), _u8(ws)))(
        Mt, w, Pinv, jnp.stack([variance, mean_const]).astype(Mt.dtype))
    return mean, var
