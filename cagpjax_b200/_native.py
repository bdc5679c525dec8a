"""ctypes binding of libcagp_b200.so (the C ABI declared in include/cagp_b200.h).

This is the test-time / torch-side binding of the drop-in boundary; a JAX host would bind the
same symbols as XLA FFI custom calls.  There is NO fallback: if the shared
library is missing or a tensor is not on a CUDA device the call raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libcagp_b200.so")

FAMILIES = {"rbf": 0, "matern12": 1, "matern32": 2, "matern52": 3}


def family_scale(family: str) -> float:
    """Coordinate scale c_family of ``cagp_scale_inputs`` (xs = (X - origin) * c_family / lengthscale)."""
    import math

    log2e = 1.0 / math.log(2.0)
    return {"rbf": math.sqrt(0.5 * log2e), "matern12": log2e, "matern32": math.sqrt(3.0) * log2e,
            "matern52": math.sqrt(5.0) * log2e}[family]
F32, F64 = 0, 1


class NativeLibraryError(RuntimeError):
    """Raised when libcagp_b200.so is missing or a native call fails."""


class KernelDesc(Structure):
    _fields_ = [
        ("family", c_int32),
        ("dtype", c_int32),
        ("n_lengthscale", c_int32),
        ("reserved", c_int32),
        ("lengthscale", c_void_p),
        ("bbox", c_void_p),
    ]


_SIGNATURES = {
    "cagp_last_error_string": (c_char_p, []),
    "cagp_version": (c_int, []),
    "cagp_padded_dim": (c_int, [c_int32]),
    "cagp_bounding_box": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p]),
    "cagp_scale_inputs": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "cagp_ks_blocksparse_fwd": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                        c_int32, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "cagp_ks_blocksparse_fwd_sym": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32,
                                            c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "cagp_ks_blocksparse_fwd_sym_peer": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32,
                                                 c_int32, c_int32, POINTER(c_void_p), POINTER(c_int64), c_int32, c_int32,
                                                 c_int64, c_void_p]),
    "cagp_ks_sym_fast_path": (c_int, [POINTER(KernelDesc), c_int64, c_int32, c_int32, c_int32]),
    "cagp_ks_fwd_sym_zero_rows": (c_int, [POINTER(KernelDesc), c_int64, c_int32, c_int32]),
    "cagp_ks_bwd_sym_ws_bytes": (c_size_t, [POINTER(KernelDesc), c_int64, c_int32, c_int32, c_int32, c_int32]),
    "cagp_ks_blocksparse_bwd_sym": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32,
                                            c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                            c_void_p]),
    "cagp_ks_bwd_ws_bytes": (c_size_t, [POINTER(KernelDesc), c_int64, c_int64, c_int32, c_int32]),
    "cagp_ks_blocksparse_bwd": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                        c_int32, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "cagp_project_rows": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "cagp_gram_ws_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "cagp_gram_mtm": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cagp_assemble_cotangent": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "cagp_assemble_cotangent_peer": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_int32),
                                             c_int32, c_int32, c_int64, c_void_p]),
    "cagp_project_rows_bwd": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "cagp_ks_dense_fwd": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                  c_void_p, c_int32, c_int64, c_void_p, c_int64, c_void_p]),
    "cagp_ks_dense_bwd_ls_ws_bytes": (c_size_t, [POINTER(KernelDesc), c_int64, c_int64, c_int32]),
    "cagp_ks_dense_bwd_ls": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                     c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cagp_cross_cov_fwd": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                   c_void_p, c_int64, c_void_p]),
    "cagp_cross_cov_bwd_ws_bytes": (c_size_t, [POINTER(KernelDesc), c_int64, c_int64, c_int32]),
    "cagp_cross_cov_bwd": (c_int, [POINTER(KernelDesc), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                   c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cagp_predict_ws_bytes": (c_size_t, [c_int32, c_int64, c_int32]),
    "cagp_predict_moments": (c_int, [c_int32, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None

# Launch accounting: number of kernels of THIS library launched (cuBLAS kernels not counted), and
# optional CUDA-event timing of the pair kernels on the launching stream (bench.py sets PROFILE = []).
LAUNCHES = 0
CALLS = {}  # C entry point -> number of calls (tests use it to prove which native path ran)
PROFILE = None
_KERNELS_PER_CALL = {"cagp_ks_blocksparse_fwd_sym_peer": 1, "cagp_bounding_box": 1, "cagp_ks_dense_fwd": 1, "cagp_ks_dense_bwd_ls": 2, "cagp_ks_blocksparse_fwd_sym": 1, "cagp_ks_blocksparse_bwd_sym": 2, "cagp_scale_inputs": 1, "cagp_ks_blocksparse_fwd": 1, "cagp_ks_blocksparse_bwd": 2,
                     "cagp_project_rows": 1, "cagp_gram_mtm": 2, "cagp_assemble_cotangent": 1, "cagp_assemble_cotangent_peer": 1,
                     "cagp_project_rows_bwd": 1, "cagp_predict_moments": 2, "cagp_cross_cov_fwd": 1, "cagp_cross_cov_bwd": 2}


class _timed:
    """Counts launches and, when PROFILE is a list, brackets the call with CUDA events."""

    def __init__(self, name, tensor, work=0):
        self.name, self.tensor, self.work = name, tensor, work

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += _KERNELS_PER_CALL.get(self.name, 0 if self.name.startswith("host:") else 1)
        CALLS[self.name] = CALLS.get(self.name, 0) + 1
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.tensor.device))
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream(self.tensor.device))
            PROFILE.append((self.name, self.e0, self.e1, self.work))
        return False


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load (once) and return the ctypes handle; raises NativeLibraryError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeLibraryError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cagpjax_b200/csrc`). There is no CPU or eager fallback for this path."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().cagp_last_error_string().decode()
        raise NativeLibraryError(f"{what} failed (status {rc}): {msg}")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float64:
        return F64
    if dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {dtype}; the native kernels are float32/float64 only")


def _ptr(t: torch.Tensor):
    return c_void_p(t.data_ptr())


def _require_cuda(*tensors: torch.Tensor):
    """Every tensor handed to one C call must be a contiguous CUDA tensor of ONE dtype on ONE device: the kernel
    instantiation is picked from the first tensor's dtype and the C ABI sees plain pointers, so a float64 buffer
    read as float32 (or the reverse, an out-of-bounds read) would be silent.  Callers cast first (see _fused)."""
    first = None
    for t in tensors:
        if not t.is_cuda:
            raise NativeLibraryError(
                "the CaGP hot path runs only on CUDA tensors (sm_100a kernels); got a tensor on "
                f"{t.device}. There is no CPU fallback."
            )
        if not t.is_contiguous():
            raise ValueError("native kernels need contiguous tensors")
        if first is None:
            first = t
        elif t.dtype != first.dtype or t.device != first.device:
            raise NativeLibraryError(
                f"native kernels need all operands in one dtype on one device; got {first.dtype} on {first.device} "
                f"and {t.dtype} on {t.device}")


def _stream(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def make_desc(family: str, lengthscale: torch.Tensor, dtype: torch.dtype, bbox: torch.Tensor = None) -> KernelDesc:
    ls = lengthscale.reshape(-1)
    if ls.dtype != dtype or not ls.is_cuda or (bbox is not None and (bbox.dtype != dtype or not bbox.is_cuda)):
        raise NativeLibraryError(f"kernel descriptor: lengthscale / bbox must be CUDA tensors of dtype {dtype}")
    return KernelDesc(FAMILIES[family], dtype_code(dtype), ls.numel(), 0, ls.data_ptr(),
                      bbox.data_ptr() if bbox is not None else None)


def bounding_box(xs: torch.Tensor) -> torch.Tensor:
    """{min, max} per (padded) coordinate of scaled inputs xs (dpad, n) -> (2 * dpad,)."""
    _require_cuda(xs)
    dp, n = xs.shape
    bbox = torch.empty(2 * dp, dtype=xs.dtype, device=xs.device)
    with torch.cuda.device(xs.device), _timed("cagp_bounding_box", xs, 0):
        _check(load().cagp_bounding_box(dtype_code(xs.dtype), _ptr(xs), n, n, dp, _ptr(bbox), _stream(xs)),
               "cagp_bounding_box")
    return bbox


def padded_dim(d: int) -> int:
    v = load().cagp_padded_dim(d)
    if v < 0:
        raise NativeLibraryError(f"input dimension {d} > 32 is not supported by the native kernels")
    return v


def scale_inputs(family: str, X: torch.Tensor, lengthscale: torch.Tensor, origin: torch.Tensor = None) -> torch.Tensor:
    """xs (padded_dim(d), n): inputs shifted by ``origin`` (d,), scaled and transposed."""
    n, d = X.shape
    ls = lengthscale.reshape(-1).to(device=X.device, dtype=X.dtype).contiguous()
    if origin is not None:
        origin = origin.detach().reshape(-1).to(device=X.device, dtype=X.dtype).contiguous()
        _require_cuda(X, ls, origin)
    else:
        _require_cuda(X, ls)
    xs = torch.empty((padded_dim(d), n), dtype=X.dtype, device=X.device)
    kd = make_desc(family, ls, X.dtype)
    with torch.cuda.device(X.device), _timed("cagp_scale_inputs", X, 0):
        _check(load().cagp_scale_inputs(byref(kd), _ptr(X), n, d, _ptr(origin) if origin is not None else None, _ptr(xs),
                                        n, _stream(X)), "cagp_scale_inputs")
    return xs


def ks_blocksparse_fwd(family, xs1, xs2, d, s, k, lengthscale, bbox=None) -> torch.Tensor:
    """Mt (k, m) = (K'(X1, X2) S)^T.  ``bbox``: bounding box of BOTH scaled point sets (optional)."""
    _require_cuda(xs1, xs2, s)
    m, n = xs1.shape[1], xs2.shape[1]
    Mt = torch.empty((k, m), dtype=xs1.dtype, device=xs1.device)
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype, bbox)
    with torch.cuda.device(xs1.device), _timed("cagp_ks_blocksparse_fwd", xs1, m * n):
        _check(load().cagp_ks_blocksparse_fwd(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(s), k, _ptr(Mt), m,
                                              _stream(xs1)), "cagp_ks_blocksparse_fwd")
    return Mt


def ks_blocksparse_bwd(family, xs1, xs2, d, s, k, Gt, lengthscale):
    """(s_bar (n), ls_bar (n_lengthscale)) for cotangent Gt (k, m)."""
    _require_cuda(xs1, xs2, s, Gt)
    m, n = xs1.shape[1], xs2.shape[1]
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype)
    lib = load()
    ws_bytes = lib.cagp_ks_bwd_ws_bytes(byref(kd), m, n, d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs1.device)
    s_bar = torch.empty(n, dtype=xs1.dtype, device=xs1.device)
    ls_bar = torch.empty(ls.numel(), dtype=xs1.dtype, device=xs1.device)
    with torch.cuda.device(xs1.device), _timed("cagp_ks_blocksparse_bwd", xs1, m * n):
        _check(lib.cagp_ks_blocksparse_bwd(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(s), k, _ptr(Gt),
                                           Gt.shape[1], _ptr(s_bar), _ptr(ls_bar), _ptr(ws), ws_bytes, _stream(xs1)),
               "cagp_ks_blocksparse_bwd")
    return s_bar, ls_bar


def ks_blocksparse_fwd_sym(family, xs, d, s, k, lengthscale, a_begin=0, a_end=None, out=None, bbox=None) -> torch.Tensor:
    """Symmetric Gram projection: Mt (k, n) entries of the tiles (a, a+off) for a in [a_begin, a_end)."""
    _require_cuda(xs, s)
    n = xs.shape[1]
    a_end = k if a_end is None else a_end
    full = a_begin == 0 and a_end == k
    Mt = out if out is not None else (torch.empty if full else torch.zeros)((k, n), dtype=xs.dtype, device=xs.device)
    ls = lengthscale.reshape(-1).to(device=xs.device, dtype=xs.dtype).contiguous()
    kd = make_desc(family, ls, xs.dtype, bbox)
    with torch.cuda.device(xs.device), _timed("cagp_ks_blocksparse_fwd_sym", xs, n * n * (a_end - a_begin) // k):
        _check(load().cagp_ks_blocksparse_fwd_sym(byref(kd), _ptr(xs), n, n, d, _ptr(s), k, a_begin, a_end, _ptr(Mt), n,
                                                  _stream(xs)), "cagp_ks_blocksparse_fwd_sym")
    return Mt


def ks_blocksparse_fwd_sym_peer(family, xs, d, s, k, lengthscale, a_begin, a_end, panel_ptrs, row_begin, me, ldm,
                                bbox=None) -> None:
    """Peer-memory symmetric forward: row sums into this rank's panel, column sums straight into the owners' panels."""
    _require_cuda(xs, s)
    n = xs.shape[1]
    world = len(panel_ptrs)
    ls = lengthscale.reshape(-1).to(device=xs.device, dtype=xs.dtype).contiguous()
    kd = make_desc(family, ls, xs.dtype, bbox)
    ptrs = (c_void_p * world)(*[c_void_p(int(p)) for p in panel_ptrs])
    rb = (c_int64 * (world + 1))(*[int(v) for v in row_begin])
    with torch.cuda.device(xs.device), _timed("cagp_ks_blocksparse_fwd_sym_peer", xs, n * n * (a_end - a_begin) // k):
        _check(load().cagp_ks_blocksparse_fwd_sym_peer(byref(kd), _ptr(xs), n, n, d, _ptr(s), k, a_begin, a_end, ptrs, rb,
                                                       world, me, ldm, _stream(xs)), "cagp_ks_blocksparse_fwd_sym_peer")


def sym_zero_rows(family, xs, d, k, lengthscale, bbox) -> int:
    """0 / 1 / 2: nothing / row k-1 / the whole panel must be zeroed before the peer-memory symmetric forward."""
    ls = lengthscale.reshape(-1).to(device=xs.device, dtype=xs.dtype).contiguous()
    kd = make_desc(family, ls, xs.dtype, bbox)
    return int(load().cagp_ks_fwd_sym_zero_rows(byref(kd), xs.shape[1], d, k))


def sym_fast_path(family, xs, d, k, lengthscale, bbox, backward=False) -> int:
    """Non-zero if the symmetric call also launches the float64 RBF fast-path kernel (csrc/ks_symx.cuh) AND the
    bounding box proves its expanded form valid, i.e. the fast-path kernel does the work: 1 = every scaled
    |x|^2 < 300 (no clamp), 2 = every |x|^2 < 900 (exponent clamp); 0 = the general kernel runs
    (kernel_math.cuh, launch_expand_level)."""
    if bbox is None:
        return 0
    ls = lengthscale.reshape(-1).to(device=xs.device, dtype=xs.dtype).contiguous()
    kd = make_desc(family, ls, xs.dtype, bbox)
    if not load().cagp_ks_sym_fast_path(byref(kd), xs.shape[1], d, k, 1 if backward else 0):
        return 0
    dp = xs.shape[0]
    m = torch.maximum(bbox[:dp].abs(), bbox[dp:].abs())
    r2 = float((m * m).sum())
    return 1 if r2 < 300.0 else (2 if r2 < 900.0 else 0)


def ks_blocksparse_bwd_sym(family, xs, d, s, k, Gt, lengthscale, a_begin=0, a_end=None, bbox=None):
    """(s_bar (n), ls_bar) contributions of the tiles of blocks [a_begin, a_end), full cotangent Gt (k, n)."""
    _require_cuda(xs, s, Gt)
    n = xs.shape[1]
    a_end = k if a_end is None else a_end
    ls = lengthscale.reshape(-1).to(device=xs.device, dtype=xs.dtype).contiguous()
    kd = make_desc(family, ls, xs.dtype, bbox)
    lib = load()
    ws_bytes = lib.cagp_ks_bwd_sym_ws_bytes(byref(kd), n, d, k, a_begin, a_end)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs.device)
    s_bar = torch.empty(n, dtype=xs.dtype, device=xs.device)
    ls_bar = torch.empty(ls.numel(), dtype=xs.dtype, device=xs.device)
    with torch.cuda.device(xs.device), _timed("cagp_ks_blocksparse_bwd_sym", xs, n * n * (a_end - a_begin) // k):
        _check(lib.cagp_ks_blocksparse_bwd_sym(byref(kd), _ptr(xs), n, n, d, _ptr(s), k, a_begin, a_end, _ptr(Gt),
                                               Gt.shape[1], _ptr(s_bar), _ptr(ls_bar), _ptr(ws), ws_bytes, _stream(xs)),
               "cagp_ks_blocksparse_bwd_sym")
    return s_bar, ls_bar


def ks_dense_fwd(family, xs1, xs2, d, V, lengthscale) -> torch.Tensor:
    """Y (m, p) = K'(X1, X2) V for dense V (n, p)."""
    _require_cuda(xs1, xs2, V)
    m, n, p = xs1.shape[1], xs2.shape[1], V.shape[1]
    Y = torch.empty((m, p), dtype=xs1.dtype, device=xs1.device)
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype)
    with torch.cuda.device(xs1.device), _timed("cagp_ks_dense_fwd", xs1, m * n):
        _check(load().cagp_ks_dense_fwd(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(V), p, p, _ptr(Y), p,
                                        _stream(xs1)), "cagp_ks_dense_fwd")
    return Y


def ks_dense_bwd_ls(family, xs1, xs2, d, V, Ybar, lengthscale) -> torch.Tensor:
    """ls_bar (n_lengthscale) = sum_ij (Ybar_i . V_j) dK'_ij / d lengthscale."""
    _require_cuda(xs1, xs2, V, Ybar)
    m, n, p = xs1.shape[1], xs2.shape[1], V.shape[1]
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype)
    lib = load()
    ws_bytes = lib.cagp_ks_dense_bwd_ls_ws_bytes(byref(kd), m, n, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs1.device)
    ls_bar = torch.empty(ls.numel(), dtype=xs1.dtype, device=xs1.device)
    Vt, Ybt = V.T.contiguous(), Ybar.T.contiguous()  # point index contiguous for coalesced tile loads
    with torch.cuda.device(xs1.device), _timed("cagp_ks_dense_bwd_ls", xs1, m * n):
        _check(lib.cagp_ks_dense_bwd_ls(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(Vt), n, _ptr(Ybt), m, p,
                                        _ptr(ls_bar), _ptr(ws), ws_bytes, _stream(xs1)), "cagp_ks_dense_bwd_ls")
    return ls_bar


def cross_cov_fwd(family, xs1, xs2, d, lengthscale) -> torch.Tensor:
    """C (m, n) = K'(X1, X2), materialised (unit variance)."""
    _require_cuda(xs1, xs2)
    m, n = xs1.shape[1], xs2.shape[1]
    C = torch.empty((m, n), dtype=xs1.dtype, device=xs1.device)
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype)
    with torch.cuda.device(xs1.device), _timed("cagp_cross_cov_fwd", xs1, m * n):
        _check(load().cagp_cross_cov_fwd(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(C), n, _stream(xs1)),
               "cagp_cross_cov_fwd")
    return C


def cross_cov_bwd(family, xs1, xs2, d, Cbar, lengthscale):
    """(x2s_bar (padded_dim(d), n), ls_bar (n_lengthscale)): cotangents of the SCALED x2 coordinates and the lengthscale."""
    Cbar = Cbar.contiguous()
    _require_cuda(xs1, xs2, Cbar)
    m, n = xs1.shape[1], xs2.shape[1]
    ls = lengthscale.reshape(-1).to(device=xs1.device, dtype=xs1.dtype).contiguous()
    kd = make_desc(family, ls, xs1.dtype)
    lib = load()
    ws_bytes = lib.cagp_cross_cov_bwd_ws_bytes(byref(kd), m, n, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs1.device)
    x2s_bar = torch.empty((xs2.shape[0], n), dtype=xs1.dtype, device=xs1.device)
    ls_bar = torch.empty(ls.numel(), dtype=xs1.dtype, device=xs1.device)
    with torch.cuda.device(xs1.device), _timed("cagp_cross_cov_bwd", xs1, m * n):
        _check(lib.cagp_cross_cov_bwd(byref(kd), _ptr(xs1), m, m, _ptr(xs2), n, n, d, _ptr(Cbar), n, _ptr(x2s_bar),
                                      _ptr(ls_bar), _ptr(ws), ws_bytes, _stream(xs1)), "cagp_cross_cov_bwd")
    return x2s_bar, ls_bar


def union_bounding_box(xs1: torch.Tensor, xs2: torch.Tensor) -> torch.Tensor:
    """Bounding box of two scaled point sets (same padded dimension)."""
    b1 = bounding_box(xs1)
    if xs1.data_ptr() == xs2.data_ptr() and xs1.shape == xs2.shape:
        return b1
    b2 = bounding_box(xs2)
    dp = xs1.shape[0]
    return torch.cat([torch.minimum(b1[:dp], b2[:dp]), torch.maximum(b1[dp:], b2[dp:])]).contiguous()


def project_rows(Mt, row_offset, n_total, s_local, yt_local):
    """A' (k, k) partial over the local rows and c' (k)."""
    _require_cuda(Mt, s_local, yt_local)
    k, m = Mt.shape
    A = torch.empty((k, k), dtype=Mt.dtype, device=Mt.device)
    c = torch.empty(k, dtype=Mt.dtype, device=Mt.device)
    with torch.cuda.device(Mt.device), _timed("cagp_project_rows", Mt, 0):
        _check(load().cagp_project_rows(dtype_code(Mt.dtype), _ptr(Mt), m, m, row_offset, n_total, k, _ptr(s_local),
                                        _ptr(yt_local), _ptr(A), _ptr(c), _stream(Mt)), "cagp_project_rows")
    return A, c


def gram_mtm(Mt):
    _require_cuda(Mt)
    k, m = Mt.shape
    B = torch.empty((k, k), dtype=Mt.dtype, device=Mt.device)
    lib = load()
    ws_bytes = lib.cagp_gram_ws_bytes(dtype_code(Mt.dtype), m, k)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=Mt.device)
    with torch.cuda.device(Mt.device), _timed("cagp_gram_mtm", Mt, m * k * k):
        _check(lib.cagp_gram_mtm(dtype_code(Mt.dtype), _ptr(Mt), m, m, k, _ptr(B), _ptr(ws), ws_bytes, _stream(Mt)),
               "cagp_gram_mtm")
    return B


def assemble_cotangent(Mt, row_offset, n_total, W, Abar, cbar, s_local, yt_local):
    _require_cuda(Mt, W, Abar, cbar, s_local, yt_local)
    k, m = Mt.shape
    Gt = torch.empty((k, m), dtype=Mt.dtype, device=Mt.device)
    with torch.cuda.device(Mt.device), _timed("cagp_assemble_cotangent", Mt, 2 * m * k * k):
        _check(load().cagp_assemble_cotangent(dtype_code(Mt.dtype), _ptr(Mt), m, m, row_offset, n_total, k, _ptr(W),
                                              _ptr(Abar), _ptr(cbar), _ptr(s_local), _ptr(yt_local), _ptr(Gt), m,
                                              _stream(Mt)), "cagp_assemble_cotangent")
    return Gt


def assemble_cotangent_peer(Mt, row_offset, n_total, W, Abar, cbar, s_local, yt_local, buffer_ptrs, block_begin, me, ld):
    """Cotangent panel of this rank written straight into the ranks' full-width (k, ld) buffers (peer memory): its own
    buffer gets the whole panel, the owner of action block b additionally gets row b.  float64 only."""
    _require_cuda(Mt, W, Abar, cbar, s_local, yt_local)
    k, m = Mt.shape
    world = len(buffer_ptrs)
    ptrs = (c_void_p * world)(*[c_void_p(int(p)) for p in buffer_ptrs])
    bb = (c_int32 * (world + 1))(*[int(v) for v in block_begin])
    with torch.cuda.device(Mt.device), _timed("cagp_assemble_cotangent_peer", Mt, 2 * m * k * k):
        _check(load().cagp_assemble_cotangent_peer(dtype_code(Mt.dtype), _ptr(Mt), m, m, row_offset, n_total, k, _ptr(W),
                                                   _ptr(Abar), _ptr(cbar), _ptr(s_local), _ptr(yt_local), ptrs, bb, world,
                                                   me, ld, _stream(Mt)), "cagp_assemble_cotangent_peer")


def project_rows_bwd(Mt, row_offset, n_total, Abar, cbar):
    _require_cuda(Mt, Abar, cbar)
    k, m = Mt.shape
    s_dir = torch.empty(m, dtype=Mt.dtype, device=Mt.device)
    yt_bar = torch.empty(m, dtype=Mt.dtype, device=Mt.device)
    with torch.cuda.device(Mt.device), _timed("cagp_project_rows_bwd", Mt, 0):
        _check(load().cagp_project_rows_bwd(dtype_code(Mt.dtype), _ptr(Mt), m, m, row_offset, n_total, k, _ptr(Abar),
                                            _ptr(cbar), _ptr(s_dir), _ptr(yt_bar), _stream(Mt)),
               "cagp_project_rows_bwd")
    return s_dir, yt_bar


def predict_moments(Mt, w, Pinv, variance, mean_const):
    _require_cuda(Mt, w, Pinv)
    k, m = Mt.shape
    lib = load()
    scalars = torch.stack([variance.reshape(()).to(Mt.dtype), mean_const.reshape(()).to(Mt.dtype)]).contiguous()
    ws_bytes = lib.cagp_predict_ws_bytes(dtype_code(Mt.dtype), m, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=Mt.device)
    mean = torch.empty(m, dtype=Mt.dtype, device=Mt.device)
    var = torch.empty(m, dtype=Mt.dtype, device=Mt.device)
    with torch.cuda.device(Mt.device), _timed("cagp_predict_moments", Mt, 0):
        _check(lib.cagp_predict_moments(dtype_code(Mt.dtype), _ptr(Mt), m, m, k, _ptr(w), _ptr(Pinv), _ptr(scalars),
                                        _ptr(mean), _ptr(var), _ptr(ws), ws_bytes, _stream(Mt)),
               "cagp_predict_moments")
    return mean, var
