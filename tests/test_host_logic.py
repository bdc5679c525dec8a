"""CPU tests of the host-side logic: API surface, error behaviour, operator algebra, the C-ABI library
(loads, exports every declared symbol) and the absence of any CPU fallback on the kernel path."""

import os
import re
import sys

import pytest
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import _fused, _native, gpx, ops
from cagpjax_b200.linalg import congruence_transform, lower_cholesky
from cagpjax_b200.linalg.utils import _add_jitter
from cagpjax_b200.operators import BlockDiagonalSparse, LazyKernel, diag_like
from cagpjax_b200.operators.annotations import ScaledOrthogonal
from cagpjax_b200.policies import BlockSparsePolicy
from cagpjax_b200.solvers import Cholesky
from oracle import cagp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = torch.float64


def _rand(*shape, seed=0, dtype=DT):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=dtype)


# ---- C ABI ---------------------------------------------------------------------------------------
def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cagp_b200.h")).read()
    declared = set(re.findall(r"\b(cagp_[a-z0-9_]+)\s*\(", header))
    declared -= {"cagp_stream_t"}
    assert declared, "no declarations parsed"
    lib = _native.load()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/cagp_b200.h but not exported"
    assert declared == set(_native.exported_symbols()), "ctypes signatures out of sync with the header"
    assert lib.cagp_version() >= 100
    assert [lib.cagp_padded_dim(d) for d in (1, 3, 5, 7, 16, 17, 32, 33)] == [1, 3, 8, 8, 16, 32, 32, -1]


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary must be bindable from C / cgo / JNI / XLA FFI: the header compiles as C99 and as C++."""
    import shutil
    import subprocess

    header = os.path.join(ROOT, "include", "cagp_b200.h")
    src_c = tmp_path / "probe.c"
    src_c.write_text('#include "cagp_b200.h"\nint main(void) { cagp_kernel_desc kd; (void)kd; return (int)sizeof(cagp_status) * 0; }\n'
                     .replace("sizeof(cagp_status)", "sizeof(enum cagp_status)"))
    for cc, flags in (("gcc", ["-std=c99", "-pedantic", "-Wall", "-Werror"]), ("g++", ["-std=c++17", "-Wall", "-Werror", "-x", "c++"])):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not available")
        subprocess.run([cc, *flags, "-fsyntax-only", "-I", os.path.dirname(header), str(src_c)], check=True)


# host-only helpers of the C ABI (no launch): called through ctypes from either host language, no FFI handler needed
_HOST_ONLY = {"cagp_last_error_string", "cagp_version", "cagp_padded_dim", "cagp_ks_sym_fast_path", "cagp_ks_fwd_sym_zero_rows", "cagp_ks_bwd_ws_bytes",
              "cagp_ks_bwd_sym_ws_bytes", "cagp_ks_dense_bwd_ls_ws_bytes", "cagp_cross_cov_bwd_ws_bytes",
              "cagp_gram_ws_bytes", "cagp_predict_ws_bytes"}


def test_xla_ffi_shim_is_guarded_and_covers_every_compute_symbol(tmp_path):
    """csrc/xla_ffi_shim.cc (the JAX side of the boundary): without the XLA headers it must compile to an
    empty translation unit; every compute entry point of the header must be called by exactly one FFI handler, and
    jax_binding/ must register that handler."""
    import shutil
    import subprocess

    shim = os.path.join(ROOT, "cagpjax_b200", "csrc", "xla_ffi_shim.cc")
    src = open(shim).read()
    if shutil.which("g++"):
        obj = tmp_path / "shim.o"
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-c", shim, "-o", str(obj)], check=True)
        syms = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
        assert "Cagp" not in syms, "the shim must be empty when xla/ffi/api/ffi.h is absent"
    header = open(os.path.join(ROOT, "include", "cagp_b200.h")).read()
    declared = set(re.findall(r"\b(cagp_[a-z0-9_]+)\s*\(", header)) - {"cagp_stream_t"}
    compute = declared - _HOST_ONLY
    called = set(re.findall(r"\b(cagp_[a-z0-9_]+)\s*\(", src)) - {"cagp_last_error_string"}
    assert compute == called, (sorted(compute - called), sorted(called - compute))
    handlers = re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((Cagp\w+),", src)
    assert len(handlers) == len(set(handlers)) == len(compute)
    ffi_py = open(os.path.join(ROOT, "jax_binding", "cagpjax_b200_jax", "ffi.py")).read()
    targets = dict(re.findall(r'"(Cagp\w+)":\s*"(cagp_\w+)"', ffi_py))
    assert set(targets) == set(handlers) and set(targets.values()) == compute


def test_jax_binding_calls_agree_with_the_shim_bindings():
    """Every `jax.ffi.ffi_call(name, results)(operands..., attributes...)` in jax_binding/.../ops.py must pass as many
    operands, declare as many results and name the same attributes as the handler's binding in csrc/xla_ffi_shim.cc."""
    import ast

    shim = open(os.path.join(ROOT, "cagpjax_b200", "csrc", "xla_ffi_shim.cc")).read()
    bindings = {}
    for m in re.finditer(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((Cagp\w+),\s*\w+,\s*(.*?)\);", shim, flags=re.S):
        body = m.group(2)
        attrs = re.findall(r'(?:I64|\.Attr<[^>]*(?:<[^>]*>)?[^>]*>)\("(\w+)"\)', body)
        bindings[m.group(1)] = (len(re.findall(r"\bARG\b", body)), len(re.findall(r"\bRET\b", body)), attrs)
    tree = ast.parse(open(os.path.join(ROOT, "jax_binding", "cagpjax_b200_jax", "ops.py")).read())
    seen = set()
    for node in ast.walk(tree):
        # the call of the callable returned by jax.ffi.ffi_call(...)
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Call)):
            continue
        inner = node.func
        if not (isinstance(inner.func, ast.Attribute) and inner.func.attr == "ffi_call"):
            continue
        name = inner.args[0].value
        spec = inner.args[1]
        n_results = len(spec.elts) if isinstance(spec, ast.Tuple) else 1
        n_args = len(node.args)
        attrs = [kw.arg for kw in node.keywords]
        assert name in bindings, name
        assert (n_args, n_results) == bindings[name][:2], (name, n_args, n_results, bindings[name])
        assert attrs == bindings[name][2], (name, attrs, bindings[name][2])   # XLA matches attributes by name; keep the order too
        seen.add(name)
    assert len(seen) >= 11, sorted(seen)   # the peer variants and the cross-covariance pair are bound from other layers


def test_xla_ffi_shim_compiles_against_a_mock_of_the_ffi_api(tmp_path):
    """With tests/mock_xla on the include path the shim is NOT empty: it must compile (C++17, -Wall -Werror) against a
    mock of xla/ffi/api/ffi.h (written from memory, so this is not a build against the real header) AND the real
    include/cagp_b200.h.  That checks the syntax of every handler, the argument count / types of every call into the
    C ABI, and - through the mock's static_assert - that each handler's parameters match its binding.  A deliberately
    broken binding and a deliberately wrong C-ABI call must both fail, so the check is not vacuous."""
    import shutil
    import subprocess

    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    shim = os.path.join(ROOT, "cagpjax_b200", "csrc", "xla_ffi_shim.cc")
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("CUDA headers not available")
    base = ["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "mock_xla"), "-I" + cuda_inc,
            "-I" + os.path.join(ROOT, "include")]
    obj = tmp_path / "shim_mock.o"
    subprocess.run(base + ["-c", shim, "-o", str(obj)], check=True)
    syms = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
    handlers = re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((Cagp\w+),", open(shim).read())
    assert all(re.search(rf"\bT {h}\b", syms) for h in handlers), "every handler symbol must be exported"
    src = open(shim).read().replace('#include "../../include/cagp_b200.h"', '#include "cagp_b200.h"')
    broken_binding = src.replace("XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpBoundingBox, BoundingBox, CAGP_BIND() ARG RET);",
                                 "XLA_FFI_DEFINE_HANDLER_SYMBOL(CagpBoundingBox, BoundingBox, CAGP_BIND() ARG ARG RET);")
    broken_call = src.replace("(int32_t)k, B->untyped_data(), ws->untyped_data(),", "(int32_t)k, B->untyped_data(),", 1)
    assert broken_binding != src and broken_call != src
    for name, text, needle in (("binding", broken_binding, "does not match its binding"), ("call", broken_call, "error")):
        f = tmp_path / f"broken_{name}.cc"
        f.write_text(text)
        res = subprocess.run(base + ["-fsyntax-only", str(f)], capture_output=True, text=True)
        assert res.returncode != 0 and needle in res.stderr, name


def test_jax_binding_parses_and_fails_cleanly_without_jax():
    import ast
    import glob
    import importlib.util

    files = sorted(glob.glob(os.path.join(ROOT, "jax_binding", "cagpjax_b200_jax", "*.py")))
    assert len(files) >= 6
    for f in files:
        ast.parse(open(f).read(), f)
    if importlib.util.find_spec("jax") is None:
        sys.path.insert(0, os.path.join(ROOT, "jax_binding"))
        try:
            with pytest.raises(ImportError, match="needs jax"):
                import cagpjax_b200_jax  # noqa: F401
        finally:
            sys.path.pop(0)


def test_no_cpu_fallback_on_kernel_path():
    x = _rand(20, 2)
    kernel = gpx.RBF()
    op = LazyKernel(kernel, x, x)
    with pytest.raises(_native.NativeLibraryError, match="no CPU fallback"):
        op @ torch.ones(20, dtype=DT)
    with pytest.raises(_native.NativeLibraryError, match="no CPU fallback"):
        (op @ BlockDiagonalSparse(_rand(20), 4)).to_dense()


def test_native_argument_errors_are_reported_not_thrown():
    import ctypes

    lib = _native.load()
    kd = _native.KernelDesc(7, 1, 1, 0, 0, None)
    rc = lib.cagp_scale_inputs(ctypes.byref(kd), None, 10, 2, None, None, 10, None)
    assert rc == 1 and b"family" in lib.cagp_last_error_string()
    assert lib.cagp_project_rows(1, None, 1, 1, 0, 1, 1, None, None, None, None, None) == 1


# ---- BlockDiagonalSparse (reference tests/test_operators.py:117-139) ------------------------------
@pytest.mark.parametrize("shape", [(5, 2), (9, 4), (6, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_block_diagonal_sparse_operator(shape, dtype):
    n, k = shape
    nz = _rand(n, dtype=dtype)
    op = BlockDiagonalSparse(nz, k)
    assert op.shape == shape and op.dtype == dtype and op.isa(ScaledOrthogonal)
    mat = op.to_dense()
    assert torch.equal(mat, O.bds_dense(nz.double(), k).to(dtype))
    assert torch.allclose(torch.eye(n, dtype=dtype) @ op, mat)
    assert torch.allclose(op @ torch.eye(k, dtype=dtype), mat)
    assert op.T.shape == shape[::-1] and torch.allclose(op.T.to_dense(), mat.T)
    assert (op @ torch.ones(k, dtype=dtype)).dtype == dtype and (op.T @ torch.ones(n, dtype=dtype)).dtype == dtype


def test_block_diagonal_sparse_grad():
    nz = _rand(9).requires_grad_(True)
    x = _rand(4, seed=1).requires_grad_(True)
    f = lambda nz, x: torch.prod(torch.sin(BlockDiagonalSparse(nz, 4) @ x))
    assert torch.autograd.gradcheck(f, (nz, x))
    y = _rand(9, seed=2).requires_grad_(True)
    g = lambda nz, y: torch.prod(torch.sin(BlockDiagonalSparse(nz, 4).T @ y))
    assert torch.autograd.gradcheck(g, (nz, y))


# ---- diag_like / congruence / lower_cholesky / jitter (reference tests/test_operators.py:157-180,
#      tests/test_linalg.py:30-85, 276-416) ---------------------------------------------------------
def test_diag_like():
    ref = ops.Dense(torch.ones(5, 5, dtype=torch.float32))
    a = diag_like(ref, 2.5)
    assert isinstance(a, ops.ScalarMul) and a.shape == (5, 5) and a.dtype == torch.float32
    assert torch.allclose(a.to_dense(), 2.5 * torch.eye(5))
    b = diag_like(ref, torch.arange(1.0, 6.0) ** 2)
    assert isinstance(b, ops.Diagonal) and torch.allclose(b.to_dense(), torch.diag(torch.arange(1.0, 6.0) ** 2))


@pytest.mark.parametrize("n,k", [(7, 3), (10, 2)])
def test_congruence_overloads(n, k):
    A = BlockDiagonalSparse(_rand(n), k)
    B = ops.Diagonal(_rand(n, seed=1))
    C = congruence_transform(A, B)
    assert isinstance(C, ops.Diagonal) and C.shape == (k, k)
    assert torch.allclose(C.to_dense(), A.to_dense().T @ B.to_dense() @ A.to_dense())
    Cs = congruence_transform(A, ops.ScalarMul(0.3, (n, n), DT))
    assert isinstance(Cs, ops.Diagonal) and torch.allclose(Cs.to_dense(), 0.3 * A.to_dense().T @ A.to_dense())
    Dd = congruence_transform(ops.Diagonal(_rand(n)), ops.Diagonal(_rand(n, seed=2)))
    assert isinstance(Dd, ops.Diagonal)
    Am, Bm = _rand(n, 3), _rand(n, n, seed=3)
    assert torch.allclose(congruence_transform(Am, Bm), Am.T @ Bm @ Am)
    lazy = congruence_transform(ops.Dense(Am), ops.Dense(Bm))
    assert isinstance(lazy, ops.LinearOperator) and torch.allclose(lazy.to_dense(), Am.T @ Bm @ Am)


def test_lower_cholesky_and_jitter_types():
    d = torch.rand(6, dtype=DT) + 0.5
    L = lower_cholesky(ops.Diagonal(d), 1e-3)
    assert isinstance(L, ops.Diagonal) and torch.allclose(L.diag, torch.sqrt(d + 1e-3))
    L = lower_cholesky(ops.ScalarMul(2.0, (4, 4), DT), 1e-3)
    assert isinstance(L, ops.ScalarMul) and torch.allclose(torch.as_tensor(L.c), torch.sqrt(torch.tensor(2.001, dtype=DT)))
    assert isinstance(lower_cholesky(ops.Identity((4, 4), DT)), ops.Identity)
    R = _rand(5, 5); A = R @ R.T + torch.eye(5, dtype=DT)
    L = lower_cholesky(ops.Dense(A), 1e-3)
    assert isinstance(L, ops.Triangular) and torch.allclose(L.A @ L.A.T, A + 1e-3 * torch.eye(5, dtype=DT))
    assert isinstance(_add_jitter(ops.ScalarMul(1.0, (3, 3), DT), 0.1), ops.ScalarMul)
    assert isinstance(_add_jitter(ops.Identity((3, 3), DT), 0.1), ops.ScalarMul)
    assert isinstance(_add_jitter(ops.Diagonal(torch.ones(3, dtype=DT)), 0.1), ops.Diagonal)
    assert isinstance(_add_jitter(ops.ScalarMul(1.0, (3, 3), DT), torch.ones(3, dtype=DT)), ops.Diagonal)
    assert isinstance(_add_jitter(ops.Dense(A), 0.1), ops.Sum)


# ---- Cholesky solver on small dense operators (reference tests/test_solvers.py) --------------------
@pytest.mark.parametrize("n", [4, 10])
def test_cholesky_solver(n):
    R = _rand(n, n); A = R @ R.T + 0.1 * torch.eye(n, dtype=DT)
    solver = Cholesky(1e-6)
    st = solver.init(ops.Dense(A))
    Aj = A + 1e-6 * torch.eye(n, dtype=DT)
    b = _rand(n, seed=1)
    assert torch.allclose(solver.solve(st, b), torch.linalg.solve(Aj, b), rtol=1e-8 * n)
    assert torch.allclose(solver.logdet(st), torch.linalg.slogdet(Aj)[1])
    B = _rand(n, 3, seed=2)
    assert torch.allclose(solver.inv_congruence_transform(st, B), B.T @ torch.linalg.solve(Aj, B))
    assert torch.allclose(solver.inv_quad(st, b), b @ torch.linalg.solve(Aj, b))
    other = solver.init(ops.Diagonal(torch.rand(n, dtype=DT) + 0.5))
    assert torch.allclose(solver.trace_solve(st, other), torch.trace(torch.linalg.solve(Aj, other.to_dense() @ other.to_dense())))
    z = _rand(n, 2, seed=3)
    assert torch.allclose(solver.unwhiten(st, z), st.A @ z)
    with pytest.raises(ValueError, match="jitter must be non-negative"):
        Cholesky(-1.0)


# ---- policies (reference tests/test_policies.py:158-273) -----------------------------------------
def test_block_sparse_policy():
    for dtype in (torch.float32, torch.float64):
        pol = BlockSparsePolicy.from_random(42, 20, 3, dtype=dtype)
        assert pol.n_actions == 3 and pol.nz_values.shape == (20,) and pol.nz_values.dtype == dtype
        starts = O.block_starts(20, 3)
        for b in range(3):
            assert abs(float(pol.nz_values[starts[b]:starts[b + 1]].norm()) - 1.0) < 1e-6
    with pytest.raises(ValueError, match="num_datapoints must be at least 1"):
        BlockSparsePolicy.from_random(0, 0, 3)
    with pytest.raises(ValueError, match="n_actions must be at least 1"):
        BlockSparsePolicy(0, torch.ones(3))
    nz = _rand(10)
    pol = BlockSparsePolicy(2, gpx.non_trainable(nz))
    act = pol.to_actions(None)
    assert isinstance(act, BlockDiagonalSparse) and torch.equal(act.nz_values, nz)
    assert torch.equal(pol.to_actions(None).to_dense(), pol.to_actions(None, key=3).to_dense())
    const = BlockSparsePolicy.from_random(1, 12, 2, sampler=lambda key, shape, dtype: torch.full(shape, 3.0, dtype=dtype or DT))
    assert torch.all(const.nz_values == const.nz_values[0])


def test_orthogonalization_policy_passes_block_sparse_actions_through():
    """Reference tests/test_policies.py:473-488."""
    from cagpjax_b200.policies import DensePolicy, OrthogonalizationPolicy

    base = BlockSparsePolicy.from_random(3, 12, 3)
    pol = OrthogonalizationPolicy(base)
    act = pol.to_actions(None)
    assert isinstance(act, BlockDiagonalSparse) and pol.n_actions == 3
    assert torch.equal(act.nz_values, base.nz_values)
    dense = OrthogonalizationPolicy(DensePolicy(_rand(12, 3))).to_actions(None)
    Q = dense.to_dense()
    assert torch.allclose(Q.T @ Q, torch.eye(3, dtype=DT), atol=1e-12)


# ---- LazyKernel host properties (reference tests/test_operators.py:236-260) -----------------------
@pytest.mark.parametrize("batch_size,max_memory_mb", [(None, 0.0), (None, 32), (9, 32), (27, 0.0)])
@pytest.mark.parametrize("checkpoint", [True, False])
def test_lazy_kernel_host_properties(batch_size, max_memory_mb, checkpoint):
    x1, x2 = _rand(9, 2), _rand(5, 2, seed=1)
    kernel = gpx.RBF(lengthscale=torch.tensor(1.0, dtype=DT), variance=torch.tensor(1.0, dtype=DT))
    op = LazyKernel(kernel, x1, x2, batch_size=batch_size, max_memory_mb=max_memory_mb, checkpoint=checkpoint)
    assert op.shape == (9, 5) and op.dtype == kernel(x1[0], x2[0]).dtype
    assert op.kernel is kernel and op.x1 is x1 and op.x2 is x2 and op.checkpoint is checkpoint
    if batch_size is None:
        if max_memory_mb == 0.0:
            assert op.batch_size_row == 1 and op.batch_size_col == 1
        else:
            assert op.batch_size_row > 1 and op.batch_size_col > 1
    else:
        assert op.batch_size_row == batch_size and op.batch_size_col == batch_size
    assert op.T.shape == (5, 9)
    eng = cagp.computations.LazyKernelComputation(batch_size=5)
    g = eng.gram(kernel, x1)
    assert isinstance(g, LazyKernel) and g.isa(ops.PSDAnnotation) and g.batch_size_row == 5 and g.is_gram
    c = eng.cross_covariance(kernel, x1, x2)
    assert isinstance(c, LazyKernel) and not c.is_gram and c.shape == (9, 5)


def test_model_requires_supervised_data():
    kernel = gpx.RBF()
    model = cagp.ComputationAwareGP(gpx.Prior(kernel, gpx.Zero()) * gpx.Gaussian(10), BlockSparsePolicy.from_random(0, 10, 2))
    x, y = _rand(10, 1), _rand(10, 1, seed=1)
    for ds in (gpx.Dataset(None, y), gpx.Dataset(x, None), gpx.Dataset(None, None)):
        with pytest.raises(ValueError, match="Training data must be supervised"):
            model.init(ds)
    assert isinstance(model.solver, Cholesky) and model.solver.jitter == 1e-6


def test_row_shard_partition():
    for n, w in [(10, 3), (1_000_000, 8), (7, 7), (5, 8)]:
        shards = [_fused.RowShard.for_rank(n, r, w) for r in range(w)]
        assert shards[0].begin == 0 and shards[-1].end == n
        assert all(a.end == b.begin for a, b in zip(shards[:-1], shards[1:]))
        assert max(s.end - s.begin for s in shards) - min(s.end - s.begin for s in shards) <= 1
    v = _rand(1037)
    assert torch.allclose(_fused.segment_sum(v, 10), O.congruence_bds_diag(torch.ones(1037, dtype=DT), 10, v))


# ---- closed-form ELBO on k-sized statistics ---------------------------------------------------------
@pytest.mark.parametrize("k,n", [(5, 60), (12, 500)])
def test_closed_form_elbo_matches_autograd_and_oracle(k, n):
    from cagpjax_b200._closed_form import elbo_from_stats, elbo_from_stats_autograd

    R = _rand(k, k); A = R @ R.T / k
    M = _rand(n, k, seed=1) * 0.3; yt = _rand(n, seed=2)
    B, c = M.T @ M, M.T @ yt
    r, q = _rand(k, seed=3), torch.rand(k, dtype=DT) + 0.5
    ins = [A, B, c, r, q, yt @ yt, torch.tensor(1.3, dtype=DT), torch.tensor(0.4, dtype=DT)]

    def run(f):
        xs = [t.clone().requires_grad_(True) for t in ins]
        v = f(xs[0], xs[1], xs[2], xs[3], xs[4], xs[5], n, xs[6], xs[7], 1e-6)
        return v.detach(), torch.autograd.grad(v, xs)

    v1, g1 = run(elbo_from_stats_autograd)
    v2, g2 = run(elbo_from_stats)
    v3 = O.elbo_from_stats(dict(A=A, B=B, c=c, r=r, q=q, yy=yt @ yt, n=n), 1.3, 0.4, 1e-6)
    assert torch.allclose(v1, v2, rtol=1e-13) and torch.allclose(v2, v3, rtol=1e-13)
    for a, b in zip(g1, g2):
        if a.dim() == 2:
            a, b = (a + a.T) / 2, (b + b.T) / 2
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-12 * float(a.abs().max()))


def test_bench_contract_on_cpu():
    """bench.py: the reference arm runs on the host and prints ONE JSON line with the contract's keys; the product
    arm refuses to run without a CUDA device (no CPU fallback)."""
    import json
    import subprocess
    import sys

    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-n", "256"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert ref.returncode == 0, ref.stderr[-2000:]
    lines = [ln for ln in ref.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["metric"] == "elbo_grad_steps_per_sec" and out["unit"] == "steps/s"
    assert out["higher_is_better"] is True and out["vs_baseline"] is None and "workload" in out["config"]
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] >= 1 and out["cpu_baseline"]["sample"]
    assert out["e2e"] == {"value": out["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not torch.cuda.is_available():
        prod = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                              capture_output=True, text=True, cwd=ROOT, timeout=300)
        assert prod.returncode != 0 and "no CPU fallback" in (prod.stderr + prod.stdout)


def test_interop_wrapper_roundtrip():
    """reference interop.py:13-81 / tests/test_operators.py TestUtils: lazify returns an operator of the same shape and
    dtype; the lineax-shaped wrapper exposes mv / as_matrix / transpose / structures and the PSD tag."""
    from cagpjax_b200.interop import ColaLinearOperator, is_positive_semidefinite, is_symmetric, lazify, to_lineax

    A = torch.randn(5, 3, dtype=torch.float64)
    op = lazify(A)
    assert op.shape == (5, 3) and op.dtype == torch.float64
    w = to_lineax(op)
    assert isinstance(w, ColaLinearOperator) and lazify(w) is op and to_lineax(w) is w
    v = torch.randn(3, dtype=torch.float64)
    assert torch.allclose(w.mv(v), A @ v) and torch.equal(w.as_matrix(), A)
    assert torch.equal(w.transpose().as_matrix(), A.T)
    assert w.in_structure() == ((3,), torch.float64) and w.out_structure() == ((5,), torch.float64)
    assert not is_symmetric(w)
    P = ops.PSD(ops.Dense(A.T @ A))
    assert is_symmetric(to_lineax(P)) and is_positive_semidefinite(to_lineax(P))
    d = ops.Diagonal(torch.arange(1.0, 4.0, dtype=torch.float64))
    assert lazify(to_lineax(d)) is d
    assert to_lineax(A).as_matrix().shape == (5, 3)
