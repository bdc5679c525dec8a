"""Operator-level consistency of LazyKernel on the CUDA path, re-expressing the reference's own operator tests
(reference tests/test_operators.py:24-51 `_test_linear_operator_consistency`, :183-273 `TestLazyKernel`): I @ op,
op @ I, transpose, dtype, to_dense == kernel cross-covariance, and the batching / memory knobs not changing results.
Plus: mixed-dtype inputs (float32 data with the library's float64 defaults) and memoisation under in-place updates."""

import numpy as np
import pytest
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import _native as N
from cagpjax_b200 import gpx
from cagpjax_b200.operators import BlockDiagonalSparse, LazyKernel
from oracle import cagp_oracle as O
from tests.util import KERNELS, make_problem, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = {torch.float32: 1e-5, torch.float64: 1e-12}


def _kernel(family, dtype, ls=1.0, var=1.0):
    t = lambda v: torch.tensor(v, dtype=dtype, device=DEV)
    return KERNELS[family](lengthscale=t(ls), variance=t(var))


def _inputs(shape, n_dims, dtype, seed=42):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(shape[0], n_dims, generator=g, dtype=torch.float64).to(DEV, dtype)
    x2 = torch.randn(shape[1], n_dims, generator=g, dtype=torch.float64).to(DEV, dtype)
    return x1, x2


def _dense_reference(family, x1, x2, ls, var):
    return O.cross_covariance(family, x1.double().cpu(), x2.double().cpu(), torch.tensor(ls, dtype=torch.float64),
                              torch.tensor(var, dtype=torch.float64))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(9, 5), (7, 10), (300, 41)])
@pytest.mark.parametrize("batch_size", [1, 3, None])
@pytest.mark.parametrize("max_memory_mb", [0.0, 2**5])
@pytest.mark.parametrize("family", ["rbf", "matern32"])
def test_linear_operator_consistency(dtype, shape, batch_size, max_memory_mb, family):
    """reference tests/test_operators.py:246-257: identity products, transpose, dtype, dense agreement - for every
    combination of the batching knobs (kept as inert API parameters here: results must not depend on them)."""
    x1, x2 = _inputs(shape, 2, dtype)
    bs = None if batch_size is None else batch_size * max(*shape)
    kernel = _kernel(family, dtype, ls=0.8, var=1.7)
    op = LazyKernel(kernel, x1, x2, batch_size=bs, max_memory_mb=max_memory_mb)
    assert op.dtype == kernel(x1[0], x2[0]).dtype == dtype
    if bs is None:
        assert (op.batch_size_row == 1 and op.batch_size_col == 1) if max_memory_mb == 0.0 else \
            (op.batch_size_row > 1 and op.batch_size_col > 1)
    else:
        assert op.batch_size_row == bs and op.batch_size_col == bs
    mat = op.to_dense()
    atol = ATOL[dtype]
    eye_m = torch.eye(shape[0], dtype=dtype, device=DEV)
    eye_n = torch.eye(shape[1], dtype=dtype, device=DEV)
    assert torch.allclose(eye_m @ op, mat, atol=atol)          # _rmatmat
    assert torch.allclose(op @ eye_n, mat, atol=atol)          # _matmat
    opT = op.T
    assert opT.shape == op.shape[::-1]
    assert torch.allclose(opT.to_dense(), mat.T, atol=atol)
    assert mat.dtype == op.dtype
    assert (op @ torch.ones(shape[1], dtype=dtype, device=DEV)).dtype == op.dtype
    assert (op.T @ torch.ones(shape[0], dtype=dtype, device=DEV)).dtype == op.dtype
    ref = _dense_reference(family, x1, x2, 0.8, 1.7)           # test_consistency_with_dense
    assert rel_err(mat.double(), ref) < (1e-12 if dtype == torch.float64 else 2e-6)
    # the knobs do not change results: bitwise equal to the default operator
    assert torch.equal(LazyKernel(kernel, x1, x2).to_dense(), mat)


@pytest.mark.parametrize("n,dtype", [(20_000, torch.float64), (40_000, torch.float32)])
def test_large_kernel_matvec_with_grad(n, dtype):
    """reference tests/test_operators.py:275-313: finite loss and a reverse-mode gradient that agrees with central
    differences (rtol 1e-6 / 1e-2), at the reference's own sizes."""
    g = torch.Generator().manual_seed(42)
    x1, x2 = (torch.randn(n, 2, generator=g, dtype=torch.float64).to(DEV, dtype) for _ in range(2))
    v = torch.randn(n, generator=g, dtype=torch.float64).to(DEV, dtype)
    ls0, var0 = (float(t) for t in torch.rand(2, generator=g, dtype=torch.float64) * 0.8 + 0.2)

    def loss(ls, var):
        kernel = KERNELS["rbf"](lengthscale=ls, variance=var)
        return torch.vdot(v, LazyKernel(kernel, x1, x2, max_memory_mb=2**8, checkpoint=True) @ v)

    ls = torch.tensor(ls0, dtype=dtype, device=DEV, requires_grad=True)
    var = torch.tensor(var0, dtype=dtype, device=DEV, requires_grad=True)
    val = loss(ls, var)
    assert torch.isfinite(val)
    gl, gv = torch.autograd.grad(val, [ls, var])
    eps = 1e-6 if dtype == torch.float64 else 1e-2
    rtol = 1e-6 if dtype == torch.float64 else 1e-2
    with torch.no_grad():
        fd_l = (loss(ls + eps, var) - loss(ls - eps, var)) / (2 * eps)
        fd_v = (loss(ls, var + eps) - loss(ls, var - eps)) / (2 * eps)
    assert abs(float(gl - fd_l)) <= rtol * max(1.0, abs(float(fd_l)))
    assert abs(float(gv - fd_v)) <= rtol * max(1.0, abs(float(fd_v)))


def test_float32_data_with_float64_library_defaults():
    """float32 X / y with a default ``gpx.Constant()`` mean (float64) and ``from_random`` actions
    (float64): every operand of a native call is cast to the dtype of X instead of being read with the wrong width."""
    n, d, k = 1500, 3, 12
    g = torch.Generator().manual_seed(7)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 4
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    x32, y32 = x.to(DEV, torch.float32), y.to(DEV, torch.float32)
    kernel = gpx.RBF()                                   # float64 defaults
    prior = gpx.Prior(kernel=kernel, mean_function=gpx.Constant())
    policy = cagp.BlockSparsePolicy(k, torch.randn(n, generator=g, dtype=torch.float64).to(DEV))   # float64 actions
    model = cagp.ComputationAwareGP(prior * gpx.Gaussian(n, obs_stddev=0.3), policy)
    data = gpx.Dataset(x32, y32.reshape(-1, 1))
    val, grads = gpx.value_and_grad(lambda m, dd: -m.elbo(m.init(dd)), model, data)
    assert torch.isfinite(val)
    s = gpx.unwrap(policy.nz_values).detach().double().cpu()
    p = O.make_params(1.0, 1.0, 0.3, 0.0, s)
    ref = O.elbo_literal("rbf", x32.double().cpu(), y32.double().cpu(), p, k)
    assert rel_err(-val.double(), ref) < 2e-3            # float32 pair kernels against the float64 oracle
    for gname, gval in grads.items():
        assert torch.isfinite(gval).all(), gname
    # the boundary itself refuses mixed operands instead of misreading them
    xs = N.scale_inputs("rbf", x32, torch.ones(1, dtype=torch.float64, device=DEV))
    with pytest.raises(N.NativeLibraryError, match="one dtype"):
        N.ks_blocksparse_fwd("rbf", xs, xs, d, torch.ones(n, dtype=torch.float64, device=DEV), k,
                             torch.ones(1, device=DEV))


def test_projection_memo_sees_in_place_updates():
    """the memoised projection is keyed on (storage, version), so an optimiser's in-place update
    of nz_values on a held operator is not served a stale M'."""
    n, d, k = 900, 2, 6
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(n, d, generator=g, dtype=torch.float64) * 3).to(DEV)
    nz = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    op = LazyKernel(_kernel("rbf", torch.float64), x, x)
    op.x2 = op.x1
    actions = BlockDiagonalSparse(nz, k)
    with torch.no_grad():
        A1 = op.project(actions).A.clone()
        nz.mul_(2.0)                                   # in place: same tensor object, same storage
        A2 = op.project(actions).A
    assert torch.allclose(A2, 4.0 * A1, rtol=1e-12)
