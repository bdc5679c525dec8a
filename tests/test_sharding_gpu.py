"""GPU tests: (1) row shards emulated sequentially on one GPU add up to the unsharded statistics and
gradients (the multi-rank reduction is tested with gloo on CPU); (2) the CUDA path reproduces the
committed golden fixtures; (3) size-independent properties at a larger size."""

import os

import numpy as np
import pytest
import torch

from cagpjax_b200 import _fused, _native as N, gpx
from oracle import cagp_oracle as O
from tests.golden.make_golden import CASES, make_case
from tests.util import KERNELS, make_problem, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "cagp_golden.npz"))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("family,n,d,k", [("rbf", 1037, 3, 10), ("matern32", 700, 8, 64)])
def test_emulated_row_shards_sum_to_unsharded(world, family, n, d, k):
    g = torch.Generator().manual_seed(0)
    X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 3).to(DEV)
    s = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    yt = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    ls = torch.tensor([0.9], dtype=torch.float64, device=DEV)
    Abar, Bbar = (torch.randn(k, k, generator=g, dtype=torch.float64).to(DEV) for _ in range(2))
    cbar = torch.randn(k, generator=g, dtype=torch.float64).to(DEV)
    full = _fused.RowShard.full(n)
    A, B, c, *saved = _fused.native_local_forward(family, X, ls, s, yt, k, full)
    sb, yb, lb = _fused.native_local_backward(family, X, ls, s, yt, k, full, saved, Abar, Bbar, cbar)
    A2, B2, c2 = torch.zeros_like(A), torch.zeros_like(B), torch.zeros_like(c)
    sb2, yb2, lb2 = torch.zeros_like(sb), torch.zeros_like(yb), torch.zeros_like(lb)
    for r in range(world):
        sh = _fused.RowShard.for_rank(n, r, world)
        a, b, cc, *sv = _fused.native_local_forward(family, X, ls, s, yt, k, sh)
        A2 += a; B2 += b; c2 += cc
        s_p, y_p, l_p = _fused.native_local_backward(family, X, ls, s, yt, k, sh, sv, Abar, Bbar, cbar)
        sb2 += s_p; lb2 += l_p; yb2[sh.begin:sh.end] = y_p
    for got, ref in [(A2, A), (B2, B), (c2, c), (sb2, sb), (yb2, yb), (lb2, lb)]:
        assert rel_err(got, ref) < 1e-12


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_path_reproduces_golden_vectors(case):
    name, family, n, d, k, ard, normalize = case
    import cagpjax_b200 as cagp

    x, y, s, ls = (torch.from_numpy(GOLDEN[f"{name}/{key}"]).to(DEV) for key in ("x", "y", "s", "lengthscale"))
    t = lambda v: torch.tensor(v, dtype=torch.float64, device=DEV)
    kernel = KERNELS[family](lengthscale=ls, variance=t(1.3))
    model = cagp.ComputationAwareGP(gpx.Prior(kernel, gpx.Constant(t(0.5))) * gpx.Gaussian(n, obs_stddev=t(0.2)),
                                    cagp.BlockSparsePolicy(k, s))
    data = gpx.Dataset(x, y.reshape(-1, 1))
    with torch.no_grad():
        st = model.init(data)
        q = model.predict(st)
        assert rel_err(q.mean, GOLDEN[f"{name}/mean"]) < 1e-9
        assert rel_err(q.variance, GOLDEN[f"{name}/var"]) < 1e-9
        assert rel_err(model.prior_kl(st), GOLDEN[f"{name}/kl"]) < 1e-9
    val, grads = gpx.value_and_grad(lambda m, dd: m.elbo(m.init(dd)), model, data)
    assert rel_err(val, GOLDEN[f"{name}/elbo"]) < 1e-10
    key_of = {"lengthscale": "lengthscale", "variance": "variance", "obs_stddev": "obs_stddev", "constant": "mean_constant", "nz_values": "actions"}
    for gname, g in grads.items():
        assert rel_err(g.reshape(-1), GOLDEN[f"{name}/grad_{key_of[gname.split('.')[1]]}"].reshape(-1)) < 1e-8, gname


def test_large_size_properties():
    """n = 200k (too big for the dense oracle): linearity in s, symmetry of S^T K S, structured-oracle
    agreement on a row sample, finite ELBO and gradients."""
    n, d, k = 200_000, 3, 256
    g = torch.Generator().manual_seed(1)
    X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 10).to(DEV)
    s1 = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    s2 = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    ls = torch.ones(1, dtype=torch.float64, device=DEV)
    xs = N.scale_inputs("rbf", X, ls)
    M1 = N.ks_blocksparse_fwd("rbf", xs, xs, d, s1, k, ls)
    M2 = N.ks_blocksparse_fwd("rbf", xs, xs, d, s2, k, ls)
    M12 = N.ks_blocksparse_fwd("rbf", xs, xs, d, (2.0 * s1 - 0.5 * s2).contiguous(), k, ls)
    assert rel_err(M12, 2.0 * M1 - 0.5 * M2) < 1e-12  # linearity in the actions
    A, _ = N.project_rows(M1, 0, n, s1, s1)
    assert rel_err(A, A.T) < 1e-11  # S^T K S is symmetric
    rows = torch.arange(0, n, 997, device=DEV)
    ref = O.projected_kernel_blocksparse("rbf", X[rows].cpu().numpy(), X.cpu().numpy(), np.ones(1), s1.cpu().numpy(), k)
    assert rel_err(M1[:, rows].T, ref) < 1e-11


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("n,k", [(3001, 130), (2048, 64), (900, 7)])
def test_peer_cotangent_gemm_emulated_ranks(world, n, k):
    """cagp_assemble_cotangent_peer with the ranks emulated one after another on one GPU (plain buffers stand in for the
    peer-mapped ones): afterwards rank h's (k, n) buffer holds G'[i in own rows, :] and G'[:, a in own blocks] -
    exactly what its symmetric backward tiles read - and agrees with the unsharded cotangent."""
    g = torch.Generator().manual_seed(5)
    dt = torch.float64
    Mt = (torch.randn(k, n, generator=g, dtype=dt) * 0.3).to(DEV)
    W = torch.randn(k, k, generator=g, dtype=dt).to(DEV)
    Abar = torch.randn(k, k, generator=g, dtype=dt).to(DEV)
    cbar = torch.randn(k, generator=g, dtype=dt).to(DEV)
    s = torch.randn(n, generator=g, dtype=dt).to(DEV)
    yt = torch.randn(n, generator=g, dtype=dt).to(DEV)
    G_ref = N.assemble_cotangent(Mt, 0, n, W, Abar, cbar, s, yt)            # (k, n), unsharded
    shards = [_fused.RowShard.for_rank_blocks(n, k, r, world) for r in range(world)]
    block_begin = [sh.a_begin for sh in shards] + [k]
    bufs = [torch.full((k, n), float("nan"), dtype=dt, device=DEV) for _ in range(world)]
    for r, sh in enumerate(shards):
        Ml = Mt[:, sh.begin:sh.end].contiguous()
        N.assemble_cotangent_peer(Ml, sh.begin, n, W, Abar, cbar, s[sh.begin:sh.end].contiguous(),
                                  yt[sh.begin:sh.end].contiguous(), [b.data_ptr() for b in bufs], block_begin, r, n)
    torch.cuda.synchronize()
    for r, sh in enumerate(shards):
        assert rel_err(bufs[r][:, sh.begin:sh.end], G_ref[:, sh.begin:sh.end]) < 1e-13     # own rows, all blocks
        assert rel_err(bufs[r][sh.a_begin:sh.a_end], G_ref[sh.a_begin:sh.a_end]) < 1e-13   # own blocks, all rows
        # nothing else was written: the rest of the buffer still holds the NaN canary
        mask = torch.ones((k, n), dtype=torch.bool, device=DEV)
        mask[:, sh.begin:sh.end] = False
        mask[sh.a_begin:sh.a_end] = False
        assert torch.isnan(bufs[r][mask]).all()
