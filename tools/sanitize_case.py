"""Smallest end-to-end exercise of every kernel family, for compute-sanitizer runs (memcheck / racecheck / initcheck /
synccheck, ONE tool per gpurun call):  compute-sanitizer --tool memcheck python tools/sanitize_case.py
Shapes hit: SPREAD and SINGLE symmetric shapes, the float64 RBF fast path (csrc/ks_symx.cuh) incl. a row block split
over CTAs (atomics) and the 20-column broadcast tail, Matern expanded distance, the DMMA contractions with partial
tiles, dense + cross-covariance kernels."""
import sys, torch
sys.path.insert(0, ".")
from cagpjax_b200 import _native as N
dev = "cuda:0"; dt = torch.float64
g = torch.Generator().manual_seed(0)
for (fam, n, d, k) in [("rbf", 300, 3, 7), ("rbf", 700, 3, 2), ("rbf", 2300, 3, 2), ("matern32", 900, 8, 5), ("rbf", 1500, 2, 1)]:
    X = torch.rand(n, d, generator=g, dtype=dt).to(dev) * 3
    s = torch.randn(n, generator=g, dtype=dt).to(dev)
    ls = torch.ones(1, dtype=dt, device=dev)
    xs = N.scale_inputs(fam, X, ls, X[0]); bb = N.bounding_box(xs)
    Mt = N.ks_blocksparse_fwd(fam, xs, xs, d, s, k, ls, bb)
    Ms = N.ks_blocksparse_fwd_sym(fam, xs, d, s, k, ls, bbox=bb)
    G = torch.randn(k, n, generator=g, dtype=dt).to(dev)
    a = N.ks_blocksparse_bwd(fam, xs, xs, d, s, k, G, ls)
    b = N.ks_blocksparse_bwd_sym(fam, xs, d, s, k, G, ls, bbox=bb)
    A, c = N.project_rows(Mt, 0, n, s, s); B = N.gram_mtm(Mt)
    Gt = N.assemble_cotangent(Mt, 0, n, B, A, c, s, s)
    N.project_rows_bwd(Mt, 0, n, A, c)
    Pinv = torch.linalg.inv(B / n + torch.eye(k, dtype=dt, device=dev))
    N.predict_moments(Mt, c, Pinv.contiguous(), torch.tensor(1.3, dtype=dt, device=dev), torch.tensor(0.1, dtype=dt, device=dev))
    V = torch.randn(n, 20, generator=g, dtype=dt).to(dev)
    Y = N.ks_dense_fwd(fam, xs, xs, d, V, ls)
    N.ks_dense_bwd_ls(fam, xs, xs, d, V, Y, ls)
    C = N.cross_cov_fwd(fam, xs, xs[:, :37].contiguous(), d, ls)
    N.cross_cov_bwd(fam, xs, xs[:, :37].contiguous(), d, C, ls)
    torch.cuda.synchronize()
    print(fam, n, d, k, "fast path" if N.sym_fast_path(fam, xs, d, k, ls, bb) else "general", float((Mt - Ms).abs().max()),
          float((a[0] - b[0]).abs().max()))
print("done")
