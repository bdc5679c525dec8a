"""Smallest end-to-end exercise of every kernel family, for compute-sanitizer runs."""
import sys, torch
sys.path.insert(0, ".")
from cagpjax_b200 import _native as N
dev = "cuda:0"; dt = torch.float64
g = torch.Generator().manual_seed(0)
for (n, d, k) in [(300, 3, 7), (700, 3, 2)]:
    X = torch.rand(n, d, generator=g, dtype=dt).to(dev) * 3
    s = torch.randn(n, generator=g, dtype=dt).to(dev)
    ls = torch.ones(1, dtype=dt, device=dev)
    xs = N.scale_inputs("rbf", X, ls, X[0]); bb = N.bounding_box(xs)
    Mt = N.ks_blocksparse_fwd("rbf", xs, xs, d, s, k, ls)
    Ms = N.ks_blocksparse_fwd_sym("rbf", xs, d, s, k, ls, bbox=bb)
    G = torch.randn(k, n, generator=g, dtype=dt).to(dev)
    a = N.ks_blocksparse_bwd("rbf", xs, xs, d, s, k, G, ls)
    b = N.ks_blocksparse_bwd_sym("rbf", xs, d, s, k, G, ls, bbox=bb)
    A, c = N.project_rows(Mt, 0, n, s, s); B = N.gram_mtm(Mt)
    Gt = N.assemble_cotangent(Mt, 0, n, B, A, c, s, s)
    N.project_rows_bwd(Mt, 0, n, A, c)
    V = torch.randn(n, 20, generator=g, dtype=dt).to(dev)
    Y = N.ks_dense_fwd("rbf", xs, xs, d, V, ls)
    N.ks_dense_bwd_ls("rbf", xs, xs, d, V, Y, ls)
    torch.cuda.synchronize()
    print(n, d, k, float((Mt - Ms).abs().max()), float((a[0] - b[0]).abs().max()))
print("done")
