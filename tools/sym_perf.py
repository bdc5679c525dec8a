"""Symmetric Gram kernels: time and accuracy of the bounding-box fast paths (safe range / expanded distance form)
against the general path (no bounding box).  usage: sym_perf.py [n] [d] [k] [family]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cagpjax_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
fam = sys.argv[4] if len(sys.argv) > 4 else "rbf"
dev = "cuda:0"
g = torch.Generator().manual_seed(11)
X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 10).to(dev)
s = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
ls = torch.ones(1, dtype=torch.float64, device=dev)
xs = N.scale_inputs(fam, X, ls, X[0])
bbox = N.bounding_box(xs)


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


Mt0 = N.ks_blocksparse_fwd_sym(fam, xs, d, s, k, ls)
Mt1 = N.ks_blocksparse_fwd_sym(fam, xs, d, s, k, ls, bbox=bbox)
print("fwd rel diff fast vs general:", float((Mt1 - Mt0).abs().max() / Mt0.abs().max()))
t0 = timeit(lambda: N.ks_blocksparse_fwd_sym(fam, xs, d, s, k, ls))
t1 = timeit(lambda: N.ks_blocksparse_fwd_sym(fam, xs, d, s, k, ls, bbox=bbox))
print(f"fwd general {t0:.1f} ms, with bbox {t1:.1f} ms  ({n * n / t1 / 1e6:.0f} Gpairs/s ordered)")
Gt = torch.randn(k, n, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(1))
sb0, lb0 = N.ks_blocksparse_bwd_sym(fam, xs, d, s, k, Gt, ls)
sb1, lb1 = N.ks_blocksparse_bwd_sym(fam, xs, d, s, k, Gt, ls, bbox=bbox)
print("bwd rel diff s_bar:", float((sb1 - sb0).abs().max() / sb0.abs().max()), "ls_bar:", float(((lb1 - lb0) / lb0).abs().max()))
t0 = timeit(lambda: N.ks_blocksparse_bwd_sym(fam, xs, d, s, k, Gt, ls))
t1 = timeit(lambda: N.ks_blocksparse_bwd_sym(fam, xs, d, s, k, Gt, ls, bbox=bbox))
print(f"bwd general {t0:.1f} ms, with bbox {t1:.1f} ms  ({n * n / t1 / 1e6:.0f} Gpairs/s ordered)")
