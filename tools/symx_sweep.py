"""A/B sweep of the fast-path symmetric kernels (csrc/ks_symx.cuh) against the general kernels (csrc/ks_sym.cuh).
Every variant is selected with CAGP_SYMX_FWD / CAGP_SYMX_BWD (read by the library at every call), checked against the
general kernel on the same inputs and timed with CUDA events.
usage: symx_sweep.py [n] [d] [k] [--fwd 0,1,..] [--bwd 0,1,..] [--reps 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cagpjax_b200 import _native as N

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="?", type=int, default=1_000_000)
ap.add_argument("d", nargs="?", type=int, default=3)
ap.add_argument("k", nargs="?", type=int, default=1024)
ap.add_argument("--fwd", default="0,1,2,3,4,5,6")
ap.add_argument("--bwd", default="0,1,2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ls", type=float, default=1.0)
args = ap.parse_args()
n, d, k = args.n, args.d, args.k
dev = "cuda:0"
g = torch.Generator().manual_seed(11)
X = (torch.rand(n, d, generator=g, dtype=torch.float64) * 10).to(dev)
s = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
ls = torch.full((1,), args.ls, dtype=torch.float64, device=dev)
lo, hi = torch.aminmax(X, dim=0)
xs = N.scale_inputs("rbf", X, ls, torch.lerp(lo, hi, 0.5))
bbox = N.bounding_box(xs)


def timeit(f, reps):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def setv(fwd=None, bwd=None):
    if fwd is not None:
        os.environ["CAGP_SYMX_FWD"] = str(fwd)
    if bwd is not None:
        os.environ["CAGP_SYMX_BWD"] = str(bwd)


setv("off", "off")
fwd = lambda: N.ks_blocksparse_fwd_sym("rbf", xs, d, s, k, ls, bbox=bbox)
Mt_ref = fwd()
t_ref = timeit(fwd, args.reps)
print(f"fwd general: {t_ref:.1f} ms ({n * n / t_ref / 1e6:.0f} G ordered pairs/s)", flush=True)
for v in [int(x) for x in args.fwd.split(",") if x != ""]:
    setv(fwd=v)
    try:
        Mt = fwd()
        err = float((Mt - Mt_ref).abs().max() / Mt_ref.abs().max())
        t = timeit(fwd, args.reps)
        print(f"fwd variant {v}: {t:.1f} ms ({n * n / t / 1e6:.0f} G/s)  max rel diff vs general {err:.2e}", flush=True)
        del Mt
    except Exception as exc:  # noqa: BLE001
        print(f"fwd variant {v}: FAILED {exc}", flush=True)
setv(fwd="off")
Gt = torch.randn(k, n, dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(1))
bwd = lambda: N.ks_blocksparse_bwd_sym("rbf", xs, d, s, k, Gt, ls, bbox=bbox)
sb_ref, lb_ref = bwd()
t_ref = timeit(bwd, args.reps)
print(f"bwd general: {t_ref:.1f} ms ({n * n / t_ref / 1e6:.0f} G ordered pairs/s)", flush=True)
for v in [int(x) for x in args.bwd.split(",") if x != ""]:
    setv(bwd=v)
    try:
        sb, lb = bwd()
        e1 = float((sb - sb_ref).abs().max() / sb_ref.abs().max())
        e2 = float(((lb - lb_ref) / lb_ref).abs().max())
        t = timeit(bwd, args.reps)
        print(f"bwd variant {v}: {t:.1f} ms ({n * n / t / 1e6:.0f} G/s)  s_bar rel diff {e1:.2e}  ls_bar rel diff {e2:.2e}", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"bwd variant {v}: FAILED {exc}", flush=True)
