"""Throughput of the materialised cross-covariance kernels against the HBM roofline (run on a B200).

The forward writes the (m, n) matrix once (algorithmic bytes m*n*itemsize); the backward reads the cotangent once.
usage: python tools/cross_perf.py [m] [n] [d]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from cagpjax_b200 import _native as N

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda:0"
out = []
for dtype in (torch.float64, torch.float32):
    for family in ("rbf", "matern32"):
        g = torch.Generator().manual_seed(0)
        X = (torch.rand(m, d, generator=g, dtype=torch.float64) * 10).to(dev, dtype)
        Z = (torch.rand(n, d, generator=g, dtype=torch.float64) * 10).to(dev, dtype)
        ls = torch.ones(1, dtype=dtype, device=dev)
        xs1 = N.scale_inputs(family, X, ls, X[0])
        xs2 = N.scale_inputs(family, Z, ls, X[0])
        C = N.cross_cov_fwd(family, xs1, xs2, d, ls)
        Cbar = torch.randn_like(C)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, reps=5):
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]

        fwd = timed(lambda: N.cross_cov_fwd(family, xs1, xs2, d, ls))
        bwd = timed(lambda: N.cross_cov_bwd(family, xs1, xs2, d, Cbar, ls))
        nbytes = m * n * C.element_size()
        out.append(dict(family=family, dtype=str(dtype), m=m, n=n, d=d, fwd_ms=round(fwd, 3), bwd_ms=round(bwd, 3),
                        fwd_gbs=round(nbytes / fwd / 1e6, 1), bwd_gbs=round(nbytes / bwd / 1e6, 1),
                        fwd_gevals=round(m * n / fwd / 1e6, 1)))
        print(json.dumps(out[-1]), flush=True)
