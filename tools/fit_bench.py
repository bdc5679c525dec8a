"""README-shape training loop (BASELINE config 1: n = 1000, d = 1, RBF, k = 10, f64, AdamW): iterations per second
of gpx.fit, eager vs CUDA-graph replay.  usage: fit_bench.py [iters]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.test_fit_readme import _readme_model  # noqa: E402
from cagpjax_b200 import gpx  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 250
neg_elbo = lambda m, d: -m.elbo(m.init(d))
for graph in (False, True):
    model, data = _readme_model()
    gpx.fit(model=model, objective=neg_elbo, train_data=data, num_iters=5, cuda_graph=graph)  # warm-up (cuBLAS, allocator)
    model, data = _readme_model()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model, hist = gpx.fit(model=model, objective=neg_elbo, train_data=data, num_iters=iters, cuda_graph=graph)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"cuda_graph={graph}: {iters} iterations in {dt * 1e3:.1f} ms = {iters / dt:.0f} it/s, "
          f"objective {float(hist[0]):.4f} -> {float(hist[-1]):.4f}")
