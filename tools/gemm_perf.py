"""Device timing + accuracy of the float64 tensor-core contractions (csrc/gemm_dmma.cuh) at a bench shape, next to
torch.matmul (cuBLAS) on the same operands.  usage: gemm_perf.py [m] [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cagpjax_b200 import _native as N

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = "cuda:0"
g = torch.Generator(dev).manual_seed(0)
Mt = torch.randn(k, m, dtype=torch.float64, device=dev, generator=g) * 0.1
W = torch.randn(k, k, dtype=torch.float64, device=dev, generator=g)
W = W + W.T
Abar = torch.randn(k, k, dtype=torch.float64, device=dev, generator=g)
cbar = torch.randn(k, dtype=torch.float64, device=dev, generator=g)
s = torch.randn(m, dtype=torch.float64, device=dev, generator=g)
yt = torch.randn(m, dtype=torch.float64, device=dev, generator=g)
w = torch.randn(k, dtype=torch.float64, device=dev, generator=g)


def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
B = N.gram_mtm(Mt)
Bref = Mt @ Mt.T
t = timeit(lambda: N.gram_mtm(Mt)); tr = timeit(lambda: Mt @ Mt.T)
print(f"gram_mtm (lower-triangular DMMA): {t:.2f} ms = {m * k * k / t / 1e9:.2f} TFLOP/s on n k^2 flops executed "
      f"({2 * m * k * k / t / 1e9:.1f} full-GEMM-equivalent); cuBLAS gemm {tr:.2f} ms; rel diff {rel(B, Bref):.2e}")
G = N.assemble_cotangent(Mt, 0, m, W, Abar, cbar, s, yt)
blk = torch.clamp(torch.arange(m, device=dev) // (m // k), max=k - 1)
def ref_cot():
    Gr = W @ Mt
    Gr += Abar[blk].T * s[None, :]
    Gr += cbar[:, None] * yt[None, :]
    return Gr
Gr = ref_cot()
t = timeit(lambda: N.assemble_cotangent(Mt, 0, m, W, Abar, cbar, s, yt)); tr = timeit(lambda: W @ Mt)
print(f"assemble_cotangent (DMMA + fused outer terms): {t:.2f} ms = {2 * m * k * k / t / 1e9:.2f} TFLOP/s; cuBLAS gemm alone {tr:.2f} ms; rel diff {rel(G, Gr):.2e}")
del G, Gr
Pinv = torch.linalg.inv(Bref / m + torch.eye(k, dtype=torch.float64, device=dev)).contiguous()
var, mu = torch.tensor(1.3, dtype=torch.float64, device=dev), torch.tensor(0.4, dtype=torch.float64, device=dev)
mean, v = N.predict_moments(Mt, w, Pinv, var, mu)
Z = Pinv @ Mt
mean_ref = mu + var * (w @ Mt)
v_ref = var - var**2 * (Mt * Z).sum(0)
t = timeit(lambda: N.predict_moments(Mt, w, Pinv, var, mu)); tr = timeit(lambda: Pinv @ Mt)
print(f"predict_moments (DMMA, Z never written): {t:.2f} ms = {2 * m * k * k / t / 1e9:.2f} TFLOP/s; cuBLAS gemm alone {tr:.2f} ms; "
      f"rel diff mean {rel(mean, mean_ref):.2e} var {rel(v, v_ref):.2e}")
