"""Timing of the dense right-hand-side kernels. usage: dense_perf.py n m d p"""
import sys, torch
sys.path.insert(0, ".")
from cagpjax_b200 import _native as N
n, m, d, p = (int(v) for v in sys.argv[1:5])
dev = "cuda:0"; dt = torch.float64
g = torch.Generator().manual_seed(0)
X2 = torch.randn(n, d, generator=g, dtype=dt).to(dev); X1 = X2[:m].contiguous()
V = (torch.randn(n, p, generator=g, dtype=dt) / n**0.5).to(dev)
Yb = torch.randn(m, p, generator=g, dtype=dt).to(dev)
ls = torch.full((1,), float(d) ** 0.5, dtype=dt, device=dev)
xs1, xs2 = N.scale_inputs("rbf", X1, ls), N.scale_inputs("rbf", X2, ls)
def timeit(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: N.ks_dense_fwd("rbf", xs1, xs2, d, V, ls))
fl = 2.0 * m * n * p + m * n * (3 * d + 34)
print(f"dense fwd n={n} m={m} d={d} p={p}: {t:.2f} ms  {fl/t/1e9:.2f} TFLOP/s (2mnp + declared pair flops)  {m*n/t/1e6:.1f} Gpairs/s")
t = timeit(lambda: N.ks_dense_bwd_ls("rbf", xs1, xs2, d, V, Yb, ls))
print(f"dense bwd ls: {t:.2f} ms  {(2.0*m*n*p + m*n*(3*d+40))/t/1e9:.2f} TFLOP/s")
