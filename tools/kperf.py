"""Quick device timing of the pair kernels (CUDA events). usage: kperf.py [n] [m] [d] [k] [family] [dtype]"""
import sys, time
import torch
sys.path.insert(0, ".")
from cagpjax_b200 import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
d = int(sys.argv[3]) if len(sys.argv) > 3 else 3
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
fam = sys.argv[5] if len(sys.argv) > 5 else "rbf"
dtype = torch.float64 if (len(sys.argv) <= 6 or sys.argv[6] == "f64") else torch.float32
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
X2 = (torch.rand(n, d, generator=g, dtype=torch.float64) * 10).to(dev, dtype)
X1 = X2[:m].contiguous()
s = torch.randn(n, generator=g, dtype=torch.float64).to(dev, dtype)
ls = torch.ones(1, dtype=dtype, device=dev)
xs1 = N.scale_inputs(fam, X1, ls); xs2 = N.scale_inputs(fam, X2, ls)

def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

t = timeit(lambda: N.ks_blocksparse_fwd(fam, xs1, xs2, d, s, k, ls))
print(f"fwd  {fam} {dtype} n={n} m={m} d={d} k={k}: {t:.2f} ms  {m*n/t/1e6:.1f} Gpairs/s")
Gt = torch.randn(k, m, dtype=dtype, device=dev)
t = timeit(lambda: N.ks_blocksparse_bwd(fam, xs1, xs2, d, s, k, Gt, ls))
print(f"bwd  {fam} {dtype} n={n} m={m} d={d} k={k}: {t:.2f} ms  {m*n/t/1e6:.1f} Gpairs/s")
Mt = N.ks_blocksparse_fwd(fam, xs1, xs2, d, s, k, ls)
t = timeit(lambda: N.gram_mtm(Mt)); print(f"syrk k={k} m={m}: {t:.2f} ms  {m*k*k/t/1e9:.2f} TFLOP/s (n k^2 flops)")
W = torch.randn(k, k, dtype=dtype, device=dev); c = torch.randn(k, dtype=dtype, device=dev)
sl = s[:m].contiguous()
t = timeit(lambda: N.assemble_cotangent(Mt, 0, n, W, W, c, sl, sl)); print(f"assemble(gemm) : {t:.2f} ms  {2*m*k*k/t/1e9:.2f} TFLOP/s")
t = timeit(lambda: N.project_rows(Mt, 0, n, sl, sl)); print(f"project_rows: {t:.2f} ms  {Mt.numel()*Mt.element_size()/t/1e6:.0f} GB/s")
t = timeit(lambda: N.project_rows_bwd(Mt, 0, n, W, c)); print(f"project_rows_bwd: {t:.2f} ms  {Mt.numel()*Mt.element_size()/t/1e6:.0f} GB/s")
if m == n:
    bb = N.bounding_box(xs2)
    t = timeit(lambda: N.ks_blocksparse_fwd_sym(fam, xs2, d, s, k, ls)); print(f"sym fwd: {t:.2f} ms  {n*n/t/1e6:.1f} G(ordered)pairs/s")
    t = timeit(lambda: N.ks_blocksparse_fwd_sym(fam, xs2, d, s, k, ls, bbox=bb)); print(f"sym fwd (safe range): {t:.2f} ms  {n*n/t/1e6:.1f} G(ordered)pairs/s")
    Gf = torch.randn(k, n, dtype=dtype, device=dev)
    t = timeit(lambda: N.ks_blocksparse_bwd_sym(fam, xs2, d, s, k, Gf, ls)); print(f"sym bwd: {t:.2f} ms  {n*n/t/1e6:.1f} G(ordered)pairs/s")
    t = timeit(lambda: N.ks_blocksparse_bwd_sym(fam, xs2, d, s, k, Gf, ls, bbox=bb)); print(f"sym bwd (safe range): {t:.2f} ms  {n*n/t/1e6:.1f} G(ordered)pairs/s")
