"""Config C4: predictive mean and variance at m test points for n training points (kernel evals/s).
usage: predict_bench.py [n] [m] [dtype f64|f32]"""
import json, sys, time, torch
sys.path.insert(0, ".")
import bench as B
from cagpjax_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dtype = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda:0")
cfg = dict(B.CONFIGS["c3"], n=n)
x, y, s = B.synth(cfg, dev, dtype)
model, data = B.build_model(cfg, x, y, s, dev, dtype)
g = torch.Generator().manual_seed(14)
xt = (torch.rand(m, cfg["d"], generator=g, dtype=torch.float64) * 10).to(dev, dtype)
with torch.no_grad():
    state = model.init(data)
    def run():
        q = model.predict(state, xt)
        return q.mean, q.variance
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    _native.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        mean, var = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = {}
    for name, a, b, w in _native.PROFILE:
        prof.setdefault(name, []).append(a.elapsed_time(b))
print(json.dumps({"metric": "predict_kernel_evals_per_sec", "value": m * n / (ms * 1e-3), "unit": "evals/s", "ms": ms,
                  "config": {"workload": f"predict mean+variance, n={n} train, m={m} test, d=3, RBF, k=1024", "dtype": str(dtype)},
                  "kernel_ms": {k: sum(v) / len(v) for k, v in prof.items()},
                  "mean_finite": bool(torch.isfinite(mean).all()), "var_min": float(var.min()), "var_max": float(var.max())}))
