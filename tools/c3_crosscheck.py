"""Full-size C3 cross-check: symmetric Gram kernels vs the independent rectangular kernels (values and gradients)."""
import sys, torch
sys.path.insert(0, ".")
import bench as B
from cagpjax_b200 import gpx, _fused
cfg = B.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda:0"); dtype = torch.float64
x, y, s = B.synth(cfg, dev, dtype)
res = []
for sym in (True, False):
    _fused.USE_SYMMETRIC = sym
    model, data = B.build_model(cfg, x, y, s, dev, dtype)
    val, grads = gpx.value_and_grad(B.neg_elbo, model, data)
    res.append((val, grads))
(v1, g1), (v2, g2) = res
print("ELBO sym", float(v1), "rect", float(v2), "rel", abs(float(v1 - v2)) / abs(float(v2)))
for k in g1:
    a, b = g1[k].reshape(-1), g2[k].reshape(-1)
    print(k, "max rel diff", float((a - b).abs().max() / b.abs().max()), "norm", float(b.norm()))
