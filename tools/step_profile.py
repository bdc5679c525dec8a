"""torch.profiler breakdown of one ELBO+grad step. usage: step_profile.py CONFIG"""
import sys, torch
sys.path.insert(0, ".")
import bench as B
from cagpjax_b200 import gpx
from torch.profiler import profile, ProfilerActivity
cfg = B.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0"); dtype = torch.float64
x, y, s = B.synth(cfg, dev, dtype)
model, data = B.build_model(cfg, x, y, s, dev, dtype)
for _ in range(3):
    gpx.value_and_grad(B.neg_elbo, model, data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        gpx.value_and_grad(B.neg_elbo, model, data)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
