"""Multi-rank parity check (torchrun): sharded ELBO + gradients vs the single-rank result computed on rank 0."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from cagpjax_b200 import gpx
from tests.util import make_problem, rel_err

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
neg = lambda m, d: -m.elbo(m.init(d))
worst = 0.0
for family, n, d, k, ard in [("rbf", 5037, 3, 10, False), ("matern32", 3000, 8, 64, True), ("rbf", 2000, 2, 3, False), ("rbf", 40000, 3, 41, False)]:
    ms, ds, *_ = make_problem(family, n, d, k, ard=ard, normalize=True, device=f"cuda:{lr}", process_group=dist.group.WORLD)
    vs, gs = gpx.value_and_grad(neg, ms, ds)
    m1, d1, *_ = make_problem(family, n, d, k, ard=ard, normalize=True, device=f"cuda:{lr}")
    v1, g1 = gpx.value_and_grad(neg, m1, d1)
    errs = [rel_err(vs, v1)] + [rel_err(gs[key], g1[key]) for key in g1]
    worst = max(worst, max(errs))
    if rank == 0:
        print(family, n, d, k, "max rel err sharded vs single:", f"{max(errs):.2e}")
t = torch.tensor([worst], device=f"cuda:{lr}"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("WORST", float(t)); assert float(t) < 1e-9
dist.destroy_process_group()
