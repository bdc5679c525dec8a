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
# dense action matrix, rows sharded
import cagpjax_b200 as cagp
from cagpjax_b200.policies import DensePolicy
from tests.util import KERNELS, make_data
for n, d, k in [(3000, 16, 32), (1001, 2, 5)]:
    x, y = make_data(n, d, 3, scale=2.0)
    S = torch.randn(n, k, generator=torch.Generator().manual_seed(22), dtype=torch.float64) / n**0.5
    t_ = lambda v: torch.as_tensor(v, dtype=torch.float64).to(f"cuda:{lr}")
    res = []
    for pg in (dist.group.WORLD, None):
        kernel = KERNELS["rbf"](lengthscale=t_(2.0), variance=t_(1.3), compute_engine=cagp.computations.LazyKernelComputation(process_group=pg))
        model = cagp.ComputationAwareGP(gpx.Prior(kernel, gpx.Constant(t_(0.5))) * gpx.Gaussian(n, obs_stddev=t_(0.2)), DensePolicy(t_(S)))
        res.append(gpx.value_and_grad(neg, model, gpx.Dataset(t_(x), t_(y).reshape(-1, 1))))
    (vs, gs), (v1, g1) = res
    errs = [rel_err(vs, v1)] + [rel_err(gs[key], g1[key]) for key in g1]
    worst = max(worst, max(errs))
    if rank == 0:
        print("dense", n, d, k, "max rel err sharded vs single:", f"{max(errs):.2e}")
t = torch.tensor([worst], device=f"cuda:{lr}"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("WORST", float(t)); assert float(t) < 1e-9
dist.destroy_process_group()
