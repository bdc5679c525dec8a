"""2-rank exercise of the peer-memory symmetric forward (remote st.global / red.global into another rank's panel between
two cross-rank barriers) and of the sharded backward, for compute-sanitizer:
  torchrun --nproc-per-node 2 --no-python compute-sanitizer --tool memcheck python tools/sanitize_peer.py
Checks the sharded ELBO + gradients against the single-rank result."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from cagpjax_b200 import _fused, gpx
from tests.util import make_problem, rel_err

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
neg = lambda m, d: -m.elbo(m.init(d))
for family, n, d, k in [("rbf", 2300, 3, 2), ("rbf", 4100, 3, 4), ("matern32", 1500, 8, 6)]:
    ms, ds, *_ = make_problem(family, n, d, k, normalize=True, device=f"cuda:{lr}", process_group=dist.group.WORLD)
    vs, gs = gpx.value_and_grad(neg, ms, ds)
    m1, d1, *_ = make_problem(family, n, d, k, normalize=True, device=f"cuda:{lr}")
    v1, g1 = gpx.value_and_grad(neg, m1, d1)
    err = max([rel_err(vs, v1)] + [rel_err(gs[key], g1[key]) for key in g1])
    if rank == 0:
        print(family, n, d, k, "peer panels:", any(v is not None for v in _fused._peer_panels.values()), f"max rel err {err:.2e}")
    assert err < 1e-9
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("done")
