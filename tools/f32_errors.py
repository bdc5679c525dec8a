"""Actual float32 errors of the model-level quantities against the float64 oracle ON THE SAME (float32-rounded) INPUTS
(north star: 1e-4 in float32).  Prints, per case: ELBO, every gradient, predictive mean / variance."""
import sys, torch
sys.path.insert(0, ".")
from cagpjax_b200 import gpx
from oracle import cagp_oracle as O
from tests.util import make_problem, rel_err
neg = lambda m, d: -m.elbo(m.init(d))
names = {"lengthscale": "lengthscale", "variance": "variance", "obs_stddev": "obs_stddev", "constant": "mean_constant", "nz_values": "actions"}
for family in ("rbf", "matern32"):
    for n, d, k in [(9, 1, 4), (1037, 3, 10), (700, 8, 64), (5000, 3, 40)]:
        model, data, p, x, y = make_problem(family, n, d, k, dtype=torch.float32, normalize=True, obs_stddev=0.5)
        # the oracle sees what the kernels see: inputs rounded to float32, arithmetic in float64
        x32, y32 = data.X.double().cpu(), data.y.double().cpu().reshape(-1)
        p.actions = model.policy.nz_values.detach().double().cpu() if hasattr(model.policy, "nz_values") else p.actions
        val, grads = gpx.value_and_grad(neg, model, data)
        oval, ograds = O.elbo_and_grad_literal(family, x32, y32, p, k)
        errs = {"elbo": rel_err(-val, oval)}
        for name, g in grads.items():
            errs[name.split(".")[1]] = rel_err(-g.reshape(-1), ograds[names[name.split(".")[1]]].reshape(-1))
        with torch.no_grad():
            q = model.predict(model.init(data))
            ost = O.cagp_init(family, x32, y32, p, k)
            om, ov = O.cagp_predict(family, ost, p)
            errs["mean"] = rel_err(q.mean, om); errs["var"] = rel_err(q.variance, ov)
        print(family, n, d, k, " ".join(f"{a}={b:.1e}" for a, b in errs.items()), flush=True)
