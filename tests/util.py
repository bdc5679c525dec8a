"""Shared problem builders for the parity tests (same seeded inputs for oracle and CUDA path)."""

import numpy as np
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import gpx
from oracle import cagp_oracle as O

KERNELS = {"rbf": gpx.RBF, "matern12": gpx.Matern12, "matern32": gpx.Matern32, "matern52": gpx.Matern52}


def make_data(n, d, seed=0, scale=4.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * scale
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    return x, y


def make_problem(family, n, d, k, *, ard=False, seed=0, dtype=torch.float64, device="cuda:0", mean_constant=0.5,
                 obs_stddev=0.2, variance=1.3, normalize=True, process_group=None):
    """Returns (model, train_data, oracle_params, x64, y64): model on ``device`` in ``dtype``."""
    x, y = make_data(n, d, seed)
    g = torch.Generator().manual_seed(seed + 100)
    s = torch.randn(n, generator=g, dtype=torch.float64)
    if normalize:
        s = O.normalize_by_blocks(s, k)
    ls = (0.7 + 0.5 * torch.rand(d, generator=g, dtype=torch.float64)) if ard else torch.tensor(0.9, dtype=torch.float64)
    p = O.make_params(ls, variance, obs_stddev, mean_constant, s)
    to = lambda t: t.to(device=device, dtype=dtype)
    kernel = KERNELS[family](lengthscale=to(ls), variance=to(torch.tensor(variance, dtype=torch.float64)),
                             compute_engine=cagp.computations.LazyKernelComputation(process_group=process_group))
    prior = gpx.Prior(kernel=kernel, mean_function=gpx.Constant(to(torch.tensor(mean_constant, dtype=torch.float64))))
    lik = gpx.Gaussian(n, obs_stddev=to(torch.tensor(obs_stddev, dtype=torch.float64)))
    policy = cagp.BlockSparsePolicy(n_actions=k, nz_values=to(s))
    model = cagp.ComputationAwareGP(prior * lik, policy)
    data = gpx.Dataset(to(x), to(y).reshape(-1, 1))
    return model, data, p, x, y


def rel_err(a, b):
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if torch.is_tensor(b) else b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))
