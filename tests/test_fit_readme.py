"""README example (reference README.md:17-73; BASELINE config 1): n = 1000, d = 1, RBF, BlockSparsePolicy with
10 actions, float64, AdamW(lr = 0.01) on -ELBO.  The CUDA path's loss trajectory is compared with the same
optimiser driving the CPU oracle."""

import pytest
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import gpx
from oracle import cagp_oracle as O

pytestmark = pytest.mark.gpu


def _softplus_inv(x):
    return x + torch.log(-torch.expm1(-x))


def test_readme_example_training_trajectory(n_data=1000, n_actions=10, iters=250):  # BASELINE config 1: 250 AdamW iterations
    g = torch.Generator().manual_seed(42)
    x = torch.linspace(0, 10, n_data, dtype=torch.float64).reshape(-1, 1)
    y = torch.sin(x[:, 0]) + torch.randn(n_data, generator=g, dtype=torch.float64)
    s0 = O.normalize_by_blocks(torch.randn(n_data, generator=torch.Generator().manual_seed(43), dtype=torch.float64), n_actions)

    # --- CUDA path through the public API -------------------------------------------------------------
    dev = "cuda:0"
    prior = gpx.Prior(mean_function=gpx.Zero(), kernel=gpx.RBF(lengthscale=1.0, variance=1.0))
    likelihood = gpx.Gaussian(n_data, obs_stddev=0.1)
    policy = cagp.BlockSparsePolicy(n_actions, s0.to(dev))
    model = cagp.ComputationAwareGP(prior * likelihood, policy)
    data = gpx.Dataset(x.to(dev), y.to(dev).reshape(-1, 1))

    def negative_elbo(m, d):
        return -m.elbo(m.init(d))

    model, history = gpx.fit(model=model, objective=negative_elbo, train_data=data, optim="adamw", num_iters=iters,
                             learning_rate=0.01)
    assert history.shape == (iters,) and torch.isfinite(history).all()
    assert history[-1] < history[0]  # the objective decreases

    # --- same optimiser on the CPU oracle ---------------------------------------------------------------
    raw = [_softplus_inv(torch.tensor(v, dtype=torch.float64)).requires_grad_(True) for v in (1.0, 1.0, 0.1)]
    s = s0.clone().requires_grad_(True)
    opt = torch.optim.AdamW(raw + [s], lr=0.01, weight_decay=1e-4)
    ref = []
    for _ in range(iters):
        opt.zero_grad()
        ls, var, sig = (torch.nn.functional.softplus(r) for r in raw)
        p = O.Params(ls, var, sig, torch.tensor(0.0, dtype=torch.float64), s)
        loss = -O.elbo_literal("rbf", x, y, p, n_actions)
        loss.backward()
        opt.step()
        ref.append(loss.detach())
    ref = torch.stack(ref)
    rel = ((history.cpu() - ref).abs() / ref.abs()).max()
    assert rel < 1e-8, float(rel)
    # learned hyper-parameters agree
    got = [float(gpx.unwrap(t)) for t in (prior.kernel.lengthscale, prior.kernel.variance, likelihood.obs_stddev)]
    want = [float(torch.nn.functional.softplus(r.detach())) for r in raw]
    for a, b in zip(got, want):
        assert abs(a - b) < 1e-7 * max(1.0, abs(b))
    # predictions with the trained model
    with torch.no_grad():
        post = model.predict(model.init(data))
        assert post.mean.shape == (n_data,) and (post.variance > 0).all()


def _readme_model(n_data=1000, n_actions=10):
    g = torch.Generator().manual_seed(42)
    x = torch.linspace(0, 10, n_data, dtype=torch.float64).reshape(-1, 1)
    y = torch.sin(x[:, 0]) + torch.randn(n_data, generator=g, dtype=torch.float64)
    s0 = O.normalize_by_blocks(torch.randn(n_data, generator=torch.Generator().manual_seed(43), dtype=torch.float64), n_actions)
    dev = "cuda:0"
    prior = gpx.Prior(mean_function=gpx.Zero(), kernel=gpx.RBF(lengthscale=1.0, variance=1.0))
    model = cagp.ComputationAwareGP(prior * gpx.Gaussian(n_data, obs_stddev=0.1), cagp.BlockSparsePolicy(n_actions, s0.to(dev)))
    return model, gpx.Dataset(x.to(dev), y.to(dev).reshape(-1, 1))


@pytest.mark.parametrize("optim", ["adamw", "sgd"])
def test_fit_with_cuda_graph_matches_eager(optim, iters=25):
    """The captured step (forward kernels, k x k algebra, reverse mode, optimiser) replays to the same trajectory
    as the eager loop."""
    neg_elbo = lambda m, d: -m.elbo(m.init(d))
    lr = 0.01 if optim == "adamw" else 1e-6
    m1, data = _readme_model()
    m1, h1 = gpx.fit(model=m1, objective=neg_elbo, train_data=data, optim=optim, num_iters=iters, learning_rate=lr)
    m2, data = _readme_model()
    m2, h2 = gpx.fit(model=m2, objective=neg_elbo, train_data=data, optim=optim, num_iters=iters, learning_rate=lr,
                     cuda_graph=True)
    assert h2.shape == (iters,)
    assert ((h1 - h2).abs() / h1.abs()).max() < 1e-9
    for a, b in [(m1.posterior.prior.kernel.lengthscale, m2.posterior.prior.kernel.lengthscale),
                 (m1.policy.nz_values, m2.policy.nz_values)]:
        assert torch.allclose(gpx.unwrap(a), gpx.unwrap(b), rtol=1e-9, atol=1e-12)
