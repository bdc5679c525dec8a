"""Third-party anchors for the oracle (CPU).

The third-party packages behind the reference (GPJax's kernel and likelihood formulas) cannot run here and are restated
both in the oracle and in the stand-ins that run the reference's own source (oracle/refshim), so those restated
formulas are additionally pinned against an INDEPENDENT published implementation that is installed in this image:
scikit-learn's ``gaussian_process`` (kernels ``RBF`` / ``Matern(nu = 1/2, 3/2, 5/2)`` and the exact
``GaussianProcessRegressor``).  scikit-learn is not GPJax, so this is an independent anchor, not a pin on GPJax itself - but a
misremembered kernel formula, a wrong ARD scaling or a wrong exact-GP limit of the CaGP posterior / ELBO in the
oracle would show up here.

What is compared:
  * ``O.cross_covariance`` (restating gpjax.kernels.stationary.*) against the sklearn kernel of the same family, scalar
    and per-dimension lengthscales, 1e-13;
  * the CaGP posterior with complete actions (k = n; reference tests/test_models.py:206-235 states this limit) against
    sklearn's exact GP regression: mean and variance at test points, 1e-9 (jitter off on the CaGP side);
  * the CaGP ELBO with complete actions against sklearn's log marginal likelihood (the bound is tight when the
    actions span everything), 1e-9 relative.
"""

import math

import numpy as np
import pytest
import torch

from oracle import cagp_oracle as O

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402

NU = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}


def sk_kernel(family, lengthscale, variance):
    ls = np.asarray(lengthscale, dtype=np.float64)
    ls = float(ls) if ls.ndim == 0 or ls.size == 1 else ls
    base = RBF(length_scale=ls) if family == "rbf" else Matern(length_scale=ls, nu=NU[family])
    return ConstantKernel(constant_value=float(variance)) * base


@pytest.mark.parametrize("family", ["rbf", "matern12", "matern32", "matern52"])
@pytest.mark.parametrize("ard", [False, True])
def test_kernel_formulas_match_sklearn(family, ard):
    rng = np.random.default_rng(3)
    d = 4
    x1 = rng.normal(size=(37, d)) * 2.0
    x2 = rng.normal(size=(23, d)) * 2.0
    ls = rng.uniform(0.4, 2.5, size=d) if ard else np.float64(0.83)
    var = 1.7
    ours = O.cross_covariance(family, torch.as_tensor(x1), torch.as_tensor(x2), torch.as_tensor(ls), torch.tensor(var, dtype=torch.float64)).numpy()
    ref = sk_kernel(family, ls, var)(x1, x2)
    assert np.max(np.abs(ours - ref)) < 1e-13 * var


@pytest.mark.parametrize("family", ["rbf", "matern32", "matern52"])
def test_complete_actions_recover_sklearn_exact_gp(family):
    rng = np.random.default_rng(11)
    n, d, m = 60, 2, 17
    x = rng.uniform(-3, 3, size=(n, d))
    y = np.sin(x[:, 0]) * np.cos(0.5 * x[:, 1]) + 0.1 * rng.normal(size=n)
    z = rng.uniform(-3, 3, size=(m, d))
    ls, var, sigma, c = 1.3, 0.9, 0.25, 0.4
    gpr = sk.GaussianProcessRegressor(kernel=sk_kernel(family, ls, var), alpha=sigma**2, optimizer=None, normalize_y=False)
    gpr.fit(x, y - c)
    mean_sk, std_sk = gpr.predict(z, return_std=True)
    mean_sk = mean_sk + c
    # k = n block-sparse actions: one non-zero per block, S = diag(+-1) after normalisation
    actions = rng.normal(size=n)
    p = O.make_params(ls, var, sigma, c, O.normalize_by_blocks(torch.as_tensor(actions), n))
    st = O.cagp_init(family, torch.as_tensor(x), torch.as_tensor(y), p, n, jitter=None)
    mean, variance = O.cagp_predict(family, st, p, torch.as_tensor(z))
    assert np.max(np.abs(mean.numpy() - mean_sk)) < 1e-9
    assert np.max(np.abs(variance.numpy() - std_sk**2)) < 1e-9
    # the ELBO is tight when the actions span everything: equals the exact log marginal likelihood of y - c
    elbo = float(O.cagp_elbo(family, st, p, jitter=None))
    lml = float(gpr.log_marginal_likelihood(gpr.kernel_.theta))
    assert abs(elbo - lml) < 1e-9 * max(1.0, abs(lml))


def test_own_exact_gp_matches_sklearn():
    """``O.exact_gp_predict`` (the anchor other tests use) against sklearn, so that chain of anchors ends outside."""
    rng = np.random.default_rng(5)
    n, m = 40, 9
    x = rng.uniform(0, 5, size=(n, 1))
    y = np.sin(2 * x[:, 0]) + 0.05 * rng.normal(size=n)
    z = rng.uniform(0, 5, size=(m, 1))
    ls, var, sigma, c = 0.7, 1.2, 0.1, -0.3
    gpr = sk.GaussianProcessRegressor(kernel=sk_kernel("matern52", ls, var), alpha=sigma**2, optimizer=None).fit(x, y - c)
    mean_sk, cov_sk = gpr.predict(z, return_cov=True)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    mean, cov = O.exact_gp_predict("matern52", t(x), t(y), t(z), t(ls), t(var), t(sigma), t(c))
    assert np.max(np.abs(mean.numpy() - (mean_sk + c))) < 1e-9
    assert np.max(np.abs(cov.numpy() - cov_sk)) < 1e-9
    assert math.isfinite(float(cov.diagonal().min()))
