"""oracle/cagp_oracle_ext.py against its golden fixtures and its independent anchors (CPU)."""

import os

import numpy as np
import pytest
import torch

from oracle import cagp_oracle as O
from oracle import cagp_oracle_ext as E
from tests.golden.make_golden_ext import CASES, evaluate

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "cagp_golden_ext.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_ext_oracle_reproduces_golden_vectors(case):
    name, family, n, d, k, policy, solver = case
    got = evaluate(family, n, d, k, policy, solver)
    for key, val in got.items():
        ref = GOLDEN[f"{name}/{key}"]
        assert np.allclose(val.numpy(), ref, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(ref).max()))), key


def test_dense_chain_matches_base_oracle_literal():
    """The dense-action chain with the Cholesky solver is the base oracle's literal path (independent code)."""
    g = torch.Generator().manual_seed(0)
    n, d, k = 60, 2, 6
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 3
    y = torch.sin(x.sum(1))
    S = torch.randn(n, k, generator=g, dtype=torch.float64) / n**0.5
    p = O.make_params(1.1, 1.3, 0.2, 0.1, S)
    for family in O.KERNEL_FAMILIES:
        ref = O.elbo_literal(family, x, y, p, None, 1e-6)
        got, mean, var = E.elbo_dense_actions(family, x, y, p.lengthscale, p.variance, p.obs_stddev, p.mean_constant, S)
        assert torch.allclose(got, ref, rtol=1e-12)
        st = O.cagp_init(family, x, y, p, None)
        om, ov = O.cagp_predict(family, st, p)
        assert torch.allclose(mean, om, rtol=1e-10) and torch.allclose(var, ov, rtol=1e-8, atol=1e-12)


def test_pinv_and_cholesky_agree_without_jitter():
    """With a well-conditioned projected system and no jitter both solvers give the same ELBO."""
    g = torch.Generator().manual_seed(1)
    n, d, k = 80, 2, 5
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 3
    y = torch.sin(x.sum(1))
    S = torch.linalg.qr(torch.randn(n, k, generator=g, dtype=torch.float64))[0]
    args = ("rbf", x, y, torch.tensor(0.8, dtype=torch.float64), torch.tensor(1.2, dtype=torch.float64),
            torch.tensor(0.3, dtype=torch.float64), torch.tensor(0.0, dtype=torch.float64), S)
    a = E.elbo_dense_actions(*args, solver="cholesky", jitter=0.0)[0]
    b = E.elbo_dense_actions(*args, solver="pinv", rtol=0.0)[0]
    assert torch.allclose(a, b, rtol=1e-10)


def test_elbo_is_invariant_to_the_basis_of_the_actions():
    """CaGP only sees span(S): orthogonalising the actions (any method) must not change the ELBO (no jitter)."""
    g = torch.Generator().manual_seed(2)
    n, d, m = 70, 1, 5
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 3
    y = torch.sin(x.sum(1))
    z = torch.rand(m, d, generator=g, dtype=torch.float64) * 3
    ls, var, sig, mu = (torch.tensor(v, dtype=torch.float64) for v in (0.4, 1.0, 0.3, 0.0))
    S = E.pseudo_input_actions("matern32", x, z, ls, var)
    ref = E.elbo_dense_actions("matern32", x, y, ls, var, sig, mu, S, solver="pinv", rtol=0.0)[0]
    for Q in (E.orthogonalize_qr(S), E.orthogonalize_cgs(S, 1), E.orthogonalize_mgs(S, 1)):
        got = E.elbo_dense_actions("matern32", x, y, ls, var, sig, mu, Q, solver="pinv", rtol=0.0)[0]
        assert torch.allclose(got, ref, rtol=1e-8)


def test_lanczos_full_run_recovers_the_spectrum():
    g = torch.Generator().manual_seed(3)
    A = torch.randn(15, 15, generator=g, dtype=torch.float64)
    A = A @ A.T
    vals, vecs = E.lanczos_eigh(A, 15, torch.randn(15, generator=g, dtype=torch.float64))
    assert torch.allclose(vals, torch.linalg.eigvalsh(A), rtol=1e-10)
    assert torch.allclose(vecs @ torch.diag(vals) @ vecs.T, A, atol=1e-9)


@pytest.mark.parametrize("policy,solver", [("pseudo", "cholesky"), ("ortho_qr", "cholesky"), ("lanczos", "pinv")])
def test_ext_oracle_gradients_match_finite_differences(policy, solver):
    """The golden gradients come from torch reverse mode through the restated chain; central differences on the
    same chain are the independent check (every leaf, including the pseudo-inputs)."""
    from tests.golden.make_golden_ext import HYPER, actions_for, make_inputs

    family, n, d, k = "rbf", 60, 2, 5
    x, y, z, _ = make_inputs(n, d, k)
    base = dict(lengthscale=torch.tensor(0.9, dtype=torch.float64), variance=torch.tensor(HYPER["variance"], dtype=torch.float64),
                obs_stddev=torch.tensor(0.3, dtype=torch.float64), mean_constant=torch.tensor(0.5, dtype=torch.float64),
                pseudo_inputs=z.clone())
    kw = dict(solver="cholesky", jitter=1e-6) if solver == "cholesky" else dict(solver="pinv")

    def f(leaves):
        S = actions_for(policy, family, x, leaves["pseudo_inputs"], leaves["lengthscale"], leaves["variance"],
                        leaves["obs_stddev"], k)
        return E.elbo_dense_actions(family, x, y, leaves["lengthscale"], leaves["variance"], leaves["obs_stddev"],
                                    leaves["mean_constant"], S, **kw)[0]

    q = {name: t.clone().requires_grad_(True) for name, t in base.items()}
    grads = torch.autograd.grad(f(q), list(q.values()), allow_unused=True)
    for (name, t), g in zip(base.items(), grads):
        if g is None:  # Lanczos actions do not depend on the pseudo-inputs
            assert policy == "lanczos" and name == "pseudo_inputs"
            continue
        flat = t.reshape(-1)
        for idx in range(min(flat.numel(), 3)):
            h = 1e-6 * max(1.0, abs(float(flat[idx])))
            vals = []
            for sign in (+1, -1):
                pert = {kname: v.clone() for kname, v in base.items()}
                pert[name].reshape(-1)[idx] += sign * h
                vals.append(float(f(pert)))
            fd = (vals[0] - vals[1]) / (2 * h)
            got = float(g.reshape(-1)[idx])
            assert abs(fd - got) <= 2e-5 * max(1.0, abs(got)), (name, idx, fd, got)
