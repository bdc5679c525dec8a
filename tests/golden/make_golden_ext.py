"""Generates tests/golden/cagp_golden_ext.npz from oracle/cagp_oracle_ext.py: pseudo-input, orthogonalised
pseudo-input and Lanczos actions with the Cholesky and the pseudo-inverse solver.

Like cagp_golden.npz these vectors pin the ORACLE (and through it the CUDA path) against regressions; they are
not outputs of the reference, which cannot run in this image.  Run from the repo root:
    python tests/golden/make_golden_ext.py
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cagp_oracle as O  # noqa: E402
from oracle import cagp_oracle_ext as E  # noqa: E402

HYPER = dict(variance=1.3, obs_stddev=0.2, mean_constant=0.5)
CASES = [
    # name, family, n, d, n_actions, policy, solver
    ("pseudo_rbf_n400_m9_chol", "rbf", 400, 2, 9, "pseudo", "cholesky"),
    ("pseudo_matern32_n523_m12_pinv", "matern32", 523, 3, 12, "pseudo", "pinv"),
    ("ortho_qr_rbf_n600_m16_chol", "rbf", 600, 2, 16, "ortho_qr", "cholesky"),
    ("ortho_mgs_matern52_n300_m8_chol", "matern52", 300, 3, 8, "ortho_mgs", "cholesky"),
    ("lanczos_rbf_n300_k5_chol", "rbf", 300, 1, 5, "lanczos", "cholesky"),
    ("lanczos_matern32_n700_k8_pinv", "matern32", 700, 3, 8, "lanczos", "pinv"),
]


def make_inputs(n, d, m, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 4.0
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    z = torch.rand(m, d, generator=torch.Generator().manual_seed(seed + 5), dtype=torch.float64) * 4.0
    zt = torch.rand(17, d, generator=torch.Generator().manual_seed(seed + 9), dtype=torch.float64) * 4.0
    return x, y, z, zt


def actions_for(policy, family, x, z, ls, var, sigma, k):
    if policy == "lanczos":
        n = x.shape[0]
        Khat = O.cross_covariance(family, x, x, ls, var) + sigma**2 * torch.eye(n, dtype=torch.float64)
        v0 = torch.randn(n, generator=torch.Generator().manual_seed(0), dtype=torch.float64)  # LanczosPolicy, key = 0
        return E.lanczos_eigh(Khat, k, v0)[1]
    S = E.pseudo_input_actions(family, x, z, ls, var)
    if policy == "ortho_qr":
        return E.orthogonalize_qr(S, 0)
    if policy == "ortho_mgs":
        return E.orthogonalize_mgs(S, 1)
    return S


def evaluate(family, n, d, k, policy, solver):
    x, y, z, zt = make_inputs(n, d, k)
    leaves = dict(lengthscale=torch.tensor(0.9, dtype=torch.float64), variance=torch.tensor(HYPER["variance"], dtype=torch.float64),
                  obs_stddev=torch.tensor(HYPER["obs_stddev"], dtype=torch.float64),
                  mean_constant=torch.tensor(HYPER["mean_constant"], dtype=torch.float64), pseudo_inputs=z.clone())
    q = {name: t.clone().requires_grad_(True) for name, t in leaves.items()}
    S = actions_for(policy, family, x, q["pseudo_inputs"], q["lengthscale"], q["variance"], q["obs_stddev"], k)
    kw = dict(solver="cholesky", jitter=1e-6) if solver == "cholesky" else dict(solver="pinv")
    elbo, mean, var = E.elbo_dense_actions(family, x, y, q["lengthscale"], q["variance"], q["obs_stddev"],
                                           q["mean_constant"], S, z=zt, **kw)
    grads = torch.autograd.grad(elbo, list(q.values()), allow_unused=True)
    out = {"x": x, "y": y, "z": z, "zt": zt, "elbo": elbo.detach(), "mean": mean.detach(), "var": var.detach()}
    for (name, t), g in zip(q.items(), grads):
        out[f"grad_{name}"] = torch.zeros_like(t) if g is None else g
    return out


def main():
    out = {}
    for name, family, n, d, k, policy, solver in CASES:
        for key, val in evaluate(family, n, d, k, policy, solver).items():
            out[f"{name}/{key}"] = val.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cagp_golden_ext.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
