"""Generates tests/golden/cagp_golden.npz from the CPU oracle (oracle/cagp_oracle.py).

The reference itself cannot run in this image (no jax/gpjax/cola), so these vectors pin the ORACLE
(and through it the CUDA path) against regressions; they are not outputs of the reference.
Run from the repo root:  python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cagp_oracle as O  # noqa: E402

CASES = [
    # name, family, n, d, k, ard, normalize
    ("rbf_n9_k4", "rbf", 9, 1, 4, False, True),
    ("rbf_n1037_k10_ard", "rbf", 1037, 3, 10, True, False),
    ("matern12_n300_k7", "matern12", 300, 2, 7, False, True),
    ("matern32_n700_k64", "matern32", 700, 8, 64, False, True),
    ("matern52_n523_k7_ard", "matern52", 523, 3, 7, True, False),
    ("rbf_readme_n1000_k10", "rbf", 1000, 1, 10, False, True),
]


def make_case(family, n, d, k, ard, normalize, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64) * 4.0
    y = torch.sin(x.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed + 100)
    s = torch.randn(n, generator=g, dtype=torch.float64)
    if normalize:
        s = O.normalize_by_blocks(s, k)
    ls = (0.7 + 0.5 * torch.rand(d, generator=g, dtype=torch.float64)) if ard else torch.tensor(0.9, dtype=torch.float64)
    return x, y, s, ls


def main():
    out = {}
    for name, family, n, d, k, ard, normalize in CASES:
        x, y, s, ls = make_case(family, n, d, k, ard, normalize)
        p = O.make_params(ls, 1.3, 0.2, 0.5, s)
        st = O.cagp_init(family, x, y, p, k)
        mean, var = O.cagp_predict(family, st, p)
        val, grads = O.elbo_and_grad_literal(family, x, y, p, k)
        out[f"{name}/x"] = x.numpy(); out[f"{name}/y"] = y.numpy(); out[f"{name}/s"] = s.numpy()
        out[f"{name}/lengthscale"] = ls.numpy()
        out[f"{name}/cov_prior_proj"] = st.cov_prior_proj.numpy()
        out[f"{name}/repr_weights_proj"] = st.repr_weights_proj.numpy()
        out[f"{name}/mean"] = mean.numpy(); out[f"{name}/var"] = var.numpy()
        out[f"{name}/kl"] = O.cagp_prior_kl(st).numpy()
        out[f"{name}/elbo"] = val.numpy()
        for gname, gval in grads.items():
            out[f"{name}/grad_{gname}"] = gval.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cagp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
