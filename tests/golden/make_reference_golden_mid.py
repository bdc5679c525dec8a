"""Mid-size golden vectors from the REFERENCE'S OWN SOURCE (under oracle/refshim) at shapes that make the CUDA path take
its benchmarked kernel instances: float64 RBF, d = 3, isotropic lengthscale, long action blocks (the SINGLE-shape
fast-path kernels of csrc/ks_symx.cuh, as in BASELINE config 3), with and without an overhang block, plus a Matern-3/2
d = 8 case with short blocks (the SPREAD shape of config 2).  Gradients of the reference's ELBO by central differences
at a sample of action indices (all hyper-parameters).

    python tests/golden/make_reference_golden_mid.py       # writes tests/golden/reference_golden_mid.npz

Takes ~25 minutes on 16 cores (the reference evaluates the kernel row by row through the Python-loop stand-in of
`jax.lax.map`, and every finite-difference gradient entry costs two more ELBO evaluations), so unlike the small fixtures
it is not re-generated inside the test suite.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.run_reference import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_golden_mid.npz")
# (name, family, n, d, k, lengthscale, domain width)
CASES = [("c3_shape_even", "rbf", 4096, 3, 4, 1.0, 10.0), ("c3_shape_overhang", "rbf", 4100, 3, 3, 0.8, 10.0),
         ("c3_shape_wide", "rbf", 3000, 3, 3, 0.3, 10.0), ("c2_shape", "matern32", 1600, 8, 16, 1.5, 4.0)]


def main():
    ref = load_reference()
    gpjax = ref["gpjax"]
    from cagpjax.computations import LazyKernelComputation
    from cagpjax.models import ComputationAwareGP
    from cagpjax.policies import BlockSparsePolicy
    from cagpjax.policies.block_sparse import _normalize_by_blocks
    from cagpjax.solvers import Cholesky
    from gpjax.parameters import Real

    kernels = {"rbf": gpjax.kernels.RBF, "matern32": gpjax.kernels.Matern32}
    out = {"case_names": np.array([c[0] for c in CASES])}
    for name, family, n, d, k, ls0, width in CASES:
        rng = np.random.default_rng(sum(map(ord, name)))
        x = rng.uniform(0, width, size=(n, d))
        y = np.sin(x.sum(axis=1)) + 0.1 * rng.normal(size=n)
        z = rng.uniform(0, width, size=(64, d))
        hp = {"lengthscale": np.float64(ls0), "variance": np.float64(1.3), "obs_stddev": np.float64(0.2), "mean_constant": np.float64(0.5)}
        actions = np.asarray(_normalize_by_blocks(rng.normal(size=n), k))
        data = gpjax.Dataset(X=x, y=y[:, None])

        def build(hp_, actions_):
            kern = kernels[family](lengthscale=np.asarray(hp_["lengthscale"]), variance=np.asarray(hp_["variance"]),
                                   compute_engine=LazyKernelComputation(batch_size=512))
            prior = gpjax.gps.Prior(mean_function=gpjax.mean_functions.Constant(np.asarray(hp_["mean_constant"])), kernel=kern)
            post = prior * gpjax.likelihoods.Gaussian(num_datapoints=n, obs_stddev=np.asarray(hp_["obs_stddev"]))
            return ComputationAwareGP(post, BlockSparsePolicy(n_actions=k, nz_values=Real(actions_)), Cholesky(1e-6))

        def elbo_of(hp_, actions_):
            m = build(hp_, actions_)
            return float(m.elbo(m.init(data)))

        model = build(hp, actions)
        st = model.init(data)
        ptrain, ptest = model.predict(st), model.predict(st, z)
        rec = {"family": np.array(family), "n_actions": np.int64(k), "x": x, "y": y, "z": z, "actions": actions,
               **{f"hp_{kk}": np.asarray(v) for kk, v in hp.items()},
               "repr_weights_proj": np.asarray(st.repr_weights_proj), "elbo": np.float64(model.elbo(st)),
               "prior_kl": np.float64(model.prior_kl(st)), "train_mean": np.asarray(ptrain.mean),
               "train_variance": np.asarray(ptrain.variance), "test_mean": np.asarray(ptest.mean),
               "test_variance": np.asarray(ptest.variance)}
        for kk in hp:
            h = 1e-5 * max(1.0, abs(float(hp[kk])))
            up, dn = dict(hp), dict(hp)
            up[kk], dn[kk] = hp[kk] + h, hp[kk] - h
            rec[f"grad_{kk}"] = np.float64((elbo_of(up, actions) - elbo_of(dn, actions)) / (2 * h))
        idx = np.unique(np.concatenate([[0, n - 1, n // k - 1, n // k], rng.integers(0, n, size=8)]))
        ga = np.zeros(idx.shape[0])
        for j, i in enumerate(idx):
            h = 1e-5
            au, ad = actions.copy(), actions.copy()
            au[i] += h
            ad[i] -= h
            ga[j] = (elbo_of(hp, au) - elbo_of(hp, ad)) / (2 * h)
        rec["grad_actions_index"], rec["grad_actions_sample"] = idx, ga
        for kk, v in rec.items():
            out[f"{name}/{kk}"] = v
        print(f"{name}: elbo {float(rec['elbo']):.10f}", flush=True)
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else OUT, **out)


if __name__ == "__main__":
    main()
