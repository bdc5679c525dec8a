"""Golden vectors produced by the REFERENCE'S OWN SOURCE (unmodified /root/reference/src/cagpjax), executed in float64
numpy under the dependency stand-ins of oracle/refshim (this image has no jax / cola / gpjax; see oracle/refshim/README.md
for what that does and does not pin).

    python tests/golden/make_reference_golden.py            # writes tests/golden/reference_golden.npz

Every case stores its inputs (so nothing depends on a random-number stream) and what the reference computed from them:
`ComputationAwareGP.init` (projected residual, representer weights, the projected prior covariance and noise
covariance), `elbo`, `prior_kl`, `variational_expectation`, `predict` at the training and at test inputs (mean,
variance, full covariance), `GaussianDistribution.log_prob`, and the gradient of the reference's `elbo` with respect
to every hyper-parameter and every action value by fourth-order central finite differences of the reference itself (no
autodiff in the stand-ins).  Runs in a fresh process only (the stand-ins shadow `jaxtyping`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.run_reference import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")

# (name, family, n, d, k, ard, n_test, solver, actions)
CASES = [
    ("rbf_overhang", "rbf", 37, 2, 5, False, 11, "cholesky", "blocksparse"),
    ("rbf_even", "rbf", 48, 3, 6, False, 9, "cholesky", "blocksparse"),
    ("rbf_ard", "rbf", 45, 3, 7, True, 8, "cholesky", "blocksparse"),
    ("rbf_1d_readme", "rbf", 60, 1, 10, False, 13, "cholesky", "blocksparse"),
    ("rbf_single_action", "rbf", 23, 2, 1, False, 5, "cholesky", "blocksparse"),
    ("rbf_block_per_point", "rbf", 16, 2, 16, False, 6, "cholesky", "blocksparse"),
    ("matern12", "matern12", 33, 2, 4, False, 7, "cholesky", "blocksparse"),
    ("matern32_d8", "matern32", 50, 8, 6, False, 7, "cholesky", "blocksparse"),
    ("matern32_ard", "matern32", 40, 4, 5, True, 7, "cholesky", "blocksparse"),
    ("matern52", "matern52", 41, 3, 8, False, 10, "cholesky", "blocksparse"),
    ("rbf_dense_actions", "rbf", 36, 3, 5, False, 8, "cholesky", "dense"),
    ("matern52_dense_actions", "matern52", 30, 2, 4, False, 6, "cholesky", "dense"),
    ("rbf_pseudoinverse", "rbf", 34, 2, 5, False, 7, "pseudoinverse", "blocksparse"),
]


def main():
    ref = load_reference()
    gpjax = ref["gpjax"]
    import cola
    import jax
    from cagpjax.computations import LazyKernelComputation
    from cagpjax.models import ComputationAwareGP
    from cagpjax.policies import BlockSparsePolicy
    from cagpjax.policies.base import AbstractBatchLinearSolverPolicy
    from cagpjax.solvers import Cholesky, PseudoInverse
    from gpjax.parameters import Real

    class DenseActions(AbstractBatchLinearSolverPolicy):
        """A policy that returns a given dense action matrix (exercises the reference's dense-S code path)."""

        S: np.ndarray

        def __init__(self, S):
            self.S = S
            self.n_actions = S.shape[1]

        def to_actions(self, A, *, key=None):
            return cola.ops.Dense(self.S)

    kernels = {"rbf": gpjax.kernels.RBF, "matern12": gpjax.kernels.Matern12, "matern32": gpjax.kernels.Matern32,
               "matern52": gpjax.kernels.Matern52}

    def build(family, hp, actions, k, kind, solver):
        kern = kernels[family](lengthscale=np.asarray(hp["lengthscale"]), variance=np.asarray(hp["variance"]),
                               compute_engine=LazyKernelComputation(batch_size=7))
        prior = gpjax.gps.Prior(mean_function=gpjax.mean_functions.Constant(np.asarray(hp["mean_constant"])), kernel=kern)
        post = prior * gpjax.likelihoods.Gaussian(num_datapoints=actions.shape[0], obs_stddev=np.asarray(hp["obs_stddev"]))
        policy = BlockSparsePolicy(n_actions=k, nz_values=Real(actions)) if kind == "blocksparse" else DenseActions(actions)
        slv = Cholesky(1e-6) if solver == "cholesky" else PseudoInverse()
        return ComputationAwareGP(post, policy, slv)

    out = {"case_names": np.array([c[0] for c in CASES])}
    for name, family, n, d, k, ard, m, solver, kind in CASES:
        rng = np.random.default_rng(sum(map(ord, name)))  # deterministic per case
        x = rng.uniform(-2.5, 2.5, size=(n, d))
        y = np.sin(x[:, 0]) * np.cos(0.7 * x[:, -1]) + 0.1 * rng.normal(size=n)
        z = rng.uniform(-2.5, 2.5, size=(m, d))
        hp = {"lengthscale": rng.uniform(0.6, 1.8, size=d) if ard else np.float64(rng.uniform(0.7, 1.5)),
              "variance": np.float64(rng.uniform(0.6, 1.6)), "obs_stddev": np.float64(rng.uniform(0.1, 0.4)),
              "mean_constant": np.float64(rng.uniform(-0.5, 0.5))}
        if kind == "blocksparse":
            # the reference's own initialiser (block-normalised samples), on the stand-in's random stream
            actions = np.asarray(BlockSparsePolicy.from_random(jax.random.key(7), n, k, dtype=np.float64).nz_values)
            actions = actions * rng.uniform(0.5, 1.5, size=n)  # and away from unit block norms, as after training
        else:
            actions = rng.normal(size=(n, k))
        data = gpjax.Dataset(X=x, y=y[:, None])

        def elbo_of(hp_, actions_):
            mdl = build(family, hp_, actions_, k, kind, solver)
            return float(mdl.elbo(mdl.init(data)))

        model = build(family, hp, actions, k, kind, solver)
        st = model.init(data)
        pred_train = model.predict(st)
        pred_test = model.predict(st, z)
        rec = {
            "family": np.array(family), "n_actions": np.int64(k), "solver": np.array(solver), "actions_kind": np.array(kind),
            "x": x, "y": y, "z": z, "actions": actions,
            **{f"hp_{kk}": np.asarray(v) for kk, v in hp.items()},
            "residual_proj": np.asarray(st.residual_proj), "repr_weights_proj": np.asarray(st.repr_weights_proj),
            "obs_cov_proj": np.asarray(st.obs_cov_proj.to_dense()),
            "actions_dense": np.asarray(st.actions.to_dense()),
            "elbo": np.float64(model.elbo(st)), "prior_kl": np.float64(model.prior_kl(st)),
            "variational_expectation": np.asarray(model.variational_expectation(st)),
            "train_mean": np.asarray(pred_train.mean), "train_variance": np.asarray(pred_train.variance),
            "test_mean": np.asarray(pred_test.mean), "test_variance": np.asarray(pred_test.variance),
            "test_covariance": np.asarray(pred_test.covariance().to_dense()),
            "test_log_prob_of_mean_plus": np.float64(pred_test.log_prob(np.asarray(pred_test.mean) + 0.05)),
        }
        if solver == "cholesky":
            rec["cov_prior_proj_chol"] = np.asarray(st.cov_prior_proj_state.to_dense())
        # gradient of the reference's ELBO by FOURTH-ORDER central differences of the reference (five-point stencil,
        # h = 1e-4 relative: against exact reverse mode this is accurate to ~1e-10, where the two-point stencil at
        # h = 1e-5 stops at ~1e-8; tools: see the probe quoted in DESIGN.md section 5)
        def d4(fun, h):
            return (-fun(2 * h) + 8 * fun(h) - 8 * fun(-h) + fun(-2 * h)) / (12 * h)

        for kk in hp:
            base = np.asarray(hp[kk], dtype=np.float64)
            g = np.zeros_like(base)
            for idx in np.ndindex(base.shape) if base.ndim else [()]:
                def shifted(dx, kk=kk, idx=idx, base=base):
                    b = base.copy()
                    b[idx] += dx
                    hp_ = dict(hp)
                    hp_[kk] = b
                    return elbo_of(hp_, actions)

                g[idx] = d4(shifted, 1e-4 * max(1.0, abs(float(base[idx]))))
            rec[f"grad_{kk}"] = g
        ga = np.zeros_like(actions)
        for idx in np.ndindex(actions.shape):
            def shifted(dx, idx=idx):
                a_ = actions.copy()
                a_[idx] += dx
                return elbo_of(hp, a_)

            ga[idx] = d4(shifted, 1e-4 * max(1.0, abs(float(actions[idx]))))
        rec["grad_actions"] = ga
        for kk, v in rec.items():
            out[f"{name}/{kk}"] = v
        print(f"{name}: elbo {float(rec['elbo']):.12f}  kl {float(rec['prior_kl']):.6f}", flush=True)
    if len(sys.argv) > 1:
        np.savez_compressed(sys.argv[1], **out)
    else:
        np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
