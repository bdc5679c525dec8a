"""Golden vectors for the widened rows, again from the REFERENCE'S OWN SOURCE under oracle/refshim:
`linalg.orthogonalize` (QR / CGS / MGS, with and without re-orthogonalisation, full-rank and rank-deficient input),
`linalg.eigh` (dense `Eigh`), `PseudoInputPolicy`, `OrthogonalizationPolicy` and the model with those policies and with
either solver.  (`LanczosPolicy` runs matfree's recurrence, which is not vendored: not covered.)

    python tests/golden/make_reference_golden_ext.py       # writes tests/golden/reference_golden_ext.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.run_reference import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_golden_ext.npz")


def main():
    ref = load_reference()
    gpjax = ref["gpjax"]
    import cola
    from cagpjax.computations import LazyKernelComputation
    from cagpjax.linalg import OrthogonalizationMethod, eigh, orthogonalize
    from cagpjax.models import ComputationAwareGP
    from cagpjax.policies import OrthogonalizationPolicy, PseudoInputPolicy
    from cagpjax.solvers import Cholesky, PseudoInverse

    out = {}
    rng = np.random.default_rng(2024)

    # --- orthogonalize
    A_full = rng.normal(size=(40, 6))
    A_def = A_full.copy()
    A_def[:, 4] = A_def[:, 1] * 0.5 - A_def[:, 2]          # rank-deficient: column 4 in the span of 1, 2
    A_ill = A_full @ np.diag([1.0, 1e-3, 1.0, 1e-6, 1.0, 1e-2]) + 1e-9 * rng.normal(size=(40, 6))
    mats = {"full": A_full, "deficient": A_def, "ill": A_ill}
    for mname, A in mats.items():
        out[f"orth/{mname}/A"] = A
        for meth in OrthogonalizationMethod:
            for nre in (0, 1):
                Q = orthogonalize(A.copy(), method=meth, n_reortho=nre)
                out[f"orth/{mname}/{meth.value}/{nre}"] = np.asarray(Q)
    # operator form: annotations of the result, pass-through of already orthogonal operators
    Qop = orthogonalize(cola.ops.Dense(A_full), method=OrthogonalizationMethod.QR)
    out["orth/operator_qr_is_stiefel"] = np.array(Qop.isa(cola.Stiefel))
    again = orthogonalize(Qop, method=OrthogonalizationMethod.MGS, n_reortho=1)
    out["orth/stiefel_passes_through"] = np.array(again is Qop)

    # --- eigh (dense)
    B = rng.normal(size=(9, 9))
    B = B @ B.T + 0.1 * np.eye(9)
    res = eigh(cola.ops.Dense(B))
    out["eigh/A"] = B
    out["eigh/eigenvalues"] = np.asarray(res.eigenvalues)
    V = np.asarray(res.eigenvectors.to_dense())
    out["eigh/projector_first3"] = V[:, :3] @ V[:, :3].T   # sign-invariant
    out["eigh/is_unitary"] = np.array(res.eigenvectors.isa(cola.Unitary))

    # --- operators of the hot path: BlockDiagonalSparse, block normalisation, sparse x diagonal congruence, LazyKernel
    from cagpjax.linalg import congruence_transform
    from cagpjax.operators import BlockDiagonalSparse, LazyKernel
    from cagpjax.policies.block_sparse import _normalize_by_blocks

    for tag, n, k in (("overhang", 23, 5), ("even", 24, 6), ("single", 9, 1), ("per_point", 7, 7)):
        vals = rng.normal(size=n)
        op = BlockDiagonalSparse(vals, k)
        X = rng.normal(size=(k, 3))
        Y = rng.normal(size=(4, n))
        dvec = rng.uniform(0.5, 2.0, size=n)
        out[f"bds/{tag}/nz_values"] = vals
        out[f"bds/{tag}/n_blocks"] = np.int64(k)
        out[f"bds/{tag}/X"], out[f"bds/{tag}/Y"], out[f"bds/{tag}/dvec"] = X, Y, dvec
        out[f"bds/{tag}/matmat"] = np.asarray(op @ X)
        out[f"bds/{tag}/rmatmat"] = np.asarray(Y @ op)
        out[f"bds/{tag}/T_matmat"] = np.asarray(op.T @ Y.T)
        out[f"bds/{tag}/dense"] = np.asarray(op.to_dense())
        out[f"bds/{tag}/normalized"] = np.asarray(_normalize_by_blocks(vals, k))
        out[f"bds/{tag}/congruence_diag"] = np.asarray(congruence_transform(op, cola.ops.Diagonal(dvec)).diag)
        out[f"bds/{tag}/congruence_scalar"] = np.asarray(congruence_transform(op, cola.ops.ScalarMul(0.37, (n, n), np.float64)).diag)
    for fam in ("rbf", "matern12", "matern32", "matern52"):
        K = {"rbf": gpjax.kernels.RBF, "matern12": gpjax.kernels.Matern12, "matern32": gpjax.kernels.Matern32,
             "matern52": gpjax.kernels.Matern52}[fam](lengthscale=np.array([0.8, 1.7, 1.1]), variance=np.float64(1.4))
        x1, x2 = rng.uniform(-2, 2, size=(19, 3)), rng.uniform(-2, 2, size=(26, 3))
        R, Lm = rng.normal(size=(26, 4)), rng.normal(size=(3, 19))
        lk = LazyKernel(K, x1, x2, batch_size=6)
        out[f"lazy/{fam}/x1"], out[f"lazy/{fam}/x2"], out[f"lazy/{fam}/R"], out[f"lazy/{fam}/L"] = x1, x2, R, Lm
        out[f"lazy/{fam}/matmat"] = np.asarray(lk @ R)
        out[f"lazy/{fam}/rmatmat"] = np.asarray(Lm @ lk)
        out[f"lazy/{fam}/dense"] = np.asarray(lk.to_dense())
        out[f"lazy/{fam}/batch_rows_1GB"] = np.int64(LazyKernel(K, x1, x2).batch_size_row)

    # --- pseudo-input actions and the models built on them
    kernels = {"rbf": gpjax.kernels.RBF, "matern32": gpjax.kernels.Matern32, "matern52": gpjax.kernels.Matern52}
    cases = [("pi_rbf_qr_chol", "rbf", 42, 2, 6, "qr", 0, "cholesky"), ("pi_m32_mgs_chol", "matern32", 36, 3, 5, "mgs", 1, "cholesky"),
             ("pi_m52_cgs_pinv", "matern52", 38, 2, 5, "cgs", 1, "pseudoinverse"), ("pi_rbf_raw_pinv", "rbf", 30, 1, 4, None, 0, "pseudoinverse")]
    out["model_case_names"] = np.array([c[0] for c in cases])
    for name, family, n, d, M, meth, nre, solver in cases:
        r = np.random.default_rng(sum(map(ord, name)))
        x = r.uniform(-2, 2, size=(n, d))
        y = np.sin(1.5 * x[:, 0]) + 0.1 * r.normal(size=n)
        zs = r.uniform(-2, 2, size=(M, d))                  # pseudo-inputs
        zt = r.uniform(-2, 2, size=(7, d))                  # test inputs
        hp = dict(lengthscale=np.float64(r.uniform(0.7, 1.4)), variance=np.float64(r.uniform(0.7, 1.5)),
                  obs_stddev=np.float64(r.uniform(0.15, 0.4)), mean_constant=np.float64(r.uniform(-0.3, 0.3)))
        lazy_kernel = kernels[family](lengthscale=hp["lengthscale"], variance=hp["variance"], compute_engine=LazyKernelComputation(batch_size=5))
        dense_kernel = kernels[family](lengthscale=hp["lengthscale"], variance=hp["variance"])   # the policy's kernel: dense engine
        data = gpjax.Dataset(X=x, y=y[:, None])
        policy = PseudoInputPolicy(zs, data, dense_kernel)
        S_raw = np.asarray(policy.to_actions(None).to_dense())
        if meth is not None:
            policy = OrthogonalizationPolicy(policy, method=OrthogonalizationMethod(meth), n_reortho=nre)
        prior = gpjax.gps.Prior(mean_function=gpjax.mean_functions.Constant(hp["mean_constant"]), kernel=lazy_kernel)
        post = prior * gpjax.likelihoods.Gaussian(num_datapoints=n, obs_stddev=hp["obs_stddev"])
        model = ComputationAwareGP(post, policy, Cholesky(1e-6) if solver == "cholesky" else PseudoInverse())
        st = model.init(data)
        ptrain, ptest = model.predict(st), model.predict(st, zt)
        rec = {"family": np.array(family), "x": x, "y": y, "pseudo_inputs": zs, "z": zt, "method": np.array(meth or "none"),
               "n_reortho": np.int64(nre), "solver": np.array(solver), **{f"hp_{k}": np.asarray(v) for k, v in hp.items()},
               "actions_raw": S_raw, "actions": np.asarray(st.actions.to_dense()), "elbo": np.float64(model.elbo(st)),
               "prior_kl": np.float64(model.prior_kl(st)), "train_mean": np.asarray(ptrain.mean),
               "train_variance": np.asarray(ptrain.variance), "test_mean": np.asarray(ptest.mean),
               "test_variance": np.asarray(ptest.variance)}
        for k, v in rec.items():
            out[f"{name}/{k}"] = v
        print(f"{name}: elbo {float(rec['elbo']):.12f}", flush=True)
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else OUT, **out)


if __name__ == "__main__":
    main()
