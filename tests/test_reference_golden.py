"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCE (CPU).

`tests/golden/reference_golden.npz` was written by `tests/golden/make_reference_golden.py`, which imports the unmodified
`/root/reference/src/cagpjax` under numpy stand-ins for its absent dependencies (oracle/refshim/README.md says what
that pins: the cagpjax package's own logic; not jax / CoLA / GPJax themselves).  Here:

  * every quantity the reference computed (init, ELBO, KL, variational expectation, predictive mean / variance /
    covariance at train and test inputs, log_prob) is compared with the oracle's restatement at 1e-12;
  * the oracle's reverse-mode gradients are compared with fourth-order central finite differences OF THE REFERENCE
    (2e-9; measured 3e-10 at worst - the differences are the limit, not the oracle);
  * when the reference sources are present (this container; not the GPU box) the generator is re-run in a fresh
    process and must reproduce the committed fixture, so the fixture provably is the script's output.
"""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cagp_oracle as O
from oracle import cagp_oracle_ext as OX
from oracle.run_reference import reference_available

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_golden.npz")
G = np.load(GOLDEN)
CASES = [str(c) for c in G["case_names"]]


def case(name):
    pre = name + "/"
    return {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}


def t(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b)))


BLOCKSPARSE_CHOL = [c for c in CASES if str(G[c + "/actions_kind"]) == "blocksparse" and str(G[c + "/solver"]) == "cholesky"]
DENSE = [c for c in CASES if str(G[c + "/actions_kind"]) == "dense"]
PINV = [c for c in CASES if str(G[c + "/solver"]) == "pseudoinverse"]


def test_fixture_covers_the_path():
    fams = {str(G[c + "/family"]) for c in CASES}
    assert fams == {"rbf", "matern12", "matern32", "matern52"}
    assert BLOCKSPARSE_CHOL and DENSE and PINV
    assert any(int(G[c + "/x"].shape[0]) % int(G[c + "/n_actions"]) for c in BLOCKSPARSE_CHOL)      # an overhang block
    assert any(G[c + "/hp_lengthscale"].ndim == 1 for c in CASES)                                    # ARD


@pytest.mark.parametrize("name", BLOCKSPARSE_CHOL)
def test_oracle_matches_reference_blocksparse(name):
    c = case(name)
    fam, k = str(c["family"]), int(c["n_actions"])
    p = O.make_params(c["hp_lengthscale"], c["hp_variance"], c["hp_obs_stddev"], c["hp_mean_constant"], c["actions"])
    x, y, z = t(c["x"]), t(c["y"]), t(c["z"])
    st = O.cagp_init(fam, x, y, p, k, batch_size=7)
    assert close(st.S.numpy(), c["actions_dense"], 1e-15)
    assert close(st.residual_proj.numpy(), c["residual_proj"], 1e-13)
    assert close(st.repr_weights_proj.numpy(), c["repr_weights_proj"], 1e-10)  # a solve: conditioned by cov_prior_proj
    assert close(torch.diag(st.obs_cov_proj).numpy() if st.obs_cov_proj.dim() == 1 else st.obs_cov_proj.numpy(), c["obs_cov_proj"], 1e-14)
    assert close(st.L.numpy(), c["cov_prior_proj_chol"], 1e-12)
    assert close(float(O.cagp_elbo(fam, st, p)), float(c["elbo"]), 1e-12)
    assert close(float(O.cagp_prior_kl(st)), float(c["prior_kl"]), 1e-11)
    assert close(O.cagp_variational_expectation(fam, st, p).numpy(), c["variational_expectation"], 1e-12)
    m, v = O.cagp_predict(fam, st, p)
    assert close(m.numpy(), c["train_mean"], 1e-12) and close(v.numpy(), c["train_variance"], 1e-12)
    m, v, cov = O.cagp_predict(fam, st, p, z, full_cov=True)
    assert close(m.numpy(), c["test_mean"], 1e-12) and close(v.numpy(), c["test_variance"], 1e-12)
    assert close(cov.numpy(), c["test_covariance"], 1e-12)
    # structured (independent) formulation of the oracle against the reference's ELBO as well
    es = O.elbo_structured(fam, c["x"], c["y"], c["hp_lengthscale"], float(c["hp_variance"]), float(c["hp_obs_stddev"]),
                           float(c["hp_mean_constant"]), c["actions"], k)
    assert close(float(es), float(c["elbo"]), 1e-11)


@pytest.mark.parametrize("name", BLOCKSPARSE_CHOL)
def test_oracle_gradients_match_reference_finite_differences(name):
    c = case(name)
    fam, k = str(c["family"]), int(c["n_actions"])
    p = O.make_params(c["hp_lengthscale"], c["hp_variance"], c["hp_obs_stddev"], c["hp_mean_constant"], c["actions"])
    val, grads = O.elbo_and_grad_literal(fam, t(c["x"]), t(c["y"]), p, k)
    assert close(float(val), float(c["elbo"]), 1e-12)
    for leaf in ("lengthscale", "variance", "obs_stddev", "mean_constant", "actions"):
        ref = np.asarray(c["grad_" + leaf], dtype=np.float64)
        got = grads[leaf].numpy().reshape(ref.shape)
        assert np.max(np.abs(got - ref)) <= 2e-9 * max(1.0, np.max(np.abs(ref))), leaf


@pytest.mark.parametrize("name", DENSE)
def test_oracle_gradients_match_reference_finite_differences_dense_actions(name):
    c = case(name)
    fam = str(c["family"])
    p = O.make_params(c["hp_lengthscale"], c["hp_variance"], c["hp_obs_stddev"], c["hp_mean_constant"], c["actions"])
    val, grads = O.elbo_and_grad_literal(fam, t(c["x"]), t(c["y"]), p, None)
    assert close(float(val), float(c["elbo"]), 1e-12)
    for leaf in ("lengthscale", "variance", "obs_stddev", "mean_constant", "actions"):
        ref = np.asarray(c["grad_" + leaf], dtype=np.float64)
        got = grads[leaf].numpy().reshape(ref.shape)
        assert np.max(np.abs(got - ref)) <= 2e-9 * max(1.0, np.max(np.abs(ref))), leaf


@pytest.mark.parametrize("name", DENSE + PINV)
def test_oracle_matches_reference_dense_actions_and_pseudoinverse(name):
    c = case(name)
    fam = str(c["family"])
    solver = "pinv" if str(c["solver"]) == "pseudoinverse" else "cholesky"
    S = t(c["actions_dense"])
    args = (fam, t(c["x"]), t(c["y"]), t(c["hp_lengthscale"]), t(c["hp_variance"]), t(c["hp_obs_stddev"]), t(c["hp_mean_constant"]), S)
    elbo, m, v = OX.elbo_dense_actions(*args, solver=solver)
    assert close(float(elbo), float(c["elbo"]), 1e-10)
    assert close(m.numpy(), c["train_mean"], 1e-10) and close(v.numpy(), c["train_variance"], 1e-10)
    elbo, m, v = OX.elbo_dense_actions(*args, solver=solver, z=t(c["z"]))
    assert close(m.numpy(), c["test_mean"], 1e-10) and close(v.numpy(), c["test_variance"], 1e-10)


@pytest.mark.skipif(not reference_available(), reason="reference sources are not on this machine")
def test_fixture_is_what_the_generator_writes(tmp_path):
    out = str(tmp_path / "regen.npz")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_golden.py"), out], check=True, env=env,
                   capture_output=True, timeout=600)
    R = np.load(out)
    assert sorted(R.files) == sorted(G.files)
    for key in G.files:
        a, b = G[key], R[key]
        if a.dtype.kind in "US":
            assert np.array_equal(a, b), key
        elif key.split("/")[-1].startswith("grad_"):
            assert np.allclose(a, b, rtol=0, atol=2e-8 * max(1.0, np.max(np.abs(a)))), key   # ~1e-13-noisy ELBOs over h = 1e-4
        else:
            assert np.allclose(a, b, rtol=1e-11, atol=1e-12), key


# ------------------------------------------------------------------------------------------------------------------
# widened rows: tests/golden/reference_golden_ext.npz from tests/golden/make_reference_golden_ext.py
# ------------------------------------------------------------------------------------------------------------------
GX = np.load(os.path.join(HERE, "golden", "reference_golden_ext.npz"))
ORTH = {"qr": OX.orthogonalize_qr, "cgs": OX.orthogonalize_cgs, "mgs": OX.orthogonalize_mgs}


@pytest.mark.parametrize("matrix", ["full", "deficient", "ill"])
@pytest.mark.parametrize("method", ["qr", "cgs", "mgs"])
@pytest.mark.parametrize("n_reortho", [0, 1])
def test_oracle_orthogonalize_matches_reference(matrix, method, n_reortho):
    A = t(GX[f"orth/{matrix}/A"])
    ref = GX[f"orth/{matrix}/{method}/{n_reortho}"]
    got = ORTH[method](A, n_reortho).numpy()
    if matrix == "deficient" and method in ("qr", "mgs"):
        # rank-deficient input: QR and MGS (rtol = 0) normalise the rounding noise left in the dependent column (4) to
        # unit length, and column 5 is orthogonalised against that arbitrary direction; columns 0-3 are well defined
        # (CGS zeroes the dependent column through its rtol and is compared in full below)
        assert close(got[:, :4], ref[:, :4], 1e-11)
        if method == "qr":
            assert close(got.T @ got, np.eye(6), 1e-12) and close(ref.T @ ref, np.eye(6), 1e-12)
    else:
        tol = 1e-6 if matrix == "ill" else 1e-11   # ill-conditioned input amplifies last-bit differences of the sums
        assert close(got, ref, tol)
    assert bool(GX["orth/operator_qr_is_stiefel"]) and bool(GX["orth/stiefel_passes_through"])


def test_reference_eigh_is_a_spectral_decomposition():
    B = GX["eigh/A"]
    w, V = np.linalg.eigh(B)
    assert close(w, GX["eigh/eigenvalues"], 1e-12)
    assert close(V[:, :3] @ V[:, :3].T, GX["eigh/projector_first3"], 1e-10)
    assert bool(GX["eigh/is_unitary"])


@pytest.mark.parametrize("name", [str(c) for c in GX["model_case_names"]])
def test_oracle_matches_reference_pseudo_input_models(name):
    pre = name + "/"
    c = {k[len(pre):]: GX[k] for k in GX.files if k.startswith(pre)}
    fam = str(c["family"])
    S_raw = OX.pseudo_input_actions(fam, t(c["x"]), t(c["pseudo_inputs"]), t(c["hp_lengthscale"]), t(c["hp_variance"]))
    assert close(S_raw.numpy(), c["actions_raw"], 1e-13)
    meth = str(c["method"])
    S = S_raw if meth == "none" else ORTH[meth](S_raw, int(c["n_reortho"]))
    assert close(S.numpy(), c["actions"], 1e-9)
    solver = "pinv" if str(c["solver"]) == "pseudoinverse" else "cholesky"
    args = (fam, t(c["x"]), t(c["y"]), t(c["hp_lengthscale"]), t(c["hp_variance"]), t(c["hp_obs_stddev"]), t(c["hp_mean_constant"]), S)
    elbo, m, v = OX.elbo_dense_actions(*args, solver=solver)
    tol = 1e-9 if meth != "none" else 1e-6   # raw pseudo-input actions: S^T K^ S is ill-conditioned (cond ~ 1e9), see DESIGN 3b
    assert close(float(elbo), float(c["elbo"]), tol)
    assert close(m.numpy(), c["train_mean"], tol) and close(v.numpy(), c["train_variance"], tol)
    elbo, m, v = OX.elbo_dense_actions(*args, solver=solver, z=t(c["z"]))
    assert close(m.numpy(), c["test_mean"], tol) and close(v.numpy(), c["test_variance"], tol)


@pytest.mark.skipif(not reference_available(), reason="reference sources are not on this machine")
def test_ext_fixture_is_what_the_generator_writes(tmp_path):
    out = str(tmp_path / "regen_ext.npz")
    subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_golden_ext.py"), out], check=True,
                   capture_output=True, timeout=600)
    R = np.load(out)
    assert sorted(R.files) == sorted(GX.files)
    for key in GX.files:
        a, b = GX[key], R[key]
        if a.dtype.kind in "USb":
            assert np.array_equal(a, b), key
        else:
            assert np.allclose(a, b, rtol=1e-9, atol=1e-11), key


@pytest.mark.parametrize("tag", ["overhang", "even", "single", "per_point"])
def test_oracle_block_operators_match_reference(tag):
    g = lambda k: GX[f"bds/{tag}/{k}"]
    vals, k = t(g("nz_values")), int(g("n_blocks"))
    assert close(O.bds_matmat(vals, k, t(g("X"))).numpy(), g("matmat"), 1e-15)
    assert close(O.bds_rmatmat(vals, k, t(g("Y"))).numpy(), g("rmatmat"), 1e-14)
    assert close(g("T_matmat"), g("rmatmat").T, 1e-15)                       # the reference's own consistency
    assert close(O.bds_dense(vals, k).numpy(), g("dense"), 1e-15)
    assert close(O.normalize_by_blocks(vals, k).numpy(), g("normalized"), 1e-15)
    assert close(O.congruence_bds_diag(vals, k, t(g("dvec"))).numpy(), g("congruence_diag"), 1e-14)
    assert close(O.congruence_bds_diag(vals, k, 0.37).numpy(), g("congruence_scalar"), 1e-14)


@pytest.mark.parametrize("family", ["rbf", "matern12", "matern32", "matern52"])
def test_oracle_lazy_kernel_matches_reference(family):
    g = lambda k: GX[f"lazy/{family}/{k}"]
    ls, var = t(np.array([0.8, 1.7, 1.1])), t(1.4)
    x1, x2 = t(g("x1")), t(g("x2"))
    assert close(O.cross_covariance(family, x1, x2, ls, var).numpy(), g("dense"), 1e-14)
    assert close(O.lazy_kernel_matmat(family, x1, x2, ls, var, t(g("R")), batch_size=6).numpy(), g("matmat"), 1e-13)
    assert close(O.lazy_kernel_rmatmat(family, x1, x2, ls, var, t(g("L")), batch_size=6).numpy(), g("rmatmat"), 1e-13)
    rows, _ = O.lazy_batch_sizes((19, 26), 8)
    assert rows == int(g("batch_rows_1GB"))


@pytest.mark.parametrize("tag", ["overhang", "even", "single", "per_point"])
def test_host_layer_block_operators_match_reference(tag):
    """The product's host-side operators that need no kernel (pure torch; they run on CPU as well) against the
    reference's `BlockDiagonalSparse`, `_normalize_by_blocks` and the sparse x diagonal `congruence_transform`."""
    from cagpjax_b200 import ops
    from cagpjax_b200.linalg import congruence_transform
    from cagpjax_b200.operators import BlockDiagonalSparse
    from cagpjax_b200.policies.block_sparse import _normalize_by_blocks

    g = lambda k: GX[f"bds/{tag}/{k}"]
    vals, k = t(g("nz_values")), int(g("n_blocks"))
    n = vals.shape[0]
    op = BlockDiagonalSparse(vals, k)
    assert close((op @ t(g("X"))).numpy(), g("matmat"), 1e-15)
    assert close((t(g("Y")) @ op).numpy(), g("rmatmat"), 1e-14)
    assert close((op.T @ t(g("Y")).T).numpy(), g("T_matmat"), 1e-14)
    assert close(op.to_dense().numpy(), g("dense"), 1e-15)
    assert close(_normalize_by_blocks(vals, k).numpy(), g("normalized"), 1e-15)
    assert close(congruence_transform(op, ops.Diagonal(t(g("dvec")))).diag.numpy(), g("congruence_diag"), 1e-14)
    assert close(congruence_transform(op, ops.ScalarMul(0.37, (n, n), torch.float64)).diag.numpy(), g("congruence_scalar"), 1e-14)


# ------------------------------------------------------------------------------------------------------------------
# mid-size cases at the benchmarked kernel shapes: tests/golden/reference_golden_mid.npz
# (tests/golden/make_reference_golden_mid.py; 25 minutes of the reference's Python row loop, so not re-generated here)
# ------------------------------------------------------------------------------------------------------------------
GM = np.load(os.path.join(HERE, "golden", "reference_golden_mid.npz"))


@pytest.mark.parametrize("name", [str(c) for c in GM["case_names"]])
def test_oracle_matches_reference_at_benchmark_shapes(name):
    pre = name + "/"
    c = {k[len(pre):]: GM[k] for k in GM.files if k.startswith(pre)}
    fam, k = str(c["family"]), int(c["n_actions"])
    p = O.make_params(c["hp_lengthscale"], c["hp_variance"], c["hp_obs_stddev"], c["hp_mean_constant"], c["actions"])
    st = O.cagp_init(fam, t(c["x"]), t(c["y"]), p, k)
    assert close(st.repr_weights_proj.numpy(), c["repr_weights_proj"], 1e-9)
    assert close(float(O.cagp_elbo(fam, st, p)), float(c["elbo"]), 1e-12)
    assert close(float(O.cagp_prior_kl(st)), float(c["prior_kl"]), 1e-10)
    m, v = O.cagp_predict(fam, st, p)
    assert close(m.numpy(), c["train_mean"], 1e-11) and close(v.numpy(), c["train_variance"], 1e-11)
    m, v = O.cagp_predict(fam, st, p, t(c["z"]))
    assert close(m.numpy(), c["test_mean"], 1e-11) and close(v.numpy(), c["test_variance"], 1e-11)
    es = O.elbo_structured(fam, c["x"], c["y"], c["hp_lengthscale"], float(c["hp_variance"]), float(c["hp_obs_stddev"]),
                           float(c["hp_mean_constant"]), c["actions"], k)
    assert close(float(es), float(c["elbo"]), 1e-11)
    val, grads = O.elbo_and_grad_literal(fam, t(c["x"]), t(c["y"]), p, k)
    for leaf in ("lengthscale", "variance", "obs_stddev", "mean_constant"):
        ref = float(c["grad_" + leaf])
        assert abs(float(grads[leaf]) - ref) <= 2e-6 * max(1.0, abs(ref)), leaf
    ga = grads["actions"].numpy()[c["grad_actions_index"]]
    assert np.max(np.abs(ga - c["grad_actions_sample"])) <= 2e-6 * max(1.0, np.max(np.abs(c["grad_actions_sample"])))
