"""The CUDA path against golden vectors produced by the REFERENCE'S OWN SOURCE.

`tests/golden/reference_golden.npz` holds inputs and the outputs of the unmodified reference package for them (made by
`tests/golden/make_reference_golden.py`; what that pins: oracle/refshim/README.md).  The fixture travels to the GPU box,
the reference does not.  Float64, tolerance 1e-10 (north star) on values; the reference's gradients are fourth-order
central finite differences of its own ELBO (accurate to ~3e-10 against exact reverse mode), asserted at 2e-8."""

import os

import numpy as np
import pytest
import torch

import cagpjax_b200 as cagp
from cagpjax_b200 import gpx
from cagpjax_b200.policies import DensePolicy
from cagpjax_b200.solvers import PseudoInverse
from tests.util import KERNELS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
CASES = [str(c) for c in G["case_names"]]


def case(name):
    pre = name + "/"
    return {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}


def dev(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(DEV)


def err(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a.reshape(b.shape) - b)) / max(1.0, np.max(np.abs(b))))


def build(c):
    fam, k = str(c["family"]), int(c["n_actions"])
    kernel = KERNELS[fam](lengthscale=dev(c["hp_lengthscale"]), variance=dev(c["hp_variance"]),
                          compute_engine=cagp.computations.LazyKernelComputation())
    prior = gpx.Prior(kernel=kernel, mean_function=gpx.Constant(dev(c["hp_mean_constant"])))
    lik = gpx.Gaussian(int(c["x"].shape[0]), obs_stddev=dev(c["hp_obs_stddev"]))
    if str(c["actions_kind"]) == "blocksparse":
        policy = cagp.BlockSparsePolicy(n_actions=k, nz_values=dev(c["actions"]))
    else:
        policy = DensePolicy(dev(c["actions"]))
    kw = {"solver": PseudoInverse()} if str(c["solver"]) == "pseudoinverse" else {}
    model = cagp.ComputationAwareGP(prior * lik, policy, **kw)
    data = gpx.Dataset(dev(c["x"]), dev(c["y"]).reshape(-1, 1))
    return model, data


@pytest.mark.parametrize("name", CASES)
def test_values_match_the_reference(name):
    c = case(name)
    model, data = build(c)
    tol = 1e-10 if str(c["solver"]) == "cholesky" else 1e-9   # the pseudo-inverse route goes through an eigendecomposition
    with torch.no_grad():
        st = model.init(data)
        assert err(st.residual_proj, c["residual_proj"]) < tol
        assert err(st.repr_weights_proj, c["repr_weights_proj"]) < 1e-8   # a solve, conditioned by S^T K^ S
        assert err(model.elbo(st), c["elbo"]) < tol
        assert err(model.prior_kl(st), c["prior_kl"]) < tol * 10
        assert err(model.variational_expectation(st), c["variational_expectation"]) < tol
        q = model.predict(st)
        assert err(q.mean, c["train_mean"]) < tol and err(q.variance, c["train_variance"]) < tol
        q = model.predict(st, dev(c["z"]))
        assert err(q.mean, c["test_mean"]) < tol and err(q.variance, c["test_variance"]) < tol
        assert err(q.scale.to_dense(), c["test_covariance"]) < tol
        if str(c["solver"]) == "cholesky":
            assert err(q.log_prob(q.mean + 0.05), c["test_log_prob_of_mean_plus"]) < 1e-8


@pytest.mark.parametrize("name", [n for n in CASES if str(G[n + "/solver"]) == "cholesky"])
def test_gradients_match_finite_differences_of_the_reference(name):
    c = case(name)
    model, data = build(c)
    val, grads = gpx.value_and_grad(lambda m, d: m.elbo(m.init(d)), model, data)
    assert err(val, c["elbo"]) < 1e-10
    names = {"lengthscale": "lengthscale", "variance": "variance", "obs_stddev": "obs_stddev", "constant": "mean_constant",
             "nz_values": "actions", "actions": "actions"}
    seen = 0
    for key, g in grads.items():
        leaf = names[key.split(".")[1]]
        assert err(g, c["grad_" + leaf]) < 2e-8, key
        seen += 1
    assert seen == 5


# ------------------------------------------------------------------------------------------------------------------
# widened rows: pseudo-input actions, orthogonalisation policies, both solvers
# (tests/golden/reference_golden_ext.npz from tests/golden/make_reference_golden_ext.py)
# ------------------------------------------------------------------------------------------------------------------
GX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_ext.npz"))


@pytest.mark.parametrize("name", [str(c) for c in GX["model_case_names"]])
def test_pseudo_input_models_match_the_reference(name):
    from cagpjax_b200.linalg import OrthogonalizationMethod
    from cagpjax_b200.policies import OrthogonalizationPolicy, PseudoInputPolicy

    pre = name + "/"
    c = {k[len(pre):]: GX[k] for k in GX.files if k.startswith(pre)}
    fam, meth = str(c["family"]), str(c["method"])
    kernel = KERNELS[fam](lengthscale=dev(c["hp_lengthscale"]), variance=dev(c["hp_variance"]),
                          compute_engine=cagp.computations.LazyKernelComputation())
    prior = gpx.Prior(kernel=kernel, mean_function=gpx.Constant(dev(c["hp_mean_constant"])))
    lik = gpx.Gaussian(int(c["x"].shape[0]), obs_stddev=dev(c["hp_obs_stddev"]))
    data = gpx.Dataset(dev(c["x"]), dev(c["y"]).reshape(-1, 1))
    policy = PseudoInputPolicy(dev(c["pseudo_inputs"]), data, kernel)
    with torch.no_grad():
        assert err(policy.to_actions(None).to_dense(), c["actions_raw"]) < 1e-12
    if meth != "none":
        policy = OrthogonalizationPolicy(policy, method=OrthogonalizationMethod(meth), n_reortho=int(c["n_reortho"]))
    kw = {"solver": PseudoInverse()} if str(c["solver"]) == "pseudoinverse" else {}
    model = cagp.ComputationAwareGP(prior * lik, policy, **kw)
    # raw pseudo-input actions: S^T K^ S has condition number ~1e9, every solve loses that many digits
    tol = 1e-8 if meth != "none" else 1e-5
    with torch.no_grad():
        st = model.init(data)
        if meth != "none":
            assert err(st.actions.to_dense(), c["actions"]) < 1e-8
        assert err(model.elbo(st), c["elbo"]) < tol
        q = model.predict(st)
        assert err(q.mean, c["train_mean"]) < tol and err(q.variance, c["train_variance"]) < tol
        q = model.predict(st, dev(c["z"]))
        assert err(q.mean, c["test_mean"]) < tol and err(q.variance, c["test_variance"]) < tol


@pytest.mark.parametrize("tag", ["overhang", "even", "single", "per_point"])
def test_block_operators_match_the_reference(tag):
    from cagpjax_b200 import ops
    from cagpjax_b200.linalg import congruence_transform
    from cagpjax_b200.operators import BlockDiagonalSparse
    from cagpjax_b200.policies.block_sparse import _normalize_by_blocks

    g = lambda k: GX[f"bds/{tag}/{k}"]
    vals, k = dev(g("nz_values")), int(g("n_blocks"))
    n = vals.shape[0]
    op = BlockDiagonalSparse(vals, k)
    assert err(op @ dev(g("X")), g("matmat")) < 1e-14
    assert err(dev(g("Y")) @ op, g("rmatmat")) < 1e-14
    assert err(op.T @ dev(g("Y")).T, g("T_matmat")) < 1e-14
    assert err(op.to_dense(), g("dense")) < 1e-15
    assert err(_normalize_by_blocks(vals, k), g("normalized")) < 1e-15
    assert err(congruence_transform(op, ops.Diagonal(dev(g("dvec")))).diag, g("congruence_diag")) < 1e-14
    assert err(congruence_transform(op, ops.ScalarMul(0.37, (n, n), torch.float64, device=DEV)).diag, g("congruence_scalar")) < 1e-14


@pytest.mark.parametrize("family", ["rbf", "matern12", "matern32", "matern52"])
def test_lazy_kernel_matches_the_reference(family):
    from cagpjax_b200.operators import LazyKernel

    g = lambda k: GX[f"lazy/{family}/{k}"]
    kernel = KERNELS[family](lengthscale=dev(np.array([0.8, 1.7, 1.1])), variance=dev(1.4))
    lk = LazyKernel(kernel, dev(g("x1")), dev(g("x2")), batch_size=6)
    assert err(lk @ dev(g("R")), g("matmat")) < 1e-12
    assert err(dev(g("L")) @ lk, g("rmatmat")) < 1e-12
    assert err(lk.to_dense(), g("dense")) < 1e-13
    assert LazyKernel(kernel, dev(g("x1")), dev(g("x2"))).batch_size_row == int(g("batch_rows_1GB"))


# ------------------------------------------------------------------------------------------------------------------
# mid-size cases from the reference's own source at the benchmarked kernel shapes (reference_golden_mid.npz)
# ------------------------------------------------------------------------------------------------------------------
GM = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_mid.npz"))
# expected kernel instance: level of the expanded-form fast path (csrc/ks_symx.cuh); 0 = general kernels
EXPECTED_LEVEL = {"c3_shape_even": 1, "c3_shape_overhang": 1, "c3_shape_wide": 2, "c2_shape": 0}


@pytest.mark.parametrize("name", [str(c) for c in GM["case_names"]])
def test_benchmark_shapes_match_the_reference(name):
    from cagpjax_b200 import _native as N

    pre = name + "/"
    c = {k[len(pre):]: GM[k] for k in GM.files if k.startswith(pre)}
    fam, k, d = str(c["family"]), int(c["n_actions"]), int(c["x"].shape[1])
    # which kernel instance this case exercises (the same scaling the model applies)
    X, ls = dev(c["x"]), dev(c["hp_lengthscale"]).reshape(1)
    lo, hi = torch.aminmax(X, dim=0)
    xs = N.scale_inputs(fam, X, ls, torch.lerp(lo, hi, 0.5))
    assert N.sym_fast_path(fam, xs, d, k, ls, N.bounding_box(xs)) == EXPECTED_LEVEL[name]
    assert N.sym_fast_path(fam, xs, d, k, ls, N.bounding_box(xs), backward=True) == EXPECTED_LEVEL[name]
    c["actions_kind"], c["solver"] = np.array("blocksparse"), np.array("cholesky")
    model, data = build(c)
    with torch.no_grad():
        st = model.init(data)
        assert err(st.repr_weights_proj, c["repr_weights_proj"]) < 1e-8
        assert err(model.elbo(st), c["elbo"]) < 1e-10
        assert err(model.prior_kl(st), c["prior_kl"]) < 1e-9
        q = model.predict(st)
        assert err(q.mean, c["train_mean"]) < 1e-10 and err(q.variance, c["train_variance"]) < 1e-10
        q = model.predict(st, dev(c["z"]))
        assert err(q.mean, c["test_mean"]) < 1e-10 and err(q.variance, c["test_variance"]) < 1e-10
    val, grads = gpx.value_and_grad(lambda m, dd: m.elbo(m.init(dd)), model, data)
    assert err(val, c["elbo"]) < 1e-10
    for key, g in grads.items():
        leaf = {"lengthscale": "lengthscale", "variance": "variance", "obs_stddev": "obs_stddev", "constant": "mean_constant",
                "nz_values": "actions"}[key.split(".")[1]]
        if leaf == "actions":
            got = g.detach().cpu().numpy()[c["grad_actions_index"]]
            assert np.max(np.abs(got - c["grad_actions_sample"])) <= 2e-6 * max(1.0, np.max(np.abs(c["grad_actions_sample"]))), key
        else:
            assert err(g, c["grad_" + leaf]) < 2e-6, key
