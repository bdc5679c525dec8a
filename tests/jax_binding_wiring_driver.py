"""Driver (separate process) of tests/test_jax_binding_wiring.py: runs the REFERENCE'S model with the repository's JAX-side
drop-in (`jax_binding/cagpjax_b200_jax`: `B200KernelComputation`, `B200LazyKernel`, the `congruence_transform` overload)
under the dependency stand-ins of oracle/refshim, with every XLA FFI target replaced by a numpy emulation of the C-ABI
entry point's documented semantics (include/cagp_b200.h).  What this checks is the BINDING'S OWN LOGIC - target names
and registration, operand order and attributes, declared result shapes / dtypes, the k-major `Mt` layout, the variance
scaling, the plum overload that routes `S^T (K + sigma^2 I) S` to the fused projection, `LazyKernel` subclassing - by
requiring the reference model to give the same ELBO / predictions as with its own `LazyKernelComputation` (the committed
reference goldens).  It does not run any CUDA code (none can run where the reference sources are).

usage: jax_binding_wiring_driver.py <dir with libcagp_b200.so and a dummy libcagp_b200_xla.so> <golden.npz>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CAGP_B200_LIBDIR"] = sys.argv[1]
from oracle.run_reference import load_reference  # noqa: E402

ref = load_reference()
sys.path.insert(0, os.path.join(ROOT, "jax_binding"))
import jax  # noqa: E402  (the stand-in)

FAMILY_NAMES = {0: "rbf", 1: "matern12", 2: "matern32", 3: "matern52"}
CALLS = []


def unit_kernel(family, a, b):
    """unit-variance kernel of points scaled by 1 / lengthscale (rows of a, b)."""
    d = a[:, None, :] - b[None, :, :]
    sq = np.sum(d * d, axis=-1)
    if family == "rbf":
        return np.exp(-0.5 * sq)
    tau = np.sqrt(np.maximum(sq, 1e-36))
    if family == "matern12":
        return np.exp(-tau)
    if family == "matern32":
        return (1 + np.sqrt(3.0) * tau) * np.exp(-np.sqrt(3.0) * tau)
    return (1 + np.sqrt(5.0) * tau + 5.0 / 3.0 * tau * tau) * np.exp(-np.sqrt(5.0) * tau)


def block_of(n, k):
    bs = n // k
    idx = np.minimum(np.arange(n) // max(bs, 1), k - 1) if bs > 0 else np.full(n, k - 1)
    return idx


def emu_scale_inputs(X, ls, origin, *, family):
    CALLS.append("CagpScaleInputs")
    X = np.asarray(X)
    n, d = X.shape
    import ctypes

    dp = ctypes.CDLL(os.path.join(sys.argv[1], "libcagp_b200.so")).cagp_padded_dim(d)
    o = np.zeros(d, X.dtype) if np.asarray(origin).size == 0 else np.asarray(origin)
    ls = np.broadcast_to(np.asarray(ls), (d,))
    xs = np.zeros((dp, n), X.dtype)
    xs[:d] = ((X - o) / ls).T          # emulation's own scaling convention: kernels below take 1 / l-scaled points
    return xs


def emu_bounding_box(xs):
    CALLS.append("CagpBoundingBox")
    xs = np.asarray(xs)
    return np.concatenate([xs.min(axis=1), xs.max(axis=1)])


def emu_fwd_sym(xs, s, ls, bbox, *, family, d, k, a_begin, a_end):
    CALLS.append("CagpKsBlocksparseFwdSym")
    xs, s = np.asarray(xs), np.asarray(s)
    assert (a_begin, a_end) == (0, k) and np.asarray(bbox).shape == (2 * xs.shape[0],)
    P = xs[:d].T
    K = unit_kernel(FAMILY_NAMES[int(family)], P, P)
    n = P.shape[0]
    S = np.zeros((n, k), xs.dtype)
    S[np.arange(n), block_of(n, k)] = s
    return np.ascontiguousarray((K @ S).T)           # Mt: k-major


def emu_project_rows(Mt, s, yt, *, row_offset, n_total):
    CALLS.append("CagpProjectRows")
    Mt, s, yt = np.asarray(Mt), np.asarray(s), np.asarray(yt)
    k, m = Mt.shape
    assert row_offset == 0 and n_total == m
    S = np.zeros((m, k), Mt.dtype)
    S[np.arange(m), block_of(m, k)] = s
    return S.T @ Mt.T, Mt @ yt


def emu_gram(Mt):
    CALLS.append("CagpGramMtm")
    Mt = np.asarray(Mt)
    return Mt @ Mt.T, None   # second result: the workspace the caller sized with cagp_gram_ws_bytes


def emu_dense_fwd(xs1, xs2, V, ls, *, family, d):
    CALLS.append("CagpKsDenseFwd")
    xs1, xs2, V = np.asarray(xs1), np.asarray(xs2), np.asarray(V)
    return unit_kernel(FAMILY_NAMES[int(family)], xs1[:d].T, xs2[:d].T) @ V


# ---- reverse-mode entry points (semantics of include/cagp_b200.h, as cagpjax_b200/_fused.py uses them) -----------------
def dk_dq(family, sq):
    """d K' / d q for q = scaled squared distance."""
    if family == "rbf":
        return -0.5 * np.exp(-0.5 * sq)
    tau = np.sqrt(np.maximum(sq, 1e-36))
    if family == "matern12":
        return -np.exp(-tau) / (2 * tau)
    if family == "matern32":
        return -1.5 * np.exp(-np.sqrt(3.0) * tau)
    return -(5.0 / 6.0) * (1 + np.sqrt(5.0) * tau) * np.exp(-np.sqrt(5.0) * tau)


def emu_assemble_cotangent(Mt, W, Abar, cbar, s, yt, *, row_offset, n_total):
    CALLS.append("CagpAssembleCotangent")
    Mt, W, Abar, cbar, s, yt = map(np.asarray, (Mt, W, Abar, cbar, s, yt))
    k, m = Mt.shape
    blk = block_of(m, k)
    return W @ Mt + (Abar[blk, :] * s[:, None]).T + np.outer(cbar, yt)


def emu_bwd_sym(xs, s, ls, bbox, Gt, *, family, d, k, a_begin, a_end):
    CALLS.append("CagpKsBlocksparseBwdSym")
    xs, s, ls, Gt = map(np.asarray, (xs, s, ls, Gt))
    fam = FAMILY_NAMES[int(family)]
    P = xs[:d].T
    n = P.shape[0]
    blk = block_of(n, k)
    D = P[:, None, :] - P[None, :, :]
    sq = np.sum(D * D, axis=-1)
    K = unit_kernel(fam, P, P)
    Gj = Gt[blk, :].T                                   # G[i][blk(j)]
    s_bar = np.sum(K * Gj, axis=0)
    Wgt = dk_dq(fam, sq) * Gj * s[None, :]              # dF / dq_ij
    per_dim = np.einsum("ij,ijt->t", Wgt, D * D) * (-2.0) / np.broadcast_to(ls, (d,))   # q = sum_t D_t^2, D_t ~ 1 / l_t
    ls_bar = per_dim if ls.size == d else np.array([per_dim.sum()])       # ARD: one per dimension; isotropic: their sum
    return s_bar, ls_bar.astype(xs.dtype).reshape(ls.shape), None


def emu_project_rows_bwd(Mt, Abar, cbar, *, row_offset, n_total):
    CALLS.append("CagpProjectRowsBwd")
    Mt, Abar, cbar = map(np.asarray, (Mt, Abar, cbar))
    k, m = Mt.shape
    blk = block_of(m, k)
    return np.sum(Abar[blk, :].T * Mt, axis=0), cbar @ Mt


def check_project_stats_vjp(ops, family_id, X, ls, s, yt, k):
    """The binding's hand-paired reverse mode of (A', B', c') against central differences of its own forward."""
    rng = np.random.default_rng(5)
    Abar, Bbar, cbar = rng.normal(size=(k, k)), rng.normal(size=(k, k)), rng.normal(size=k)

    def f(ls_, s_, yt_):
        A, B, c = ops.project_stats(X, ls_, s_, yt_, family_id, k)
        return float(np.sum(Abar * A) + np.sum(Bbar * B) + cbar @ c)

    _, res = ops._project_stats_fwd(X, ls, s, yt, family_id, k)
    none, ls_bar, s_bar, yt_bar = ops._project_stats_bwd(family_id, k, res, (Abar, Bbar, cbar))
    assert none is None and np.shape(ls_bar) == np.shape(ls)

    def fd(arg, idx, h=1e-6):
        up, dn = [np.array(a, dtype=np.float64, copy=True) for a in (ls, s, yt)], [np.array(a, dtype=np.float64, copy=True) for a in (ls, s, yt)]
        up[arg][idx] += h
        dn[arg][idx] -= h
        return (f(*up) - f(*dn)) / (2 * h)

    worst = 0.0
    for arg, grad in ((0, np.atleast_1d(ls_bar)), (1, s_bar), (2, yt_bar)):
        size = np.atleast_1d([ls, s, yt][arg]).shape[0]
        for idx in sorted(set([0, size // 2, size - 1])):
            ref = fd(arg, (idx,) if np.ndim([ls, s, yt][arg]) else ())
            worst = max(worst, abs(grad[idx] - ref) / max(1.0, abs(ref)))
    return worst


def main():
    import cagpjax_b200_jax  # noqa: F401  registers the FFI targets and the congruence overload
    from cagpjax_b200_jax import B200KernelComputation, B200LazyKernel
    from cagpjax_b200_jax import ffi as F
    from cagpjax.models import ComputationAwareGP
    from cagpjax.policies import BlockSparsePolicy
    from cagpjax.solvers import Cholesky
    from gpjax.parameters import Real

    gpjax = ref["gpjax"]
    # every handler the binding registers must exist in the header-derived target table and the other way round
    assert set(jax.ffi._targets) == set(F.TARGETS), sorted(set(jax.ffi._targets) ^ set(F.TARGETS))
    for name, fn in {"CagpScaleInputs": emu_scale_inputs, "CagpBoundingBox": emu_bounding_box,
                     "CagpKsBlocksparseFwdSym": emu_fwd_sym, "CagpProjectRows": emu_project_rows, "CagpGramMtm": emu_gram,
                     "CagpKsDenseFwd": emu_dense_fwd, "CagpAssembleCotangent": emu_assemble_cotangent,
                     "CagpKsBlocksparseBwdSym": emu_bwd_sym, "CagpProjectRowsBwd": emu_project_rows_bwd}.items():
        jax.ffi.emulate(name, fn)

    G = np.load(sys.argv[2])
    kernels = {"rbf": gpjax.kernels.RBF, "matern12": gpjax.kernels.Matern12, "matern32": gpjax.kernels.Matern32,
               "matern52": gpjax.kernels.Matern52}
    worst = 0.0
    for name in ("rbf_overhang", "rbf_ard", "matern12", "matern32_d8", "matern52", "rbf_single_action"):
        c = {k[len(name) + 1:]: G[k] for k in G.files if k.startswith(name + "/")}
        fam, k = str(c["family"]), int(c["n_actions"])
        kern = kernels[fam](lengthscale=Real(c["hp_lengthscale"]), variance=Real(c["hp_variance"]), compute_engine=B200KernelComputation())
        prior = gpjax.gps.Prior(mean_function=gpjax.mean_functions.Constant(c["hp_mean_constant"]), kernel=kern)
        post = prior * gpjax.likelihoods.Gaussian(num_datapoints=c["x"].shape[0], obs_stddev=c["hp_obs_stddev"])
        model = ComputationAwareGP(post, BlockSparsePolicy(n_actions=k, nz_values=Real(c["actions"])), Cholesky(1e-6))
        CALLS.clear()
        st = model.init(gpjax.Dataset(X=c["x"], y=c["y"][:, None]))
        assert "CagpKsBlocksparseFwdSym" in CALLS and "CagpProjectRows" in CALLS, CALLS   # the fused route was taken
        assert "CagpKsDenseFwd" not in CALLS, CALLS                                       # ... and not the dense fallback
        elbo = float(model.elbo(st))
        ptest = model.predict(st, c["z"])
        errs = {"elbo": abs(elbo - float(c["elbo"])) / abs(float(c["elbo"])),
                "repr_weights": np.max(np.abs(np.asarray(st.repr_weights_proj) - c["repr_weights_proj"])) / np.max(np.abs(c["repr_weights_proj"])),
                "test_mean": np.max(np.abs(np.asarray(ptest.mean) - c["test_mean"])),
                "test_variance": np.max(np.abs(np.asarray(ptest.variance) - c["test_variance"]))}
        worst = max(worst, *errs.values())
        print(name, {kk: f"{v:.2e}" for kk, v in errs.items()}, flush=True)
        assert errs["elbo"] < 1e-11 and errs["repr_weights"] < 1e-8 and errs["test_mean"] < 1e-10 and errs["test_variance"] < 1e-10, errs
        # operator-level: products of the drop-in operator agree with the reference's LazyKernel
        from cagpjax.operators import LazyKernel

        dense_kern = kernels[fam](lengthscale=Real(c["hp_lengthscale"]), variance=Real(c["hp_variance"]))
        a, b = B200LazyKernel(dense_kern, c["z"], c["x"]), LazyKernel(dense_kern, c["z"], c["x"])
        V = np.random.default_rng(0).normal(size=(c["x"].shape[0], 3))
        W = np.random.default_rng(1).normal(size=(2, c["z"].shape[0]))
        assert np.max(np.abs(np.asarray(a @ V) - np.asarray(b @ V))) < 1e-12
        assert np.max(np.abs(np.asarray(W @ a) - np.asarray(W @ b))) < 1e-12
    # reverse mode: the hand-paired VJP of the fused statistics (custom_vjp rules called directly; no autodiff here)
    from cagpjax_b200_jax import ops

    for name, fam_id in (("rbf_overhang", 0), ("rbf_ard", 0), ("matern12", 1), ("matern32_d8", 2), ("matern52", 3)):
        c = {k[len(name) + 1:]: G[k] for k in G.files if k.startswith(name + "/")}
        w = check_project_stats_vjp(ops, fam_id, c["x"], np.atleast_1d(c["hp_lengthscale"]).astype(np.float64), c["actions"],
                                    c["y"] - float(c["hp_mean_constant"]), int(c["n_actions"]))
        print(name, f"project_stats VJP vs central differences: {w:.2e}", flush=True)
        assert w < 1e-6, w
    print(f"WIRING OK worst {worst:.2e}")


if __name__ == "__main__":
    main()
