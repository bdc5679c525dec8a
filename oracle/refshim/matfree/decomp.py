"""Stand-in for `matfree.decomp.tridiag_sym`: symmetric Lanczos tridiagonalisation with full re-orthogonalisation.

RESTATED from the textbook recurrence (matfree is not vendored): `tridiag_sym(k, materialize=True, reortho="full")`
returns `estimate(matvec, v0) -> (Q, H, residual, 1 / |v0|)` with `Q` (n, k) orthonormal columns spanning the Krylov
space of `v0` and `H = Q^T A Q` (k, k) symmetric tridiagonal.  Rounding-level details (order of the re-orthogonalisation
passes) may differ from matfree; nothing that depends on them is part of the pinned goldens."""
import numpy as np


def tridiag_sym(num_matvecs, /, *, materialize=True, reortho="full", custom_vjp=True):
    if not materialize:
        raise NotImplementedError("only materialize=True is part of the stand-in")

    def estimate(matvec, vec, *params):
        from jax.numpy import _wrap

        v0 = np.asarray(vec)
        n, k = v0.shape[0], int(num_matvecs)
        if k > n:
            raise ValueError("num_matvecs exceeds the dimension")
        dtype = v0.dtype
        Q = np.zeros((n, k), dtype=dtype)
        alpha = np.zeros(k, dtype=dtype)
        beta = np.zeros(max(k - 1, 0), dtype=dtype)
        length = np.linalg.norm(v0)
        q = v0 / length
        w = np.zeros_like(q)
        for j in range(k):
            if reortho == "full" and j > 0:
                q = q - Q[:, :j] @ (Q[:, :j].T @ q)
                q = q / np.linalg.norm(q)
            Q[:, j] = q
            w = np.asarray(matvec(_wrap(q), *params)).astype(dtype)
            alpha[j] = q @ w
            w = w - alpha[j] * q - (beta[j - 1] * Q[:, j - 1] if j > 0 else 0.0)
            if reortho == "full":
                w = w - Q[:, :j + 1] @ (Q[:, :j + 1].T @ w)
            b = np.linalg.norm(w)
            if j < k - 1:
                beta[j] = b
                q = w / b if b > 0 else w
        H = np.diag(alpha) + np.diag(beta, 1) + np.diag(beta, -1)
        return _wrap(Q), _wrap(H), _wrap(w), 1.0 / length

    return estimate
