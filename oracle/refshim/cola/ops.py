"""`cola.ops`: lazy linear operators on numpy arrays with CoLA's conventions (`_matmat`, `_rmatmat`, `.T`, `@`, `+`)."""
import numpy as np


class LinearOperator:
    __array_ufunc__ = None  # numpy defers `array @ op` to __rmatmul__

    def __init__(self, dtype, shape, matmat=None, annotations=()):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.annotations = set(annotations)
        self.device = "cpu"
        if matmat is not None:
            self._matmat = matmat

    # -- to be provided by subclasses
    def _matmat(self, X):
        raise NotImplementedError

    def _rmatmat(self, X):  # X @ A
        return self.T._matmat(X.T).T

    # -- algebra
    def to_dense(self):
        return self @ np.eye(self.shape[1], dtype=self.dtype)

    @property
    def T(self):
        return Transpose(self)

    def isa(self, annotation):
        return any(issubclass(a, annotation) for a in self.annotations)

    def __matmul__(self, X):
        if isinstance(X, LinearOperator):
            return Product(self, X)
        X = np.asarray(X)
        if X.ndim == 1:
            return self._matmat(X[:, None])[:, 0]
        return self._matmat(X)

    def __rmatmul__(self, X):
        X = np.asarray(X)
        if X.ndim == 1:
            return self._rmatmat(X[None, :])[0]
        return self._rmatmat(X)

    def __add__(self, other):
        return Sum(self, lazify(other))

    def __radd__(self, other):
        return Sum(lazify(other), self)

    def __neg__(self):
        return Product(ScalarMul(-1.0, (self.shape[0], self.shape[0]), self.dtype), self)

    def __sub__(self, other):
        return Sum(self, -lazify(other))

    def __mul__(self, c):
        return Product(ScalarMul(c, (self.shape[0], self.shape[0]), self.dtype), self)

    __rmul__ = __mul__


class Dense(LinearOperator):
    def __init__(self, A):
        A = np.asarray(A)
        super().__init__(A.dtype, A.shape)
        self.A = A

    def _matmat(self, X):
        return self.A @ X

    def _rmatmat(self, X):
        return X @ self.A

    def to_dense(self):
        return self.A

    @property
    def T(self):
        out = Dense(self.A.T)
        out.annotations = set(self.annotations)
        return out


class Diagonal(LinearOperator):
    def __init__(self, diag):
        diag = np.asarray(diag)
        super().__init__(diag.dtype, (diag.shape[0], diag.shape[0]))
        self.diag = diag

    def _matmat(self, X):
        return self.diag[:, None] * X

    def _rmatmat(self, X):
        return X * self.diag[None, :]

    @property
    def T(self):
        return self


class ScalarMul(LinearOperator):
    def __init__(self, c, shape, dtype=None, device=None):
        c = np.asarray(c)
        super().__init__(dtype if dtype is not None else c.dtype, shape)
        self.c = c.astype(self.dtype)[()] if c.ndim == 0 else c

    def _matmat(self, X):
        return self.c * X

    def _rmatmat(self, X):
        return X * self.c

    @property
    def T(self):
        return self


class Identity(LinearOperator):
    def __init__(self, shape, dtype):
        super().__init__(dtype, shape)

    def _matmat(self, X):
        return X

    def _rmatmat(self, X):
        return X

    @property
    def T(self):
        return self


def I_like(A):
    return Identity(A.shape, A.dtype)


class Triangular(LinearOperator):
    def __init__(self, A, lower=True):
        A = np.asarray(A)
        super().__init__(A.dtype, A.shape)
        self.A, self.lower = A, lower

    def _matmat(self, X):
        return self.A @ X

    def _rmatmat(self, X):
        return X @ self.A

    def to_dense(self):
        return self.A

    @property
    def T(self):
        return Triangular(self.A.T, lower=not self.lower)


class Transpose(LinearOperator):
    def __init__(self, A):
        super().__init__(A.dtype, (A.shape[1], A.shape[0]))
        self.A = A

    def _matmat(self, X):
        return self.A._rmatmat(X.T).T

    def _rmatmat(self, X):
        return self.A._matmat(X.T).T

    @property
    def T(self):
        return self.A


class Product(LinearOperator):
    def __init__(self, *Ms):
        for a, b in zip(Ms, Ms[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError(f"shape mismatch in product: {a.shape} @ {b.shape}")
        super().__init__(np.result_type(*[m.dtype for m in Ms]), (Ms[0].shape[0], Ms[-1].shape[1]))
        self.Ms = Ms

    def _matmat(self, X):
        for M in reversed(self.Ms):
            X = M @ X
        return X

    def _rmatmat(self, X):
        for M in self.Ms:
            X = X @ M
        return X

    @property
    def T(self):
        return Product(*[M.T for M in reversed(self.Ms)])


class Sum(LinearOperator):
    def __init__(self, *Ms):
        for m in Ms:
            if m.shape != Ms[0].shape:
                raise ValueError("shape mismatch in sum")
        super().__init__(np.result_type(*[m.dtype for m in Ms]), Ms[0].shape)
        self.Ms = Ms

    def _matmat(self, X):
        return sum(M @ X for M in self.Ms)

    def _rmatmat(self, X):
        return sum(X @ M for M in self.Ms)

    @property
    def T(self):
        return Sum(*[M.T for M in self.Ms])


def lazify(A):
    if isinstance(A, LinearOperator):
        return A
    return Dense(np.asarray(A))


def densify(A):
    return A.to_dense() if isinstance(A, LinearOperator) else np.asarray(A)
