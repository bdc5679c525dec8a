import numpy as np

from ..parameters import NonNegativeReal, Parameter, PositiveReal, _value
from .computations import DenseKernelComputation


class AbstractKernel:
    name = "AbstractKernel"

    def __init__(self, active_dims=None, lengthscale=1.0, variance=1.0, n_dims=None, compute_engine=None):
        self.lengthscale = lengthscale if isinstance(lengthscale, Parameter) else PositiveReal(np.asarray(lengthscale))
        self.variance = variance if isinstance(variance, Parameter) else NonNegativeReal(np.asarray(variance))
        self.compute_engine = compute_engine if compute_engine is not None else DenseKernelComputation()

    def gram(self, x):
        return self.compute_engine.gram(self, x)

    def cross_covariance(self, x, y):
        return self.compute_engine.cross_covariance(self, x, y)

    def __call__(self, x, y):
        return self.pairwise(np.asarray(x)[None, :], np.asarray(y)[None, :])[0, 0]

    def _scaled_sqdist(self, x, y):
        ls = _value(self.lengthscale)
        a, b = x / ls, y / ls
        d = a[:, None, :] - b[None, :, :]
        return np.sum(d * d, axis=-1)

    def pairwise(self, x, y):
        raise NotImplementedError
