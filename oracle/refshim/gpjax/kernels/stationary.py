"""gpjax.kernels.stationary.{rbf, matern12, matern32, matern52}: x / l, y / l, then the textbook forms;
Matern: tau = sqrt(max(|x - y|^2, 1e-36))."""
import numpy as np

from ..parameters import _value
from .base import AbstractKernel


class RBF(AbstractKernel):
    name = "RBF"

    def pairwise(self, x, y):
        return _value(self.variance) * np.exp(-0.5 * self._scaled_sqdist(x, y))


class _Matern(AbstractKernel):
    def _tau(self, x, y):
        return np.sqrt(np.maximum(self._scaled_sqdist(x, y), 1e-36))


class Matern12(_Matern):
    name = "Matern12"

    def pairwise(self, x, y):
        return _value(self.variance) * np.exp(-self._tau(x, y))


class Matern32(_Matern):
    name = "Matern32"

    def pairwise(self, x, y):
        tau = self._tau(x, y)
        return _value(self.variance) * (1.0 + np.sqrt(3.0) * tau) * np.exp(-np.sqrt(3.0) * tau)


class Matern52(_Matern):
    name = "Matern52"

    def pairwise(self, x, y):
        tau = self._tau(x, y)
        return _value(self.variance) * (1.0 + np.sqrt(5.0) * tau + 5.0 / 3.0 * np.square(tau)) * np.exp(-np.sqrt(5.0) * tau)
