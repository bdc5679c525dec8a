import numpy as np

from .parameters import Parameter, Real, _value


class AbstractMeanFunction:
    def __call__(self, x):
        raise NotImplementedError


class Constant(AbstractMeanFunction):
    def __init__(self, constant=0.0):
        self.constant = constant if isinstance(constant, Parameter) else Real(np.asarray(constant))

    def __call__(self, x):
        c = _value(self.constant)
        return np.ones((np.atleast_2d(x).shape[0], 1), dtype=np.float64) * c


class Zero(Constant):
    def __init__(self):
        super().__init__(0.0)
