import numpy as np
import paramax


class Parameter(paramax.AbstractUnwrappable):
    def __init__(self, value, tag=None):
        self.value = np.asarray(value)
        self.tag = tag

    def unwrap(self):
        return self.value

    @property
    def dtype(self):
        return self.value.dtype


class Real(Parameter):
    pass


class PositiveReal(Parameter):
    pass


class NonNegativeReal(Parameter):
    pass


class Static(Parameter):
    pass


def _value(p):
    return np.asarray(paramax.unwrap(p))
