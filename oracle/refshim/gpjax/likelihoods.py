import numpy as np

from .parameters import NonNegativeReal, Parameter, _value


class Gaussian:
    def __init__(self, num_datapoints, obs_stddev=1.0):
        self.num_datapoints = num_datapoints
        self.obs_stddev = obs_stddev if isinstance(obs_stddev, Parameter) else NonNegativeReal(np.asarray(obs_stddev))

    def expected_log_likelihood(self, y, mean, variance):
        """E_q[log N(y | f, sigma^2)] for q(f) = N(mean, variance), summed over the output axis (GPJax likelihoods.py)."""
        s2 = np.squeeze(_value(self.obs_stddev)) ** 2
        sq_error = np.square(y - mean)
        val = np.sum(np.log(2.0 * np.pi) + np.log(s2) + (sq_error + variance) / s2, axis=1)
        return -0.5 * val
