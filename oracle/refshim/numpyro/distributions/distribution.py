class Distribution:
    def __init__(self, batch_shape=(), event_shape=(), *, validate_args=None):
        self._batch_shape = tuple(batch_shape)
        self._event_shape = tuple(event_shape)

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def event_shape(self):
        return self._event_shape

    @property
    def event_dim(self):
        return len(self._event_shape)

    def shape(self, sample_shape=()):
        return tuple(sample_shape) + self._batch_shape + self._event_shape


class MultivariateNormal(Distribution):
    """`MultivariateNormal(loc, covariance_matrix)` with scipy's log-density (used by the reference's tests only)."""

    def __init__(self, loc=0.0, covariance_matrix=None, precision_matrix=None, scale_tril=None, validate_args=None):
        import numpy as np

        self.loc = np.asarray(loc)
        if covariance_matrix is None:
            covariance_matrix = np.asarray(scale_tril) @ np.asarray(scale_tril).T
        self.covariance_matrix = np.asarray(covariance_matrix)
        super().__init__((), self.loc.shape)

    def log_prob(self, value):
        import numpy as np
        from scipy.stats import multivariate_normal

        out = multivariate_normal.logpdf(np.asarray(value, dtype=np.float64), np.asarray(self.loc, dtype=np.float64),
                                         np.asarray(self.covariance_matrix, dtype=np.float64), allow_singular=True)
        return np.asarray(out).astype(self.loc.dtype)
