"""`gpjax.gps`: Prior, ConjugatePosterior (+ their exact predictive distributions, used by the reference's tests)."""
import numpy as np

from .dataset import Dataset  # noqa: F401


def _dense(op):
    """Dense matrix of whatever a compute engine's gram / cross_covariance returned."""
    if hasattr(op, "as_matrix"):
        return np.asarray(op.as_matrix())
    if hasattr(op, "to_dense"):
        return np.asarray(op.to_dense())
    return np.asarray(op)


class Prior:
    def __init__(self, mean_function=None, kernel=None, jitter=1e-6):
        self.mean_function, self.kernel, self.jitter = mean_function, kernel, jitter

    def __mul__(self, likelihood):
        return ConjugatePosterior(prior=self, likelihood=likelihood)

    __rmul__ = __mul__

    def predict(self, test_inputs):
        from .distributions import GaussianDistribution, _psd

        x = np.atleast_2d(test_inputs)
        mean = np.squeeze(self.mean_function(x), -1)
        Kxx = _dense(self.kernel.gram(x))
        return GaussianDistribution(mean, _psd(Kxx + np.eye(x.shape[0], dtype=Kxx.dtype) * self.jitter))


class AbstractPosterior:
    def __init__(self, prior=None, likelihood=None, jitter=1e-6):
        self.prior, self.likelihood, self.jitter = prior, likelihood, jitter


class ConjugatePosterior(AbstractPosterior):
    def predict(self, test_inputs, train_data):
        """Exact GP regression posterior (textbook; GPJax adds `jitter` to K_xx and to the predictive covariance)."""
        import paramax

        from .distributions import GaussianDistribution, _psd

        x, y, t = np.atleast_2d(train_data.X), np.asarray(train_data.y), np.atleast_2d(test_inputs)
        s2 = np.asarray(paramax.unwrap(self.likelihood.obs_stddev)) ** 2
        mx, mt = self.prior.mean_function(x), self.prior.mean_function(t)
        Kxx = _dense(self.prior.kernel.gram(x))
        Sigma = Kxx + np.eye(x.shape[0], dtype=Kxx.dtype) * (self.jitter + s2)
        Ktt = _dense(self.prior.kernel.gram(t))
        Kxt = _dense(self.prior.kernel.cross_covariance(x, t))
        A = np.linalg.solve(Sigma, Kxt)
        mean = mt + A.T @ (y.reshape(-1, 1) - mx)
        cov = Ktt - Kxt.T @ A + np.eye(t.shape[0], dtype=Kxx.dtype) * self.jitter
        return GaussianDistribution(np.squeeze(mean, -1), _psd(cov))
