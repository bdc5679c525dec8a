from .dataset import Dataset  # noqa: F401


class Prior:
    def __init__(self, mean_function, kernel, jitter=1e-6):
        self.mean_function, self.kernel, self.jitter = mean_function, kernel, jitter

    def __mul__(self, likelihood):
        return ConjugatePosterior(prior=self, likelihood=likelihood)

    __rmul__ = __mul__


class AbstractPosterior:
    def __init__(self, prior, likelihood, jitter=1e-6):
        self.prior, self.likelihood, self.jitter = prior, likelihood, jitter


class ConjugatePosterior(AbstractPosterior):
    pass
