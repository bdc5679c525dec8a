"""`gpjax.distributions.GaussianDistribution` (used by the reference's TESTS as the explicit-KL anchor)."""
import lineax as lx
import numpy as np


class GaussianDistribution:
    def __init__(self, loc=None, scale=None):
        self.loc = np.asarray(loc)
        self.scale = scale

    @property
    def mean(self):
        return self.loc

    def covariance(self):
        return self.scale.as_matrix()

    @property
    def variance(self):
        return np.diag(self.scale.as_matrix())

    def kl_divergence(self, other):
        """KL(self || other) between two multivariate normals."""
        Sq, Sp = np.asarray(self.scale.as_matrix()), np.asarray(other.scale.as_matrix())
        n = Sq.shape[0]
        Lp = np.linalg.cholesky(Sp)
        Lq = np.linalg.cholesky(Sq)
        diff = np.asarray(other.loc) - np.asarray(self.loc)
        from scipy.linalg import solve_triangular

        M = solve_triangular(Lp, Lq, lower=True)
        z = solve_triangular(Lp, diff, lower=True)
        logdet = 2.0 * (np.sum(np.log(np.diag(Lp))) - np.sum(np.log(np.diag(Lq))))
        return 0.5 * (np.sum(M * M) + z @ z - n + logdet)


def _psd(matrix):
    return lx.TaggedLinearOperator(lx.MatrixLinearOperator(matrix), lx.positive_semidefinite_tag)
