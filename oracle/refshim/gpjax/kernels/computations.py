import lineax as lx
import numpy as np


class AbstractKernelComputation:
    def gram(self, kernel, x):
        Kxx = self.cross_covariance(kernel, x, x)
        return lx.TaggedLinearOperator(lx.MatrixLinearOperator(Kxx), lx.positive_semidefinite_tag)

    def cross_covariance(self, kernel, x, y):
        raise NotImplementedError

    def diagonal(self, kernel, inputs):
        return np.array([kernel(x, x) for x in inputs])


class DenseKernelComputation(AbstractKernelComputation):
    def cross_covariance(self, kernel, x, y):
        """GPJax: vmap(lambda x: vmap(lambda y: kernel(x, y))(y))(x); evaluated here by the kernel's pairwise form."""
        return kernel.pairwise(np.atleast_2d(x), np.atleast_2d(y))
