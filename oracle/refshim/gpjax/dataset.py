import numpy as np


class Dataset:
    def __init__(self, X=None, y=None):
        self.X = None if X is None else np.asarray(X)
        self.y = None if y is None else np.asarray(y)

    @property
    def n(self):
        return self.X.shape[0]
