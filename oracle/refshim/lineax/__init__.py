"""Stand-in for `lineax`: the operator classes and tags the reference's interop.py converts between."""
import functools

import numpy as np


class AbstractLinearOperator:
    def mv(self, vector):
        raise NotImplementedError

    def as_matrix(self):
        raise NotImplementedError

    @property
    def T(self):
        return self.transpose()

    def in_size(self):
        return self.as_matrix().shape[1]

    def out_size(self):
        return self.as_matrix().shape[0]

    def __add__(self, other):
        return AddLinearOperator(self, other)


class _Tag:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


positive_semidefinite_tag = _Tag("positive_semidefinite_tag")
symmetric_tag = _Tag("symmetric_tag")


class TaggedLinearOperator(AbstractLinearOperator):
    def __init__(self, operator, tags):
        self.operator = operator
        self.tags = frozenset(tags) if isinstance(tags, (set, frozenset, list, tuple)) else frozenset([tags])

    def mv(self, vector):
        return self.operator.mv(vector)

    def as_matrix(self):
        return self.operator.as_matrix()

    def transpose(self):
        return TaggedLinearOperator(self.operator.transpose(), self.tags)


class MatrixLinearOperator(AbstractLinearOperator):
    def __init__(self, matrix, tags=()):
        self.matrix = np.asarray(matrix)
        self.tags = frozenset(tags) if isinstance(tags, (set, frozenset, list, tuple)) else frozenset([tags])

    def mv(self, vector):
        return self.matrix @ vector

    def as_matrix(self):
        return self.matrix

    def transpose(self):
        return MatrixLinearOperator(self.matrix.T, self.tags)


class DiagonalLinearOperator(AbstractLinearOperator):
    def __init__(self, diagonal):
        self.diagonal = np.asarray(diagonal)

    def mv(self, vector):
        return self.diagonal * vector

    def as_matrix(self):
        return np.diag(self.diagonal)

    def transpose(self):
        return self


class IdentityLinearOperator(AbstractLinearOperator):
    def __init__(self, input_structure, output_structure=None):
        self.input_structure = input_structure

    def mv(self, vector):
        return vector

    def as_matrix(self):
        n = self.input_structure.shape[0]
        return np.eye(n, dtype=self.input_structure.dtype)

    def transpose(self):
        return self


class AddLinearOperator(AbstractLinearOperator):
    def __init__(self, operator1, operator2):
        self.operator1, self.operator2 = operator1, operator2

    def mv(self, vector):
        return self.operator1.mv(vector) + self.operator2.mv(vector)

    def as_matrix(self):
        return self.operator1.as_matrix() + self.operator2.as_matrix()

    def transpose(self):
        return AddLinearOperator(self.operator1.transpose(), self.operator2.transpose())


@functools.singledispatch
def is_symmetric(operator):
    return False


@functools.singledispatch
def is_positive_semidefinite(operator):
    return False


@is_symmetric.register(TaggedLinearOperator)
def _(operator):
    return bool({symmetric_tag, positive_semidefinite_tag} & operator.tags) or is_symmetric(operator.operator)


@is_positive_semidefinite.register(TaggedLinearOperator)
def _(operator):
    return positive_semidefinite_tag in operator.tags or is_positive_semidefinite(operator.operator)
