"""Stand-in for `paramax`: parameter wrappers and `unwrap`."""
import numpy as np


class AbstractUnwrappable:
    def __class_getitem__(cls, item):
        return cls

    def unwrap(self):
        raise NotImplementedError


class NonTrainable(AbstractUnwrappable):
    def __init__(self, tree):
        self.tree = tree

    def unwrap(self):
        return self.tree


def non_trainable(tree):
    return NonTrainable(tree)


def unwrap(tree):
    while isinstance(tree, AbstractUnwrappable):
        tree = tree.unwrap()
    return tree
