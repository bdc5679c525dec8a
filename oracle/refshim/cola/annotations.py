"""Operator annotations: `PSD(op)` returns a copy of `op` (same class) carrying the annotation; `op.isa(PSD)`."""
import copy


class _AnnotationMeta(type):
    def __call__(cls, op):
        out = copy.copy(op)
        out.annotations = set(getattr(op, "annotations", ())) | {cls}
        return out


class Annotation(metaclass=_AnnotationMeta):
    pass


class SelfAdjoint(Annotation):
    pass


class PSD(SelfAdjoint):
    pass


class Unitary(Annotation):
    pass


class Stiefel(Annotation):
    pass
