"""Operator base class: the algebra the reference's expressions use
(`emb @ op`, `.H`, `apply2`, `as_range_array`, scalar multiples)."""
import numpy as np

from pymor.core.base import ImmutableObject, abstractmethod
from pymor.parameters.base import ParametricObject


class Operator(ParametricObject, ImmutableObject):
    linear = False
    solver_options = None

    def apply(self, U, mu=None):
        raise NotImplementedError

    def apply_adjoint(self, V, mu=None):
        raise NotImplementedError

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        from pymor.operators.numpy import NumpyMatrixOperator
        assembled = self.assemble(mu)
        if assembled is not self and isinstance(assembled, NumpyMatrixOperator):
            return assembled.apply_inverse(V, least_squares=least_squares)
        raise NotImplementedError(f"apply_inverse of {type(self).__name__}")

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, least_squares=False):
        return self.H.apply_inverse(U, mu=mu, least_squares=least_squares)

    def apply2(self, V, U, mu=None):
        return V.inner(self.apply(U, mu=mu))

    def pairwise_apply2(self, V, U, mu=None):
        return V.pairwise_inner(self.apply(U, mu=mu))

    def assemble(self, mu=None):
        return self

    def as_range_array(self, mu=None):
        assert self.source.dim == 1
        return self.apply(self.source.ones(), mu=mu)

    def as_source_array(self, mu=None):
        assert self.range.dim == 1
        return self.apply_adjoint(self.range.ones(), mu=mu)

    def as_vector(self, mu=None):
        return self.as_range_array(mu) if self.source.dim == 1 else self.as_source_array(mu)

    @property
    def H(self):
        from pymor.operators.constructions import AdjointOperator
        return AdjointOperator(self)

    # -- algebra
    def __matmul__(self, other):
        from pymor.operators.constructions import ConcatenationOperator
        if isinstance(other, ConcatenationOperator):
            return NotImplemented
        if not isinstance(other, Operator):
            return NotImplemented
        return ConcatenationOperator((self, other))

    def __mul__(self, other):
        from pymor.operators.constructions import LincombOperator
        from pymor.parameters.functionals import ParameterFunctional
        assert isinstance(other, (int, float, complex, np.number, ParameterFunctional))
        return LincombOperator([self], [other])

    __rmul__ = __mul__

    def __add__(self, other):
        from pymor.operators.constructions import LincombOperator
        if isinstance(other, (int, float)) and other == 0:
            return self
        ops, cs = [], []
        for o in (self, other):
            if isinstance(o, LincombOperator):
                ops.extend(o.operators); cs.extend(o.coefficients)
            else:
                ops.append(o); cs.append(1.)
        return LincombOperator(ops, cs)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-1. * other)

    def __neg__(self):
        return self * (-1.)

    def __repr__(self):
        return f"{type(self).__name__}({getattr(self, 'range', None)} <- {getattr(self, 'source', None)})"
