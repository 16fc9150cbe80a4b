"""Identity / Inverse / Concatenation / Lincomb / Zero / VectorArray / Adjoint operators."""
import numpy as np

from pymor.operators.interface import Operator
from pymor.parameters.functionals import ParameterFunctional
from pymor.vectorarrays.numpy import NumpyVectorSpace


class IdentityOperator(Operator):
    linear = True

    def __init__(self, space, name=None):
        self.__auto_init(locals())
        self.source = self.range = space

    def apply(self, U, mu=None):
        assert U in self.source
        return U.copy()

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        return V.copy()

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        return V.copy()

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, least_squares=False):
        return U.copy()


class ZeroOperator(Operator):
    linear = True

    def __init__(self, range, source, name=None):
        self.__auto_init(locals())

    def apply(self, U, mu=None):
        return self.range.zeros(len(U))

    def apply_adjoint(self, V, mu=None):
        return self.source.zeros(len(V))


class ConstantOperator(Operator):
    linear = False

    def __init__(self, value, source, name=None):
        self.__auto_init(locals())
        self.range = value.space

    def apply(self, U, mu=None):
        return self.value[[0] * len(U)].copy()


class InverseOperator(Operator):
    def __init__(self, operator, name=None):
        self.__auto_init(locals())
        self.source, self.range, self.linear = operator.range, operator.source, operator.linear

    def apply(self, U, mu=None):
        return self.operator.apply_inverse(U, mu=mu)

    def apply_adjoint(self, V, mu=None):
        return self.operator.apply_inverse_adjoint(V, mu=mu)

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        return self.operator.apply(V, mu=mu)

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, least_squares=False):
        return self.operator.apply_adjoint(U, mu=mu)


class AdjointOperator(Operator):
    linear = True

    def __init__(self, operator, source_product=None, range_product=None, name=None):
        assert operator.linear
        self.__auto_init(locals())
        self.source, self.range = operator.range, operator.source

    def apply(self, U, mu=None):
        return self.operator.apply_adjoint(U, mu=mu)

    def apply_adjoint(self, V, mu=None):
        return self.operator.apply(V, mu=mu)

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        return self.operator.apply_inverse_adjoint(V, mu=mu, least_squares=least_squares)

    @property
    def H(self):
        return self.operator


class ConcatenationOperator(Operator):
    """operators[0] o operators[1] o ... : applied right to left."""

    def __init__(self, operators, solver_options=None, name=None):
        operators = tuple(operators)
        assert all(operators[i].source == operators[i + 1].range for i in range(len(operators) - 1))
        self.__auto_init(locals())
        self.source, self.range = operators[-1].source, operators[0].range
        self.linear = all(op.linear for op in operators)

    def apply(self, U, mu=None):
        for op in self.operators[::-1]:
            U = op.apply(U, mu=mu)
        return U

    def apply_adjoint(self, V, mu=None):
        for op in self.operators:
            V = op.apply_adjoint(V, mu=mu)
        return V

    def __matmul__(self, other):
        if isinstance(other, ConcatenationOperator):
            return self.with_(operators=self.operators + other.operators)
        if isinstance(other, Operator):
            return self.with_(operators=self.operators + (other,))
        return NotImplemented

    def __rmatmul__(self, other):
        if isinstance(other, Operator):
            return self.with_(operators=(other,) + self.operators)
        return NotImplemented


class LincombOperator(Operator):
    def __init__(self, operators, coefficients, solver_options=None, name=None):
        operators, coefficients = tuple(operators), tuple(coefficients)
        assert len(operators) == len(coefficients) > 0
        assert all(op.source == operators[0].source and op.range == operators[0].range for op in operators)
        self.__auto_init(locals())
        self.source, self.range = operators[0].source, operators[0].range
        self.linear = all(op.linear for op in operators)

    @property
    def parametric(self):
        return any(isinstance(c, ParameterFunctional) for c in self.coefficients) or any(op.parametric for op in self.operators)

    def evaluate_coefficients(self, mu):
        return np.array([c.evaluate(mu) if isinstance(c, ParameterFunctional) else c for c in self.coefficients])

    def apply(self, U, mu=None):
        coeffs = self.evaluate_coefficients(mu)
        R = self.operators[0].apply(U, mu=mu)
        R.scal(coeffs[0])
        for op, c in zip(self.operators[1:], coeffs[1:]):
            R.axpy(c, op.apply(U, mu=mu))
        return R

    def apply_adjoint(self, V, mu=None):
        coeffs = self.evaluate_coefficients(mu).conj()
        R = self.operators[0].apply_adjoint(V, mu=mu)
        R.scal(coeffs[0])
        for op, c in zip(self.operators[1:], coeffs[1:]):
            R.axpy(c, op.apply_adjoint(V, mu=mu))
        return R

    def assemble(self, mu=None):
        from pymor.operators.numpy import NumpyMatrixOperator
        ops = [op.assemble(mu) for op in self.operators]
        coeffs = self.evaluate_coefficients(mu)
        if all(isinstance(op, NumpyMatrixOperator) for op in ops):
            matrix = ops[0].matrix * coeffs[0]
            for op, c in zip(ops[1:], coeffs[1:]):
                matrix = matrix + op.matrix * c
            return NumpyMatrixOperator(matrix, source_id=self.source.id, range_id=self.range.id)
        return LincombOperator(ops, coeffs)

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        op = self.assemble(mu)
        if isinstance(op, LincombOperator):
            raise NotImplementedError
        return op.apply_inverse(V, least_squares=least_squares)

    def __mul__(self, other):
        return self.with_(coefficients=tuple(c * other for c in self.coefficients))

    __rmul__ = __mul__


class VectorArrayOperator(Operator):
    """adjoint=False: coefficients -> linear combination of `array`; adjoint=True: inner
    products with `array`."""
    linear = True

    def __init__(self, array, adjoint=False, space_id=None, name=None):
        self.__auto_init(locals())
        if adjoint:
            self.source, self.range = array.space, NumpyVectorSpace(len(array), space_id)
        else:
            self.source, self.range = NumpyVectorSpace(len(array), space_id), array.space

    def apply(self, U, mu=None):
        assert U in self.source
        if not self.adjoint:
            return self.array.lincomb(U.to_numpy())
        return self.range.make_array(self.array.inner(U).T)

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        if not self.adjoint:
            return self.source.make_array(self.array.inner(V).T)
        return self.array.lincomb(V.to_numpy())

    def as_range_array(self, mu=None):
        return self.array.copy() if not self.adjoint else super().as_range_array(mu)

    def as_source_array(self, mu=None):
        return self.array.copy() if self.adjoint else super().as_source_array(mu)


class VectorOperator(VectorArrayOperator):
    def __init__(self, vector, name=None):
        assert len(vector) == 1
        super().__init__(vector, adjoint=False, name=name)
        self.vector = vector


class VectorFunctional(VectorArrayOperator):
    def __init__(self, vector, product=None, name=None):
        assert len(vector) == 1
        if product is not None:
            vector = product.apply(vector)
        super().__init__(vector, adjoint=True, name=name)
        self.vector, self.product = vector, None
