from pymor.core.base import ImmutableObject
from pymor.operators.constructions import VectorOperator
from pymor.vectorarrays.interface import VectorArray


class StationaryModel(ImmutableObject):
    """operator(u, mu) = rhs(mu); output = output_functional(u, mu)."""

    def __init__(self, operator, rhs, output_functional=None, products=None, error_estimator=None,
                 visualizer=None, name=None):
        if isinstance(rhs, VectorArray):
            rhs = VectorOperator(rhs, name="rhs")
        self.__auto_init(locals())
        self.solution_space = operator.source
        self.linear = operator.linear and (output_functional is None or output_functional.linear)

    def solve(self, mu=None, return_error_estimate=False):
        U = self.operator.apply_inverse(self.rhs.as_range_array(mu), mu=mu)
        if return_error_estimate:
            return U, self.error_estimator.estimate_error(U, mu, self)
        return U

    def output(self, mu=None, solution=None):
        U = self.solve(mu) if solution is None else solution
        return self.output_functional.apply(U, mu=mu).to_numpy()

    def estimate_error(self, mu=None, solution=None):
        U = self.solve(mu) if solution is None else solution
        return self.error_estimator.estimate_error(U, mu, self)
