from pymor.core.base import ImmutableObject
from pymor.parameters.base import ParametricObject


class ParameterFunctional(ParametricObject, ImmutableObject):
    _own_parametric = True

    def evaluate(self, mu=None):
        raise NotImplementedError

    def __call__(self, mu=None):
        return self.evaluate(mu)

    def __mul__(self, other):
        return ProductParameterFunctional([self, other])

    __rmul__ = __mul__


class ProjectionParameterFunctional(ParameterFunctional):
    def __init__(self, parameter, size=1, index=None, name=None):
        if index is None and size == 1:
            index = 0
        self.__auto_init(locals())

    def evaluate(self, mu=None):
        return float(mu[self.parameter][self.index])


class GenericParameterFunctional(ParameterFunctional):
    def __init__(self, mapping, parameters=None, name=None):
        self.__auto_init(locals())

    def evaluate(self, mu=None):
        return self.mapping(mu)


class ProductParameterFunctional(ParameterFunctional):
    def __init__(self, factors, name=None):
        self.__auto_init(locals())

    def evaluate(self, mu=None):
        out = 1.0
        for f in self.factors:
            out = out * (f.evaluate(mu) if isinstance(f, ParameterFunctional) else f)
        return out
