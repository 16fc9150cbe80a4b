from pymor.algorithms.gram_schmidt import gram_schmidt
from pymor.algorithms.projection import project
from pymor.core.base import BasicObject
from pymor.operators.constructions import LincombOperator
from pymor.operators.interface import Operator


class ResidualOperator(Operator):
    """U -> operator(U, mu) - rhs(mu)."""

    def __init__(self, operator, rhs, name=None):
        self.__auto_init(locals())
        self.source, self.range, self.linear = operator.source, operator.range, operator.linear
        self.rhs_vector = rhs if rhs is not None and not rhs.parametric else None

    def apply(self, U, mu=None):
        V = self.operator.apply(U, mu=mu)
        if self.rhs is not None:
            F = self.rhs.as_range_array(mu)
            if len(V) > 1:
                V -= F[[0] * len(V)]
            else:
                V -= F
        return V


class ResidualReductor(BasicObject):
    """Residual restricted to span(RB), measured through Riesz representatives in `product`
    (simplified: the image basis is the plain union of the affine-term images, orthonormalised)."""

    def __init__(self, RB, operator, rhs=None, product=None, riesz_representatives=False):
        self.__auto_init(locals())

    def reduce(self):
        def terms(op):
            return list(op.operators) if isinstance(op, LincombOperator) else [op]
        images = self.operator.range.empty()
        for o in terms(self.operator):
            if len(self.RB):
                images.append(o.apply(self.RB))
        if self.rhs is not None:
            for o in terms(self.rhs):
                images.append(o.as_range_array())
        if self.product is not None and self.riesz_representatives:
            images = self.product.apply_inverse(images)
        basis = gram_schmidt(images, product=self.product if self.riesz_representatives else None,
                             atol=1e-13, rtol=1e-13, check=False) if len(images) else images
        prod = self.product if self.riesz_representatives else None
        if self.product is not None and self.riesz_representatives:
            # <riesz(r), b>_product = <r, b>: project the plain operator against the image basis
            lhs = project(self.operator, basis, self.RB)
            rhs = project(self.rhs, basis, None) if self.rhs is not None else None
        else:
            lhs = project(self.operator, basis, self.RB, product=prod)
            rhs = project(self.rhs, basis, None, product=prod) if self.rhs is not None else None
        return ResidualOperator(lhs, rhs)
