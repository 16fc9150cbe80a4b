"""project(op, range_basis, source_basis): Petrov-Galerkin projection by rules (pyMOR 2023.1).
Rule order matters: the reference inserts its own rule at index 3 (utilities/__init__.py:71)."""
from pymor.algorithms.rules import RuleTable, match_always, match_class, match_generic
from pymor.core.exceptions import NoMatchingRuleError, RuleNotMatchingError
from pymor.operators.constructions import (AdjointOperator, ConcatenationOperator, ConstantOperator,
                                           LincombOperator, VectorArrayOperator, ZeroOperator)
from pymor.operators.numpy import NumpyMatrixOperator
from pymor.vectorarrays.numpy import NumpyVectorSpace


def project(op, range_basis, source_basis, product=None):
    assert source_basis is None or source_basis in op.source
    assert range_basis is None or range_basis in op.range
    if range_basis is None and source_basis is None:
        return op
    if product is not None and range_basis is not None:
        range_basis = product.apply(range_basis)
    return ProjectRules(range_basis, source_basis).apply(op)


class ProjectRules(RuleTable):
    def __init__(self, range_basis, source_basis):
        super().__init__(use_caching=True)
        self.range_basis, self.source_basis = range_basis, source_basis

    @match_always
    def action_no_bases(self, op):
        if self.range_basis is None and self.source_basis is None:
            return op
        raise RuleNotMatchingError

    @match_class(ZeroOperator)
    def action_ZeroOperator(self, op):
        rb, sb = self.range_basis, self.source_basis
        return ZeroOperator(NumpyVectorSpace(len(rb)) if rb is not None else op.range,
                            NumpyVectorSpace(len(sb)) if sb is not None else op.source, name=op.name)

    @match_class(ConstantOperator)
    def action_ConstantOperator(self, op):
        raise RuleNotMatchingError

    @match_generic(lambda op: op.linear and not op.parametric, "linear and not parametric")
    def action_apply_basis(self, op):
        rb, sb = self.range_basis, self.source_basis
        if sb is None:
            try:
                V = op.apply_adjoint(rb)
            except NotImplementedError as e:
                raise RuleNotMatchingError("apply_adjoint not implemented") from e
            if isinstance(op.source, NumpyVectorSpace):
                return NumpyMatrixOperator(V.to_numpy(), source_id=op.source.id, name=op.name)
            return VectorArrayOperator(V, adjoint=True, name=op.name)
        if rb is None:
            V = op.apply(sb)
            if isinstance(op.range, NumpyVectorSpace):
                return NumpyMatrixOperator(V.to_numpy().T, range_id=op.range.id, name=op.name)
            return VectorArrayOperator(V, adjoint=False, name=op.name)
        return NumpyMatrixOperator(op.apply2(rb, sb), name=op.name)

    @match_class(ConcatenationOperator)
    def action_ConcatenationOperator(self, op):
        if len(op.operators) == 1:
            return self.apply(op.operators[0])
        rb, sb = self.range_basis, self.source_basis
        last, first = op.operators[0], op.operators[-1]
        if sb is not None and first.linear and not first.parametric:
            V = first.apply(sb)
            return type(self)(rb, V).apply(op.with_(operators=op.operators[:-1]))
        if rb is not None and last.linear and not last.parametric:
            V = last.apply_adjoint(rb)
            return type(self)(V, sb).apply(op.with_(operators=op.operators[1:]))
        # too complicated to project directly: expand into a linear combination of simple concatenations
        from pymor.algorithms.simplify import expand
        expanded = expand(op)
        if isinstance(expanded, ConcatenationOperator):
            raise RuleNotMatchingError("expansion did not simplify the concatenation")
        return self.apply(expanded)

    @match_class(AdjointOperator)
    def action_AdjointOperator(self, op):
        if op.source_product is not None or op.range_product is not None:
            raise RuleNotMatchingError
        return type(self)(self.source_basis, self.range_basis).apply(op.operator).H

    @match_class(LincombOperator)
    def action_LincombOperator(self, op):
        return self.replace_children(op).with_(solver_options=None)
