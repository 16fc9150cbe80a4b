"""expand / contract (pyMOR 2023.1 semantics): `expand` distributes concatenations over linear
combinations, `contract` applies linear non-parametric operators to neighbouring
VectorArrayOperators."""
from pymor.algorithms.rules import RuleTable, match_class
from pymor.operators.constructions import ConcatenationOperator, LincombOperator, VectorArrayOperator
from pymor.operators.interface import Operator
from pymor.operators.numpy import NumpyMatrixOperator


def expand(obj):
    return ExpandRules().apply(obj)


def contract(obj):
    return ContractRules().apply(obj)


def _merge_concatenations(op):
    if any(isinstance(o, ConcatenationOperator) for o in op.operators):
        ops = []
        for o in op.operators:
            ops.extend(o.operators if isinstance(o, ConcatenationOperator) else (o,))
        op = op.with_(operators=ops)
    return op


class ExpandRules(RuleTable):
    def __init__(self):
        super().__init__(use_caching=True)

    @match_class(LincombOperator)
    def action_LincombOperator(self, op):
        op = self.replace_children(op)
        if any(isinstance(o, LincombOperator) for o in op.operators):
            ops, coeffs = [], []
            for c, o in zip(op.coefficients, op.operators):
                if isinstance(o, LincombOperator):
                    coeffs.extend(c * cc for cc in o.coefficients)
                    ops.extend(o.operators)
                else:
                    coeffs.append(c)
                    ops.append(o)
            op = op.with_(operators=ops, coefficients=coeffs)
        return op

    @match_class(ConcatenationOperator)
    def action_ConcatenationOperator(self, op):
        op = _merge_concatenations(self.replace_children(op))
        if any(isinstance(o, LincombOperator) for o in op.operators):
            i = next(i for i, o in enumerate(op.operators) if isinstance(o, LincombOperator))
            left, right = op.operators[:i], op.operators[i + 1:]
            ops = [ConcatenationOperator(left + (o,) + right) for o in op.operators[i].operators]
            op = self.apply(op.operators[i].with_(operators=ops))
        return op

    @match_class(Operator)
    def action_recurse(self, op):
        return self.replace_children(op)


class ContractRules(RuleTable):
    def __init__(self):
        super().__init__(use_caching=True)

    @match_class(ConcatenationOperator)
    def action_ConcatenationOperator(self, op):
        op = _merge_concatenations(self.replace_children(op))
        ops = list(op.operators)
        # a |VectorArrayOperator| (array as columns) absorbs the linear, non-parametric operator on its left
        i = len(ops) - 1
        while i > 0:
            right, left = ops[i], ops[i - 1]
            if (isinstance(right, VectorArrayOperator) and not right.adjoint
                    and left.linear and not left.parametric):
                ops[i - 1:i + 1] = [VectorArrayOperator(left.apply(right.array), adjoint=False, name=right.name)]
            elif (isinstance(right, NumpyMatrixOperator) and not right.sparse
                    and left.linear and not left.parametric and not isinstance(left, NumpyMatrixOperator)):
                # a dense matrix is an array of columns: left o M = matrix of left applied to them
                # (what makes `contract(expand(Gamma @ residual.operator))`, mor/sketched_reductor.py:148,
                # a linear combination of small k' x r matrices the minres ROM can solve with)
                cols = left.apply(right.range.from_numpy(right.matrix.T))
                ops[i - 1:i + 1] = [NumpyMatrixOperator(cols.to_numpy().T, source_id=right.source.id,
                                                        range_id=left.range.id, name=right.name)]
            i -= 1
        # ... and an adjoint one (array as rows) the operator on its right
        i = 0
        while i < len(ops) - 1:
            left, right = ops[i], ops[i + 1]
            if (isinstance(left, VectorArrayOperator) and left.adjoint
                    and right.linear and not right.parametric):
                ops[i:i + 2] = [VectorArrayOperator(right.apply_adjoint(left.array), adjoint=True, name=left.name)]
            else:
                i += 1
        if len(ops) == 1:
            return ops[0]
        return op.with_(operators=ops)

    @match_class(Operator)
    def action_recurse(self, op):
        return self.replace_children(op)
