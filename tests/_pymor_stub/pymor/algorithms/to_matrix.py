import numpy as np
import scipy.sparse as sps

from pymor.operators.constructions import (AdjointOperator, ConcatenationOperator, IdentityOperator, LincombOperator,
                                           VectorArrayOperator, ZeroOperator)
from pymor.operators.numpy import NumpyMatrixOperator


def to_matrix(op, format=None, mu=None):
    """Matrix of a linear operator built from NumpyMatrixOperators (format=None keeps the
    stored matrix, dense or sparse)."""
    if isinstance(op, NumpyMatrixOperator):
        M = op.matrix
    elif isinstance(op, LincombOperator):
        cs = op.evaluate_coefficients(mu)
        M = sum(c * to_matrix(o, format, mu) for c, o in zip(cs, op.operators))
    elif isinstance(op, ConcatenationOperator):
        M = to_matrix(op.operators[0], format, mu)
        for o in op.operators[1:]:
            M = M @ to_matrix(o, format, mu)
    elif isinstance(op, IdentityOperator):
        M = sps.eye(op.source.dim)
    elif isinstance(op, ZeroOperator):
        M = np.zeros((op.range.dim, op.source.dim))
    elif isinstance(op, VectorArrayOperator):
        M = op.array.to_numpy() if op.adjoint else op.array.to_numpy().T
    elif isinstance(op, AdjointOperator):
        M = to_matrix(op.operator, format, mu).T.conj()
    else:
        raise NotImplementedError(type(op))
    if format == "dense" and sps.issparse(M):
        M = M.toarray()
    return M
