"""NumpyMatrixOperator: apply(U) = (M @ U.to_numpy().T).T, dense or scipy.sparse M."""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from pymor.operators.interface import Operator
from pymor.vectorarrays.numpy import NumpyVectorSpace


class NumpyMatrixOperator(Operator):
    linear = True

    def __init__(self, matrix, source_id=None, range_id=None, solver_options=None, name=None):
        if not sps.issparse(matrix):
            matrix = np.asarray(matrix)
            if matrix.ndim <= 1:
                matrix = matrix.reshape(1, -1)
        self.__auto_init(locals())
        self.source = NumpyVectorSpace(matrix.shape[1], source_id)
        self.range = NumpyVectorSpace(matrix.shape[0], range_id)
        self.sparse = sps.issparse(matrix)

    @classmethod
    def from_file(cls, *a, **k):
        raise NotImplementedError

    @property
    def H(self):
        if self.sparse:
            adj = self.matrix.transpose().conj()
        elif np.isrealobj(self.matrix):
            adj = self.matrix.T
        else:
            adj = self.matrix.T.conj()
        return self.with_(matrix=adj, source_id=self.range_id, range_id=self.source_id)

    def apply(self, U, mu=None):
        assert U in self.source
        return self.range.make_array(self.matrix.dot(U.to_numpy().T).T)

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        return self.H.apply(V, mu=mu)

    def apply_inverse(self, V, mu=None, initial_guess=None, least_squares=False):
        assert V in self.range
        rhs = V.to_numpy().T
        if self.sparse:
            assert not least_squares
            R = spla.splu(self.matrix.tocsc()).solve(np.ascontiguousarray(rhs))
        elif least_squares:
            R = np.linalg.lstsq(self.matrix, rhs, rcond=None)[0]
        else:
            R = np.linalg.solve(self.matrix, rhs)
        return self.source.make_array(R.T)

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, least_squares=False):
        return self.H.apply_inverse(U, mu=mu, least_squares=least_squares)

    def assemble(self, mu=None):
        return self
