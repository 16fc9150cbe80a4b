"""CPU restatement of the reference's `utilities/factorization.py` for the step in front of the
sketch -- TEST INFRASTRUCTURE ONLY, never imported by the product.

The reference computes these with SciPy SuperLU on the host; SciPy is the third-party dependency
the reference itself calls (unpinned, README.md:8), so the restatement IS the reference's
expression, line for line:
  * inverse_lu_apply          InverseLuOperator.apply          factorization.py:118-124
  * inverse_lu_apply_adjoint  InverseLuOperator.apply_adjoint  factorization.py:126-132
  * splu_symetric             factorization.py:17-22
  * lu_to_cholesky            factorization.py:24-52
`/root/reference/utilities/factorization.py` imports pyMOR at module level and cannot be imported
here; parity for this file is pinned by the identities checked in tests (A x = b, Q^H Q = A).
"""
import numpy as np
from scipy.sparse import csc_matrix, diags
from scipy.sparse.linalg import splu


def splu_symetric(matrix):
    """factorization.py:17-22."""
    return splu(matrix, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0, options={'SymmetricMode': True})


def factorize(matrix, symetric=False, splu_kwargs=None):
    """InverseLuOperator.__init__, factorization.py:108-116."""
    if symetric:
        return splu_symetric(matrix)
    return splu(matrix, **(splu_kwargs or {}))


def inverse_lu_apply(slu, V):
    """factorization.py:118-124: V is the (len, dim) block `U.to_numpy()`."""
    return slu.solve(V.T).T


def inverse_lu_apply_adjoint(slu, V):
    """factorization.py:126-132."""
    return slu.solve(V.T, trans='H').T


def lu_to_cholesky(matrix=None, factor=None):
    """factorization.py:24-52: Q with Q^H Q == matrix."""
    if factor is None:
        factor = splu_symetric(matrix)
    n = factor.perm_c.shape[0]
    P = csc_matrix((np.ones(n), (factor.perm_r, np.arange(n))))
    D = diags(factor.U.diagonal() ** 0.5)
    return (P.T @ factor.L @ D).conj().T
