"""CPU restatement of the sketch arithmetic of
/root/reference/mor/sketched_reductor.py -- TEST INFRASTRUCTURE ONLY.

The reference drives this arithmetic through pyMOR's `project` / `expand` /
`contract` rule engine and `gram_schmidt` (pyMOR is a third-party dependency,
unpinned in the reference's README.md:8, absent from /root/reference and from
this image).  What is restated here is what those calls compute on NumPy data
in the reference's layout (a block of vectors is (len, dim), one vector per
row), following
  * sketched_reductor.py:62-81   (Theta U, Theta R^-1 A_q U, Theta R^-1 f_q)
  * utilities/__init__.py:32-36  (projected term = V.to_numpy().T, i.e. k x m)
  * utilities/utilities.py:61-70 (column concatenation of the affine terms)
  * sketched_reductor.py:90-118  (gram_schmidt of the sketched basis, T = pinv(R))
  * sketched_reductor.py:154-168 (Galerkin system = Gram matrices)
  * sketched_reductor.py:216-219 (residual norm)
`gram_schmidt` restates pyMOR's published algorithm (pymor/algorithms/
gram_schmidt.py, 2023.x: modified Gram-Schmidt with re-iteration while the norm
drops below `reiteration_threshold * old_norm`, R accumulated as R[j, i] += p).
Parity for this file is "unpinned by executable reference"; it is pinned by the
algebraic identities checked in tests/test_oracle_golden.py (Q R == A, Q Q^H == I).
"""
import numpy as np


def gram_schmidt(A, offset=0, atol=1e-13, rtol=1e-13, reiterate=True,
                 reiteration_threshold=9e-1):
    """pyMOR `gram_schmidt(A, offset=offset, return_R=True)` on the rows of A
    (Euclidean product), as called at sketched_reductor.py:94.
    Returns (Q, R) with rows of Q orthonormal and A = R.T-combination:
    A[i] = sum_j R[j, i] Q[j]."""
    A = np.array(A, copy=True)
    r = A.shape[0]
    R = np.eye(r, dtype=A.dtype)
    remove = []
    for i in range(offset, r):
        initial_norm = np.linalg.norm(A[i])
        if initial_norm <= atol:
            remove.append(i)
            continue
        if i == 0:
            A[0] *= 1 / initial_norm
            R[i, i] = initial_norm
            continue
        norm = initial_norm
        while True:
            for j in range(i):
                if j in remove:
                    continue
                p = np.vdot(A[j], A[i])            # A[j].inner(A[i])
                A[i] -= p * A[j]                   # axpy(-p, A[j])
                R[j, i] += p
            old_norm, norm = norm, np.linalg.norm(A[i])
            if norm <= rtol * initial_norm:
                remove.append(i)
                break
            if reiterate and norm < reiteration_threshold * old_norm:
                continue
            A[i] *= 1 / norm
            R[i, i] = norm
            break
    if remove:
        A = np.delete(A, remove, axis=0)
        R = np.delete(R, remove, axis=0)
    return A, R


def sketch_affine_terms(theta_apply, A_terms, U, Rinv_apply=None):
    """sketched_reductor.py:69-70 after `expand` (utilities/__init__.py:44-68):
    for each affine term A_q (scipy sparse or dense, n x n) evaluate
    right-to-left on the block U (m, n): V1 = A_q U, V2 = R^-1 V1, V3 = Theta V2,
    and return the k x m matrix V3.T (utilities/__init__.py:32-36)."""
    out = []
    for A in A_terms:
        V1 = np.asarray((A @ U.T).T)
        V2 = V1 if Rinv_apply is None else Rinv_apply(V1)
        V3 = theta_apply(V2)
        out.append(np.ascontiguousarray(V3.T))
    return out


def orthonormalize_sketch(srb, S_terms, offset=0):
    """sketched_reductor.py:90-108: Q, R = gram_schmidt(srb); T = pinv(R);
    every sketched term S_q (k x r) becomes S_q @ T (project(., None, V) with
    V = from_numpy(T.T)).  Returns (Q, R, T, new S_terms)."""
    Q, R = gram_schmidt(srb, offset=offset)
    T = np.linalg.pinv(R)                                    # :95
    new_terms = [S @ T for S in S_terms]                     # :104-108
    return Q, R, T, new_terms


def update_basis(rb, T):
    """sketched_reductor.py:99-100: rb.lincomb(T.T) on an (r, n) row block."""
    return T.T @ rb


def galerkin_system(srb, S_terms, s_rhs_terms):
    """sketched_reductor.py:161-162: reduced_lhs_q = (Theta U)^H S_q (r x r),
    reduced_rhs_q = (Theta U)^H b_q (r x 1)."""
    lhs = [srb.conj() @ S for S in S_terms]
    rhs = [srb.conj() @ b for b in s_rhs_terms]
    return lhs, rhs


def residual_norm(S_terms, theta_lhs, s_rhs_terms, theta_rhs, a):
    """ResidualErrorEstimator.estimate_error, sketched_reductor.py:216-219 with
    pyMOR's ResidualOperator: || sum_q th_q S_q a - sum_q th'_q b_q ||_2."""
    res = sum(t * (S @ a) for t, S in zip(theta_lhs, S_terms))
    res = res - sum(t * b.reshape(res.shape) for t, b in zip(theta_rhs, s_rhs_terms))
    return float(np.linalg.norm(res))
