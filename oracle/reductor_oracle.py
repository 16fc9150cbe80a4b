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


# ---------------------------------------------------------------------------------------------
# The whole SketchedReductor flow on NumPy data (checked against the EXECUTED reference:
# tests/golden/reductor_reference.npz, produced by oracle/make_golden_pymor.py).
class SketchedReductorOracle:
    """mor/sketched_reductor.py:22-208 for an affine problem
    A(mu) = sum_q thA_q(mu) A_q, f(mu) = sum_p thf_p(mu) f_p, outputs L (n_out x n).

    theta_apply / online_apply(seed) map an (m, n) row block to its (m, k) sketch."""

    def __init__(self, A_terms, f_terms, out_matrix, theta_apply, online_apply, Rinv_apply=None, R=None,
                 projection="galerkin", orthonormalize=True, save_rb=True):
        assert projection in ("galerkin", "minres")                                # :25
        self.A, self.f, self.L = A_terms, f_terms, out_matrix
        self.theta_apply, self.online_apply = theta_apply, online_apply
        self.Rinv = (lambda V: V) if Rinv_apply is None else Rinv_apply
        self.R = R
        self.projection, self.orthonormalize, self.save_rb = projection, orthonormalize, save_rb
        k = None
        self.srb = None                        # (r, k)   :40
        self.rb = None                         # (r, n)   :41
        self.S = None                          # list of k x r matrices (residual.operator terms)
        self.b = None                          # list of (k,) vectors  (residual.rhs terms)
        self.out = None                        # n_out x r

    def extend_basis(self, U):                                                     # :49-86
        U = np.atleast_2d(U)
        self.rb = U.copy() if self.rb is None else np.vstack([self.rb, U])         # :51-52
        out_proj = (self.L @ U.T)                                                  # :56  project(out, None, U)
        self.out = out_proj if self.out is None else np.concatenate([self.out, out_proj], axis=1)   # :57-59
        su = self.theta_apply(U)                                                   # :63-64
        self.srb = su if self.srb is None else np.vstack([self.srb, su])           # :65
        sop = sketch_affine_terms(self.theta_apply, self.A, U, self.Rinv)          # :69-70
        if self.S is None:
            self.b = [self.theta_apply(self.Rinv(np.atleast_2d(f)))[0] for f in self.f]     # :72-75
            self.S = sop
        else:
            self.S = [np.concatenate([S, s], axis=1) for S, s in zip(self.S, sop)]          # :77-79
        if self.orthonormalize:
            self.orthonormalize_basis(offset=self.srb.shape[0] - U.shape[0])       # :85-86

    def orthonormalize_basis(self, offset=0, T=None):                              # :90-118
        if T is None:
            Q, R = gram_schmidt(self.srb, offset=offset)                           # :94
            T = np.linalg.pinv(R)                                                  # :95
        else:
            Q = T.T @ self.srb                                                     # :97
        if self.save_rb:
            self.rb = T.T @ self.rb                                                # :99-100
        self.srb = Q                                                               # :102
        self.S = [S @ T for S in self.S]                                           # :104-108
        self.out = self.out @ T                                                    # :111
        return T

    def sketch_residual(self, seed):                                               # :143-152
        lhs = [self.online_apply(seed, S.T).T for S in self.S]                     # Gamma S_q   (k' x r)
        rhs = [self.online_apply(seed, np.atleast_2d(b))[0] for b in self.b]
        return lhs, rhs

    def reduce(self, seed=None):                                                   # :121-141
        if self.srb is None or self.srb.shape[0] == 0:
            return self._reduce_empty()
        if self.projection == "galerkin":
            est = self.sketch_residual(seed)
            lhs, rhs = galerkin_system(self.srb, self.S, self.b)                   # :161-162
            return RomOracle(lhs, rhs, self.out, est, least_squares=False)
        if not hasattr(seed, "__len__"):
            seed = (seed, seed)                                                    # :132-133
        lhs, rhs = self.sketch_residual(seed[0])                                   # :173-175
        est = self.sketch_residual(seed[1])                                        # :178
        return RomOracle(lhs, rhs, self.out, est, least_squares=True)

    def _reduce_empty(self):                                                       # :189-208
        r = 0
        n = self.A[0].shape[0]
        lhs = [np.zeros((0, 0)) for _ in self.A]
        rhs = [np.zeros((0,)) for _ in self.f]
        # residual of the empty basis = -f(mu), measured through Riesz representatives: ||f||_{R^-1}
        return EmptyRomOracle(self.f, self.R, self.L.shape[0])


class RomOracle:
    """StationaryModel + ResidualErrorEstimator (:165-166, :181-182, :210-219)."""

    def __init__(self, lhs, rhs, out, est, least_squares):
        self.lhs, self.rhs, self.out, self.est, self.least_squares = lhs, rhs, out, est, least_squares

    def solve(self, thA, thf):
        A = sum(t * M for t, M in zip(thA, self.lhs))
        b = sum(t * v for t, v in zip(thf, self.rhs))
        if self.least_squares:
            return np.linalg.lstsq(A, b, rcond=None)[0]                            # LsOperator, other_operators.py:32-33
        return np.linalg.solve(A, b)

    def estimate_error(self, a, thA, thf):
        return residual_norm(self.est[0], thA, self.est[1], thf, a)

    def output(self, a):
        return self.out @ a


class EmptyRomOracle:
    def __init__(self, f, R, n_out):
        self.f, self.R, self.n_out = f, R, n_out

    def solve(self, thA, thf):
        return np.zeros((0,))

    def estimate_error(self, a, thA, thf):
        import scipy.sparse.linalg as spla
        fm = sum(t * v for t, v in zip(thf, self.f))
        if self.R is None:
            return float(np.linalg.norm(fm))
        return float(np.sqrt(fm @ spla.spsolve(self.R.tocsc(), fm)))
