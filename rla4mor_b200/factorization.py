"""B200 counterpart of the reference's `utilities/factorization.py` for the step in front of the
sketch: `InverseLuOperator` -- the `inverse_product` R^-1 of
`SketchedReductor` (mor/sketched_reductor.py:69,73) -- and the Cholesky-type `sqrt_product` Q of
the embeddings (`lu_to_cholesky`, `operator_to_cholesky`).

As in the reference the sparse LU FACTORISATION is SciPy SuperLU's on the host
(factorization.py:17-22,115); what moves to the GPU is every application: the two sparse
triangular solves of `slu.solve(V.T).T` (:118-132) over the whole block of vectors
(csrc/sptrsv.cu), and Q as a CSR SpMM (vectorarray.MatrixOperator).  Same class / function names
and constructor arguments as the reference.  Real matrices only.
"""
import ctypes
import os

import numpy as np

from ._lib import check, lib, require_cuda, stream_ptr
from .vectorarray import DeviceVectorArray, DeviceVectorSpace, MatrixOperator


def splu_symetric(matrix):
    """factorization.py:17-22."""
    from scipy.sparse.linalg import splu
    return splu(matrix, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0, options={'SymmetricMode': True})


def lu_to_cholesky(matrix=None, factor=None):
    """factorization.py:24-52: Q with Q^H Q == matrix from the symmetric-mode LU (host, SciPy)."""
    from scipy.sparse import csc_matrix, diags
    if factor is None:
        assert not (matrix is None)
        factor = splu_symetric(matrix)
    n = factor.perm_c.shape[0]
    P = csc_matrix((np.ones(n), (factor.perm_r, np.arange(n))))
    D = diags(factor.U.diagonal() ** 0.5)
    return (P.T @ factor.L @ D).conj().T


def operator_to_cholesky(operator=None, factor=None):
    """factorization.py:55-81: the Cholesky factor as a (device CSR) matrix operator."""
    try:
        matrix, source_id, range_id = operator._host, operator.source.id, operator.range.id
    except AttributeError:
        matrix, source_id, range_id = None, None, None
    if matrix is not None:
        matrix = matrix.tocsc()
    M = lu_to_cholesky(matrix, factor)
    return MatrixOperator(M, source_id, range_id)


def plan_triangular(T, lower, wide_min=4096, group=128, max_multi=0):
    """Host analysis of one triangular factor (C++ in librla_b200.so, no GPU involved): dict of
    NumPy arrays -- see rla_sptrsv_plan_host in include/rla_b200.h."""
    T = T.tocsr()
    n = T.shape[0]
    rowptr = np.ascontiguousarray(T.indptr, dtype=np.int64)
    col = np.ascontiguousarray(T.indices, dtype=np.int32)
    val = np.ascontiguousarray(T.data, dtype=np.float64)
    nnz = int(T.nnz)
    n1 = max(n, 1)
    level, order, pos = (np.empty(n1, dtype=np.int32) for _ in range(3))
    rowptr2 = np.empty(n + 1, dtype=np.int64)
    col2, val2 = np.empty(max(nnz, 1), dtype=np.int32), np.empty(max(nnz, 1), dtype=np.float64)
    diag, split = np.empty(n1, dtype=np.float64), np.empty(n1, dtype=np.int64)
    grp_start, grp_rows = np.empty(n1, dtype=np.int64), np.empty(n1, dtype=np.int32)
    step_lo, step_mid, step_hi = (np.empty(n1, dtype=np.int64) for _ in range(3))
    step_kind = np.empty(n1, dtype=np.int32)
    nsteps, ngroups, nlev = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
    check(lib().rla_sptrsv_plan_host(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, 1 if lower else 0,
                                     int(wide_min), int(group), max(1, int(max_multi)),
                                     level.ctypes.data, order.ctypes.data, pos.ctypes.data,
                                     rowptr2.ctypes.data, col2.ctypes.data, val2.ctypes.data, diag.ctypes.data,
                                     split.ctypes.data, grp_start.ctypes.data, grp_rows.ctypes.data, ctypes.byref(ngroups),
                                     step_lo.ctypes.data, step_mid.ctypes.data, step_hi.ctypes.data,
                                     step_kind.ctypes.data, ctypes.byref(nsteps), ctypes.byref(nlev)),
          "rla_sptrsv_plan_host")
    ns, ng, nz = int(nsteps.value), int(ngroups.value), int(rowptr2[n])
    c = np.ascontiguousarray
    # inverses of the groups' triangular blocks (host, once per factor) and the effective divisors
    sizes = grp_rows[:ng].astype(np.int64)
    dinv_ptr = np.concatenate([[0], np.cumsum(sizes * sizes)]).astype(np.int64)
    dinv = np.empty(max(int(dinv_ptr[-1]), 1), dtype=np.float64)
    diag_eff = np.empty(n1, dtype=np.float64)
    check(lib().rla_sptrsv_group_inverses_host(n, rowptr2.ctypes.data, col2.ctypes.data, val2.ctypes.data,
                                               diag.ctypes.data, split.ctypes.data, ng, grp_start.ctypes.data,
                                               grp_rows.ctypes.data, dinv_ptr.ctypes.data, dinv.ctypes.data,
                                               diag_eff.ctypes.data), "rla_sptrsv_group_inverses_host")
    return dict(n=n, nlevels=int(nlev.value), nsteps=ns, ngroups=ng, max_multi=int(max_multi),
                dinv_ptr=dinv_ptr, dinv=dinv, diag_eff=diag_eff[:n],
                level=level[:n], order=order[:n], pos=pos[:n],
                rowptr=rowptr2, col=col2[:max(nz, 1)], val=val2[:max(nz, 1)], nnz=nz, diag=diag[:n], split=split[:n],
                grp_start=c(grp_start[:max(ng, 1)]), grp_rows=c(grp_rows[:max(ng, 1)]),
                step_lo=c(step_lo[:ns]), step_mid=c(step_mid[:ns]), step_hi=c(step_hi[:ns]), step_kind=c(step_kind[:ns]))


class TriangularFactor:
    """One triangular CSR matrix on the device with its schedule (chains of dependent rows as
    groups, levels of the group dependency graph as steps, inverse of every group's block:
    plan_triangular)."""

    def __init__(self, T, lower, unit_diagonal, device):
        torch = require_cuda()
        p = plan_triangular(T, lower)
        self.n, self.lower = p["n"], bool(lower)
        self.nlevels, self.nsteps, self.nnz, self.ngroups = p["nlevels"], p["nsteps"], p["nnz"], p["ngroups"]
        self.step_lo, self.step_hi, self.step_kind = p["step_lo"], p["step_hi"], p["step_kind"]
        self.launches = int(self.nsteps)
        self.group_levels = int((self.step_kind == 2).sum())
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.rowptr, self.col, self.val = dev(p["rowptr"]), dev(p["col"]), dev(p["val"])
        # divisor per position: 1 inside groups (the division is in the block inverse); for a unit
        # diagonal every divisor is 1 and the kernels skip the division altogether
        self.diag = None if unit_diagonal else dev(p["diag_eff"])
        self.order_host, self.pos_host = p["order"], p["pos"]            # schedule order (host): X[p] = row order[p]
        self.split = dev(p["split"])
        self.grp_start, self.grp_rows = dev(p["grp_start"]), dev(p["grp_rows"])
        self.dinv_ptr, self.dinv = dev(p["dinv_ptr"]), dev(p["dinv"])

    def solve_inplace(self, X, m):
        """T X = X on the (n, ldx) device block with the m right-hand sides contiguous, rows in
        SCHEDULE order: X[p] belongs to row order_host[p] of T."""
        ldx = X.stride(0)
        check(lib().rla_sptrsv_solve_f64(self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(),
                                         None if self.diag is None else self.diag.data_ptr(),
                                         self.split.data_ptr(), self.grp_start.data_ptr(), self.grp_rows.data_ptr(),
                                         self.dinv_ptr.data_ptr(), self.dinv.data_ptr(),
                                         self.step_lo.ctypes.data, self.step_hi.ctypes.data,
                                         self.step_kind.ctypes.data, self.nsteps, X.data_ptr(), int(m), ldx,
                                         stream_ptr()), "rla_sptrsv_solve_f64")
        return X


class SparseLU:
    """Device image of a SciPy SuperLU factorisation  Pr A Pc = L U  (real)."""

    def __init__(self, factorization, device=None):
        torch = require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(factorization.shape[0])
        L, U = factorization.L, factorization.U
        assert not np.iscomplexobj(L.data) and not np.iscomplexobj(U.data), "real factorisations only"
        self._L, self._U = L.tocsr(), U.tocsr()
        self._perm_r = np.ascontiguousarray(factorization.perm_r, dtype=np.int64)
        self._perm_c = np.ascontiguousarray(factorization.perm_c, dtype=np.int64)
        self._fwd = None
        self._adj = None
        # the few hundred launches of the two triangular solves are replayed as ONE CUDA graph per
        # (direction, number of right-hand sides): the launches of a solve are 20-50 us each, so the
        # gaps between dependent launches are a measurable part of it
        self.use_graph = os.environ.get("RLA_SPTRSV_GRAPH", "1") != "0"
        self._graphs = {}                                                # (adjoint, m) -> (X1, X2, CUDAGraph), a few shapes

    def _factors(self, adjoint):
        """(first factor, second factor, perm_in, map_mid, perm_out): the block goes in with
        X1[perm_in[i]] = b[i] (first factor's schedule order), is re-ordered X2[p] = X1[map_mid[p]] into the
        second factor's order, and comes out as x[i] = X2[perm_out[i]]."""
        import torch
        cached = self._adj if adjoint else self._fwd
        if cached is not None:
            return cached
        if not adjoint:
            f1 = TriangularFactor(self._L, True, True, self.device)
            f2 = TriangularFactor(self._U, False, False, self.device)
            p_in, p_out = self._perm_r, self._perm_c
        else:
            f1 = TriangularFactor(self._U.T.tocsr(), True, False, self.device)
            f2 = TriangularFactor(self._L.T.tocsr(), False, True, self.device)
            p_in, p_out = self._perm_c, self._perm_r
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)
        res = (f1, f2, dev(f1.pos_host[p_in]), dev(f1.pos_host[f2.order_host]), dev(f2.pos_host[p_out]))
        if adjoint:
            self._adj = res
        else:
            self._fwd = res
        return res

    def solve(self, B, adjoint=False):
        """(m, n) block of right-hand sides (one per row) -> (m, n) block of solutions of
        A x = b (or A^H x = b): what `slu.solve(V.T[, trans='H']).T` computes."""
        import torch
        assert B.is_cuda and B.dim() == 2 and B.shape[1] == self.n
        B = B.to(torch.float64)
        if B.stride(1) != 1:
            B = B.contiguous()
        m = B.shape[0]
        out = torch.empty((m, self.n), dtype=torch.float64, device=B.device)
        if m == 0 or self.n == 0:
            return out
        first, second, p_in, p_mid, p_out = self._factors(adjoint)
        ldx = m + (m & 1)

        def solves(X1, X2):
            first.solve_inplace(X1, m)
            check(lib().rla_sptrsv_permute_rows_f64(X1.data_ptr(), p_mid.data_ptr(), X2.data_ptr(), self.n, ldx,
                                                    stream_ptr()), "rla_sptrsv_permute_rows_f64")
            second.solve_inplace(X2, m)

        with torch.cuda.device(B.device):
            key = (bool(adjoint), m)
            graph = None
            if self.use_graph and key in self._graphs:
                X1, X2, graph = self._graphs[key]
            else:
                X1 = torch.empty((self.n, ldx), dtype=torch.float64, device=B.device)
                X2 = torch.empty((self.n, ldx), dtype=torch.float64, device=B.device)
            check(lib().rla_sptrsv_transpose_in_f64(B.data_ptr(), m, self.n, B.stride(0), p_in.data_ptr(),
                                                    X1.data_ptr(), ldx, stream_ptr()), "rla_sptrsv_transpose_in_f64")
            if graph is not None:
                graph.replay()
                lib().rla_launch_count_add(first.launches + second.launches + 1)
            else:
                solves(X1, X2)
                if self.use_graph:
                    # this call ran eagerly (and warmed everything up); the next one with the same
                    # shape replays the capture on the same two buffers
                    try:
                        g = torch.cuda.CUDAGraph()
                        torch.cuda.current_stream().synchronize()
                        n0 = lib().rla_launch_count()
                        with torch.cuda.graph(g):                        # records the launches, runs nothing
                            solves(X1, X2)
                        lib().rla_launch_count_add(-int(lib().rla_launch_count() - n0))   # recorded, not run
                        # extend_basis alternates between m right-hand sides (the images A_q U) and one
                        # (the right-hand side of the model): keep a few shapes, oldest out first
                        while len(self._graphs) >= 4:
                            self._graphs.pop(next(iter(self._graphs)))
                        self._graphs[key] = (X1, X2, g)
                    except Exception:                                    # capture unsupported: stay eager
                        self.use_graph = False
                        self._graphs = {}
            check(lib().rla_sptrsv_transpose_out_f64(X2.data_ptr(), m, self.n, ldx, p_out.data_ptr(),
                                                     out.data_ptr(), out.stride(0), stream_ptr()),
                  "rla_sptrsv_transpose_out_f64")
        return out


class InverseLuOperator:
    """Implicit inverse of a sparse matrix through its LU factorisation
    (factorization.py:84-138): `apply` solves with the factors, `apply_inverse` applies the
    matrix.  Factorisation on the host (SciPy SuperLU, as the reference), solves on the GPU."""
    linear = True

    def __init__(self, operator, factorization=None, symetric=False, splu_kwargs=None):
        from scipy.sparse.linalg import splu
        self.operator = operator
        self.source = operator.range
        self.range = operator.source
        self.symetric = symetric
        if splu_kwargs is None:
            splu_kwargs = dict()
        self.splu_kwargs = splu_kwargs
        if factorization is None:
            matrix = operator._host.tocsc()
            if symetric:
                factorization = splu_symetric(matrix)                    # :113-114
            else:
                factorization = splu(matrix, **splu_kwargs)              # :115-116
        self.factorization = factorization
        self._device_lu = SparseLU(factorization)

    def apply(self, U, mu=None):                                         # :118-124
        assert U in self.source
        return DeviceVectorArray(self.source, self._device_lu.solve(U.data))

    def apply_adjoint(self, U, mu=None):                                 # :126-132
        assert U in self.source
        return DeviceVectorArray(self.source, self._device_lu.solve(U.data, adjoint=True))

    def apply_inverse(self, U, mu=None):                                 # :134-135
        return self.operator.apply(U)

    def apply_inverse_adjoint(self, U, mu=None):                         # :137-138
        return self.operator.apply_adjoint(U)
