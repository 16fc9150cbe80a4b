"""Device-side arithmetic of the reference's `SketchedReductor`
(mor/sketched_reductor.py:49-118,143-168,210-219) for affinely decomposed problems.

The reference drives these steps through pyMOR's project/expand/contract rule engine;
here the same sequence of operations is a small host-side class over device blocks so the
call sites map one to one:

    extend_basis(U)          :49-86   Theta U, Theta R^-1 A_q U for every affine term,
                                      Theta R^-1 f_p on the first call, column concatenation
    orthonormalize_basis()   :90-118  Gram-Schmidt of the SKETCHED basis, T = pinv(R),
                                      rb <- T^T rb, S_q <- S_q T
    reduce()                 :154-168 Galerkin system = Gram matrices (Theta U)^H S_q
    estimate_error(a, mu)    :216-219 sketched residual norm

Layout: a block of vectors is (len, dim); a sketched term is kept as the (m, k) block
V3 = Theta(R^-1 A_q U) and only transposed to the k x m matrix of
utilities/__init__.py:32-36 when exported with `sketched_operator_matrices()`.
"""
import numpy as np
import torch

from . import reductor_ops as ops
from .embeddings import IdentityEmbedding
from .vectorarray import DeviceVectorArray, as_device_block


class SketchedReductor:
    def __init__(self, operators, rhs, embedding_primal, embedding_online=None, inverse_product=None,
                 save_rb=True, orthonormalize=True):
        """operators: affine terms A_q (objects with .apply on DeviceVectorArray, e.g.
        vectorarray.MatrixOperator); rhs: affine right-hand-side vectors f_p as (n,) arrays;
        inverse_product: operator R^-1 or None (identity)."""
        self.operators = list(operators)
        self.rhs = [as_device_block(f).reshape(1, -1).to(torch.float64) for f in rhs]
        self.embedding_primal = embedding_primal
        self.embedding_online = embedding_online if embedding_online is not None else \
            IdentityEmbedding(embedding_primal.range)                    # :37-38
        self.inverse_product = inverse_product
        self.save_rb = save_rb
        self.orthonormalize = orthonormalize
        self.space = embedding_primal.source
        k = embedding_primal.range.dim
        dev = torch.device("cuda", torch.cuda.current_device())
        self.srb = torch.empty((0, k), dtype=torch.float64, device=dev)             # Theta U, (r, k)
        self.rb = torch.empty((0, self.space.dim), dtype=torch.float64, device=dev)
        self.s_lhs = [torch.empty((0, k), dtype=torch.float64, device=dev) for _ in self.operators]
        self.s_rhs = None                                                            # list of (k,) tensors
        self.T_total = None

    # -- helpers
    def _rinv(self, V):
        if self.inverse_product is None:
            return V
        return self.inverse_product.apply(DeviceVectorArray(self.space, V)).data

    def _sketch(self, V):
        return self.embedding_primal.apply(V)

    def extend_basis(self, U):                                           # :49-86
        U = as_device_block(U).to(torch.float64)
        if self.save_rb:
            self.rb = torch.cat([self.rb, U], dim=0)                     # :51-52
        su = self._sketch(U)                                             # :63-64
        offset = self.srb.shape[0]
        self.srb = torch.cat([self.srb, su], dim=0)                      # :65
        Uva = DeviceVectorArray(self.space, U)
        for q, A in enumerate(self.operators):                           # :69-70 (one term per q)
            V1 = A.apply(Uva).data
            V3 = self._sketch(self._rinv(V1))
            self.s_lhs[q] = torch.cat([self.s_lhs[q], V3], dim=0)        # :78, concatenate axis=1 of k x m
        if self.s_rhs is None:                                           # :72-75
            self.s_rhs = [self._sketch(self._rinv(f)).reshape(-1) for f in self.rhs]
        if self.orthonormalize:
            self.orthonormalize_basis(offset=offset)                     # :85-86

    def orthonormalize_basis(self, offset=0, T=None):                    # :90-118
        if T is None:
            Q, R = ops.gram_schmidt(self.srb, offset=offset)             # :94
            T = ops.pinv_R(R)                                            # :95  (r x r)
        else:
            T = torch.as_tensor(T, dtype=torch.float64, device=self.srb.device)
            Q = ops.gemm_nn(T.T.contiguous(), self.srb)                  # :97
        if self.save_rb:
            self.rb = ops.gemm_nn(T.T.contiguous(), self.rb)             # :99-100  rb.lincomb(T.T)
        self.srb = Q                                                     # :102
        # S_q (k x r) <- S_q T   <=>   V3 (r x k) <- T^T V3             # :104-108
        self.s_lhs = [ops.gemm_nn(T.T.contiguous(), V3) for V3 in self.s_lhs]
        return T

    def sketched_operator_matrices(self):
        """The k x r matrices the reference holds (utilities/__init__.py:32-36)."""
        return [V3.T.contiguous() for V3 in self.s_lhs]

    def reduce(self, seed=None):                                         # :121-129,154-168
        emb = self.embedding_online if seed is None else self.embedding_online.with_(_seed=seed)
        lhs = [ops.gram(self.srb, V3) for V3 in self.s_lhs]              # :161  (Theta U)^H S_q
        rhs = [ops.gram(self.srb, b.reshape(1, -1)).reshape(-1) for b in self.s_rhs]   # :162
        est_lhs = [emb.apply(V3).T.contiguous() for V3 in self.s_lhs]    # :148  Gamma S_q  (k' x r)
        est_rhs = [emb.apply(b.reshape(1, -1)).reshape(-1) for b in self.s_rhs]   # :149
        return SketchedRom(lhs, rhs, est_lhs, est_rhs)


class SketchedRom:
    """Reduced Galerkin system + sketched residual estimator (StationaryModel with
    ResidualErrorEstimator, mor/sketched_reductor.py:165-166,210-219)."""

    def __init__(self, lhs, rhs, est_lhs, est_rhs):
        self.lhs, self.rhs, self.est_lhs, self.est_rhs = lhs, rhs, est_lhs, est_rhs

    def solve(self, theta_lhs, theta_rhs):
        A = sum(t * M for t, M in zip(theta_lhs, self.lhs))
        b = sum(t * v for t, v in zip(theta_rhs, self.rhs))
        return torch.linalg.solve(A, b)

    def estimate_error(self, a, theta_lhs, theta_rhs):
        return float(ops.residual_norm(self.est_lhs, theta_lhs, self.est_rhs, theta_rhs, a).cpu())
