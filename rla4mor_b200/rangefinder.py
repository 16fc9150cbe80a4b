"""Sketched randomized range finder (BASELINE.json configs[4]): sketch a block of m vectors
of dimension n with Theta (k x n), then factor the small k x m sketch.

    S = Theta U^T            (the (m, k) row block `Theta.apply(U)`)
    S^T = Q R                Gram-Schmidt of the sketched vectors (as the reference's
                             SketchedReductor does, mor/sketched_reductor.py:94-95), T = pinv(R)
    or  S = V^T diag(s) W    one-sided Jacobi SVD of the sketch

`U T^T` (rows: T^T @ U) is then a basis whose SKETCH is orthonormal; the singular values of
the sketch estimate those of U to the embedding's (1 +- eps).  With `world > 1` the vector
dimension is row-sharded: every rank sketches its slab and the (m, k) partial sketches are
summed by one all-reduce (sharding.py); the small factorisation is replicated.
"""
import numpy as np
import torch

from . import reductor_ops as ops
from . import sharding


def sketch_block(U_local, n, k, seed, kind="srht", rank=0, world=1, group=None, reducer=None, check=True):
    """(m, k) sketch of the block whose slab `U_local` (m, hi - lo) lives on this rank.
    `reducer` (peer.PeerSketchReducer): exchange the partials over NVLink peer memory
    (csrc/peer.cu) instead of an NCCL all-reduce."""
    if world == 1:
        if kind == "srht":
            from .srht import get_plan
            return get_plan(n, k, seed, U_local.dtype, U_local.device).apply(U_local)
        from . import dense
        return dense.embed_apply_rng(seed, 1 if kind == "rademacher" else 0, 1.0 / np.sqrt(k), k, U_local)
    if kind == "srht":
        return sharding.srht_row_sharded(U_local, n, k, seed, rank, world, group, reducer, check)
    return sharding.gaussian_row_sharded(U_local, n, k, seed, rank, world, 1 if kind == "rademacher" else 0, group,
                                         reducer, check)


# "Twice is enough" (Daniel-Gragg-Kaufman-Stewart): a second projection pass is needed only when the
# first one shrank the row below 1/sqrt(2) of its norm.  pyMOR's gram_schmidt re-iterates below 0.9
# (kept as the default of reductor_ops.gram_schmidt and in SketchedReductor, where parity with the
# reference matters): for a random 256 x 1024 sketch that is every row past the 195th -- a third of
# the rows, 40 % of the time of the factorisation -- although no pass loses more than 14 % there.
DGKS_THRESHOLD = 2.0 ** -0.5


def thin_qr(S, reiteration_threshold=DGKS_THRESHOLD):
    """Thin QR of the k x m sketch held as the row block S (m, k): S = R^T Q, Q (m, k) with
    orthonormal rows, R (m, m) upper triangular (the layout of pyMOR's gram_schmidt,
    mor/sketched_reductor.py:94; pass reiteration_threshold=0.9 for its re-iteration rule too)."""
    return ops.gram_schmidt(S, reiteration_threshold=reiteration_threshold)


def sketch_svd(S, want_v=True, precondition=None, qr=None, T=None, w_from=None):
    """SVD of the k x m sketch held as the row block S (m, k): returns (U_rows (m, k), s (m,),
    W (m, m) or None) with S = W^T diag(s) U_rows, singular values descending
    (reductor_ops.svd_jacobi's convention).

    For k >= 2m the Jacobi iteration runs on the m x m triangular factor instead of the m x k
    sketch (QR preconditioning): S = R^T Q (Gram-Schmidt), R = W^T diag(s) Z (block Jacobi on the
    rows of R, length m), S = Z^T diag(s) (W Q).  Rows four times shorter at BASELINE configs[4]
    and fewer sweeps.

    w_from: how the rotations W come out of the Jacobi iteration on R.  "accumulate": the kernel
    applies every rotation to an identity as well (rows twice as long).  "solve": only the rows
    of R are rotated (W R = diag(s) Z) and W = diag(s) Z R^-1 afterwards with T = R^-1, which the
    reductor needs anyway (mor/sketched_reductor.py:95) -- the error of W grows with cond(R), so
    this is the default (None) only when max|r_ii| / min|r_ii| < 1e4 and no row was removed (the
    choice LAPACK's xGEJSV makes for its right factor)."""
    m, k = S.shape
    if precondition is None:
        precondition = k >= 2 * m and m >= 16
    if not precondition:
        return ops.svd_jacobi(S, want_v=want_v)
    Q, R = thin_qr(S) if qr is None else qr             # S = R^T Q, R (m', m) with m' <= m kept rows
    mk = R.shape[0]
    if w_from is None:
        w_from = "accumulate"
        if mk == m:
            d = torch.diagonal(R).abs()
            if float((d.max() / d.min()).item()) < 1e4:
                w_from = "solve"
    # Jacobi on the ROWS OF R (Gram matrix R R^T, one QR-iteration step closer to diagonal than
    # S S^T = R^T R -- this is what makes the QR a preconditioner): W R = diag(s) Z, hence
    # S = R^T Q = Z^T diag(s) (W Q): left factor Z, right factor W Q.
    M = torch.zeros((mk, m + (m & 1)), dtype=torch.float64, device=S.device)
    M[:, :m] = R
    if w_from == "solve" and mk == m:
        Z, s, _ = ops.svd_jacobi(M, want_v=False)       # rows of Z: normalised rotated rows of R
        if T is None:
            T = ops.pinv_R(R)
        W = ops.gemm_nn(Z[:, :m] * s.unsqueeze(1), T)   # (diag(s) Z) R^-1
    else:
        Z, s, W = ops.svd_jacobi(M, want_v=True)
    Urows = ops.gemm_nn(W, Q)                           # (m', m') @ (m', k)
    Zm = Z[:, :m]
    if mk < m:                                          # rank-deficient: pad with zero singular values
        s = torch.cat([s, torch.zeros(m - mk, dtype=s.dtype, device=s.device)])
        Urows = torch.cat([Urows, torch.zeros((m - mk, k), dtype=Urows.dtype, device=Urows.device)])
        Zm = torch.cat([Zm, torch.zeros((m - mk, m), dtype=Zm.dtype, device=Zm.device)])
    return Urows, s, (Zm if want_v else None)


def sketched_range_finder(U_local, n, k, seed=0, kind="srht", rank=0, world=1, group=None, svd=True, reducer=None):
    """Returns dict(sketch, Q, R, T[, s, W]) -- all small (m x k, m x m) device tensors."""
    S = sketch_block(U_local, n, k, seed, kind, rank, world, group, reducer, check=False)
    Q, R = thin_qr(S)
    if reducer is not None and world > 1:
        reducer.check_status()                          # thin_qr has synchronised already: this read is free
    out = {"sketch": S, "Q": Q, "R": R, "T": ops.pinv_R(R)}
    if svd:
        Urows, s, W = sketch_svd(S, want_v=True, qr=(Q, R), T=out["T"] if R.shape[0] == R.shape[1] else None)
        out.update(s=s, W=W, Urows=Urows)
    return out
