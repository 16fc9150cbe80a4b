"""GPU parity of the sparse-LU inverse operator (rla4mor_b200/factorization.py, csrc/sptrsv.cu)
against what the reference computes -- SciPy SuperLU's `slu.solve(V.T).T`
(utilities/factorization.py:118-132) on the same factorisation -- and of the Cholesky-type
sqrt_product (factorization.py:24-52)."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import splu

from golden_util import rel_fro
from oracle import factorization_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


def _fem(nx, shift=0.1):
    ex = np.ones(nx)
    T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    return (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx)) + shift * sp.eye(nx * nx)).tocsc()


@pytest.mark.parametrize("nx,m", [(9, 1), (48, 5), (64, 64), (80, 70), (120, 33)])
def test_inverse_lu_operator_matches_superlu(rb, nx, m):
    from rla4mor_b200.factorization import InverseLuOperator
    A = _fem(nx)
    n = A.shape[0]
    op = rb.MatrixOperator(A, source_id="S", range_id="S")
    inv = InverseLuOperator(op)                                   # general splu (COLAMD, partial pivoting)
    V = np.random.RandomState(nx).standard_normal((m, n))
    got = inv.apply(op.source.from_numpy(V)).to_numpy()
    ref = fo.inverse_lu_apply(inv.factorization, V)               # the reference's own expression (:118-124)
    assert rel_fro(got, ref) < 1e-12
    assert rel_fro((A @ got.T).T, V) < 1e-11
    # apply_inverse applies the matrix itself (:134-135)
    assert rel_fro(inv.apply_inverse(op.source.from_numpy(V)).to_numpy(), (A @ V.T).T) < 1e-13


def test_inverse_lu_adjoint_nonsymmetric_and_symmetric_mode(rb):
    from rla4mor_b200.factorization import InverseLuOperator, splu_symetric
    rs = np.random.RandomState(3)
    n = 1500
    A = (sp.random(n, n, 0.004, random_state=rs) + sp.eye(n) * 4.0).tocsc()       # non-symmetric, pivoting active
    op = rb.MatrixOperator(A)
    inv = InverseLuOperator(op)
    V = rs.standard_normal((7, n))
    assert rel_fro(inv.apply(op.source.from_numpy(V)).to_numpy(), fo.inverse_lu_apply(inv.factorization, V)) < 1e-12
    got = inv.apply_adjoint(op.source.from_numpy(V)).to_numpy()
    assert rel_fro(got, fo.inverse_lu_apply_adjoint(inv.factorization, V)) < 1e-12
    S = _fem(40)
    ops = rb.MatrixOperator(S)
    invs = InverseLuOperator(ops, symetric=True)                  # splu_symetric (:17-22)
    W = rs.standard_normal((12, S.shape[0]))
    assert rel_fro(invs.apply(ops.source.from_numpy(W)).to_numpy(), fo.inverse_lu_apply(fo.factorize(S, symetric=True), W)) < 1e-12


def test_cholesky_sqrt_product(rb):
    from rla4mor_b200.factorization import operator_to_cholesky, lu_to_cholesky
    S = _fem(30)
    Q = operator_to_cholesky(rb.MatrixOperator(S, source_id="S", range_id="S"))
    Qh = fo.lu_to_cholesky(S)
    assert abs(lu_to_cholesky(S) - Qh).max() == 0.0               # host part of the product = the reference's lines
    assert abs(Qh.conj().T @ Qh - S).max() < 1e-12               # Q^H Q == matrix (:37)
    V = np.random.RandomState(0).standard_normal((6, S.shape[0]))
    got = Q.apply(Q.source.from_numpy(V)).to_numpy()
    assert rel_fro(got, (Qh @ V.T).T) < 1e-13
    # a sqrt_product in front of the sketch: || Theta Q u || ~ || u ||_S
    emb = rb.GaussianEmbedding(source=Q.source, sqrt_product=Q, options={"range_dim": 400}, _seed=1)
    y = emb.apply(Q.source.from_numpy(V)).to_numpy()
    ref = np.sqrt(np.einsum("ij,ij->i", V, (S @ V.T).T))
    assert np.all(np.abs(np.linalg.norm(y, axis=1) / ref - 1.0) < 0.2)


def test_sketched_reductor_with_inverse_product(rb):
    """Theta R^-1 A_q U with R^-1 on the device (mor/sketched_reductor.py:69-70)."""
    from rla4mor_b200.factorization import InverseLuOperator
    from oracle import embeddings_oracle as eo
    from oracle import reductor_oracle as ro
    nx = 40
    n = nx * nx
    terms = [_fem(nx, 0.1).tocsr(), sp.diags(np.linspace(1.0, 2.0, n)).tocsr()]
    R = _fem(nx, 1.0)
    k = 150
    space = rb.DeviceVectorSpace(n, id="STATE")
    ops_dev = [rb.MatrixOperator(A, source_id="STATE", range_id="STATE") for A in terms]
    rinv = InverseLuOperator(rb.MatrixOperator(R, source_id="STATE", range_id="STATE"), symetric=True)
    emb = rb.SrhtEmbedding(source=space, options={"range_dim": k}, _seed=2)
    f = [np.random.RandomState(1).standard_normal(n)]
    red = rb.SketchedReductor(ops_dev, f, emb, inverse_product=rinv, orthonormalize=False)
    U = np.random.RandomState(2).standard_normal((6, n))
    red.extend_basis(U)
    lu = splu(R)
    for got, A in zip(red.sketched_operator_matrices(), terms):
        ref = eo.srht_apply(lu.solve((A @ U.T)).T, k, 2).T        # Theta R^-1 A_q U as a k x m matrix
        assert rel_fro(got.cpu().numpy(), ref) < 1e-10
    assert rel_fro(red.s_rhs[0].cpu().numpy(), eo.srht_apply(lu.solve(f[0]).reshape(1, -1), k, 2).reshape(-1)) < 1e-10


def test_lu_solve_edge_cases(rb):
    """Empty block, 1 x 1 and diagonal matrices (a single level), many right-hand sides (several
    chunks of 64, odd count), a chain (every level one row: groups of 32)."""
    import torch
    from rla4mor_b200.factorization import InverseLuOperator
    rs = np.random.RandomState(0)
    for A in (sp.csc_matrix(np.array([[2.5]])), sp.diags(np.linspace(1.0, 3.0, 50)).tocsc(),
              (sp.eye(200) * 3.0 + sp.diags([np.ones(199)], [-1]) + sp.diags([0.5 * np.ones(199)], [1])).tocsc()):
        op = rb.MatrixOperator(A)
        inv = InverseLuOperator(op)
        n = A.shape[0]
        for m in (0, 1, 131):
            V = rs.standard_normal((m, n))
            got = inv.apply(op.source.from_numpy(torch.from_numpy(V).cuda().reshape(m, n))).to_numpy()
            assert got.shape == (m, n)
            if m:
                assert rel_fro(got, fo.inverse_lu_apply(inv.factorization, V)) < 1e-13
