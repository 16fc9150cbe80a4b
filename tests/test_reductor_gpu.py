"""GPU parity of the sketched-reductor operations against the oracle's restatement of
mor/sketched_reductor.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import embeddings_oracle as eo
from oracle import reductor_oracle as ro
from golden_util import rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _fem_terms(nx, Q=4, seed=0):
    """Q block-wise 5-point stiffness-like SPD terms on an nx x nx grid (thermal-block style)."""
    n = nx * nx
    ex = np.ones(nx)
    T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    L = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))).tocsr()
    idx = np.arange(n)
    blk = (idx // nx >= nx // 2) * 2 + (idx % nx >= nx // 2)
    terms = []
    for q in range(Q):
        D = sp.diags((blk == q % 4).astype(float) + 0.01)
        terms.append((D @ L @ D + 1e-3 * sp.eye(n)).tocsr())
    return terms, n


def test_spmm_and_gemm_nn_and_gram(rb):
    from rla4mor_b200 import reductor_ops as ops
    A = sp.random(3000, 2500, density=3e-3, random_state=1, format="csr")
    u = np.random.RandomState(0).standard_normal((11, 2500))
    op = rb.MatrixOperator(A)
    got = op.apply(op.source.from_numpy(u)).to_numpy()
    assert rel_fro(got, np.asarray((A @ u.T).T)) < 1e-13
    w = np.random.RandomState(1).standard_normal((4, 3000))
    assert rel_fro(op.apply_adjoint(op.range.from_numpy(w)).to_numpy(), np.asarray((A.T @ w.T).T)) < 1e-13
    v = np.random.RandomState(2).standard_normal((7, 33))
    th = np.random.RandomState(3).standard_normal((33, 5001))
    assert rel_fro(ops.gemm_nn(_dev(v), _dev(th)).cpu().numpy(), v @ th) < 1e-12
    a = np.random.RandomState(4).standard_normal((9, 1000))
    b = np.random.RandomState(5).standard_normal((13, 1000))
    assert rel_fro(ops.gram(_dev(a), _dev(b)).cpu().numpy(), a @ b.T) < 1e-12


@pytest.mark.parametrize("r,k", [(6, 40), (20, 1000), (64, 4000), (3, 5000), (100, 2000), (256, 1024), (17, 33)])
def test_gram_schmidt_matches_oracle(rb, r, k):
    from rla4mor_b200 import reductor_ops as ops
    A = np.random.RandomState(r).standard_normal((r, k))
    A[r // 2] = A[0] * 0.999 + 1e-3 * A[r // 2]                 # nearly dependent: triggers re-iteration
    Qd, Rd = ops.gram_schmidt(_dev(A))
    Qo, Ro = ro.gram_schmidt(A)
    assert rel_fro(Qd.cpu().numpy(), Qo) < 1e-10 and rel_fro(Rd.cpu().numpy(), Ro) < 1e-12
    assert rel_fro((Qd @ Qd.T).cpu().numpy(), np.eye(Qd.shape[0])) < 1e-13
    # offset and removal of a dependent row
    A2 = np.vstack([Qo[:3], A[3:5], Qo[0] + Qo[1]])
    Q2, R2 = ops.gram_schmidt(_dev(A2), offset=3)
    Qo2, Ro2 = ro.gram_schmidt(A2, offset=3)
    assert Q2.shape == Qo2.shape and R2.shape == Ro2.shape
    assert rel_fro(Q2.cpu().numpy(), Qo2) < 1e-10 and rel_fro(R2.cpu().numpy(), Ro2) < 1e-10


@pytest.mark.parametrize("m,k", [(5, 40), (32, 600), (64, 1024), (33, 500), (256, 1024), (21, 501), (130, 258)])
def test_jacobi_svd(rb, m, k):
    from rla4mor_b200 import reductor_ops as ops
    S = np.random.RandomState(m).standard_normal((m, k)) * np.logspace(0, -6, m)[:, None]
    U, s, V = ops.svd_jacobi(_dev(S), want_v=True)
    s_ref = np.linalg.svd(S, compute_uv=False)
    assert np.max(np.abs(s.cpu().numpy() - s_ref) / s_ref[0]) < 1e-13
    rec = (V.T * s) @ U                                          # S = V^T diag(s) U_rows
    assert rel_fro(rec.cpu().numpy(), S) < 1e-12
    assert rel_fro((U @ U.T).cpu().numpy(), np.eye(m)) < 1e-10


def test_gram_schmidt_grid_kernel_semantics(rb):
    """The many-CTA kernel (csrc/factor.cu) keeps pyMOR's semantics: removal of dependent rows,
    zero rows, re-iteration, and agrees with the one-CTA kernel it replaces."""
    from rla4mor_b200 import reductor_ops as ops
    import torch
    rs = np.random.RandomState(7)
    r, k = 96, 700
    base = rs.standard_normal((40, k))
    base[10] = 0.0                                                         # a zero row
    A = np.vstack([base, rs.standard_normal((56, 40)) @ base * 1.0])     # rows 40.. are dependent on the first 40
    A[50] = rs.standard_normal(k)                                          # except these two
    A[77] = rs.standard_normal(k)
    assert rb.lib().rla_gram_schmidt_workspace_bytes(r, k) > 0            # the grid kernel applies
    Qd, Rd = ops.gram_schmidt(_dev(A))
    Qo, Ro = ro.gram_schmidt(A)
    assert Qd.shape == Qo.shape == (41, k) and Rd.shape == Ro.shape
    assert rel_fro((Qd @ Qd.T).cpu().numpy(), np.eye(41)) < 1e-12
    assert rel_fro(Rd.cpu().numpy(), Ro) < 1e-9 and rel_fro(Qd.cpu().numpy(), Qo) < 1e-8
    # same answer as the one-CTA kernel on a well-conditioned block (both follow the same recurrence)
    B = rs.standard_normal((200, 900))
    Q1, R1 = ops.gram_schmidt(_dev(B))
    Bd = _dev(B).clone()
    R2 = torch.empty((200, 200), dtype=torch.float64, device="cuda")
    fl = torch.empty((200,), dtype=torch.int32, device="cuda")
    rb._lib.check(rb.lib().rla_gram_schmidt_f64(Bd.data_ptr(), 200, 900, 900, 0, R2.data_ptr(), fl.data_ptr(),
                                                1e-13, 1e-13, 0.9, rb._lib.stream_ptr()), "gs")
    assert int(fl.sum()) == 0
    assert rel_fro(Q1.cpu().numpy(), Bd.cpu().numpy()) < 1e-12 and rel_fro(R1.cpu().numpy(), R2.cpu().numpy()) < 1e-13
    assert rel_fro((R1.T @ Q1).cpu().numpy(), B) < 1e-13                  # A = R^T Q


@pytest.mark.parametrize("m,k,want_v,force", [(64, 128, False, 2), (64, 128, True, 4), (64, 128, False, 8),
                                              (64, 128, True, 16), (256, 256, False, 0), (256, 256, True, 0),
                                              (256, 1024, False, 0), (100, 130, True, 0), (500, 512, False, 0)])
def test_jacobi_cluster_kernel(rb, m, k, want_v, force, monkeypatch):
    """Cluster-resident Jacobi (csrc/jacobi_cluster.cu): every cluster size, with and without the
    accumulated rotations, ragged m, against numpy and against the grid-synchronised block kernel."""
    from rla4mor_b200 import reductor_ops as ops
    if force:
        monkeypatch.setenv("RLA_JACOBI_CLUSTER", str(force))
    C = rb.lib().rla_svd_jacobi_cluster_size(k, m, 1 if want_v else 0)
    assert C == (force or C) and C in (2, 4, 8, 16), C                   # the cluster kernel applies
    S = np.random.RandomState(m + k).standard_normal((m, k)) * np.logspace(0, -5, m)[:, None]
    U, s, V = ops.svd_jacobi(_dev(S), want_v=want_v)
    info = [int(v) for v in ops.svd_jacobi.last_info.tolist()]
    assert info[1] == 1 and info[0] < 20, info
    s_ref = np.linalg.svd(S, compute_uv=False)
    assert np.max(np.abs(s.cpu().numpy() - s_ref) / s_ref[0]) < 1e-13
    assert np.max(np.abs(s.cpu().numpy() - s_ref) / s_ref) < 1e-9        # high relative accuracy
    r_num = min(m, k)
    Un = U[:r_num].cpu().numpy()
    assert rel_fro(Un @ Un.T, np.eye(r_num)) < 1e-10
    if want_v:
        assert rel_fro(((V.T * s) @ U).cpu().numpy(), S) < 1e-12
        assert rel_fro((V @ V.T).cpu().numpy(), np.eye(m)) < 1e-12
    monkeypatch.setenv("RLA_JACOBI_CLUSTER", "0")
    U2, s2, V2 = ops.svd_jacobi(_dev(S), want_v=want_v)
    assert np.max(np.abs((s - s2).cpu().numpy()) / s_ref[0]) < 1e-13


def test_jacobi_block_vs_round_kernel(rb):
    from rla4mor_b200 import reductor_ops as ops
    S = np.random.RandomState(11).standard_normal((70, 512)) * np.logspace(0, -9, 70)[:, None]
    U1, s1, V1 = ops.svd_jacobi(_dev(S), want_v=True, cluster=False)
    U2, s2, V2 = ops.svd_jacobi(_dev(S), want_v=True, block=False, cluster=False)
    s_ref = np.linalg.svd(S, compute_uv=False)
    assert np.max(np.abs(s1.cpu().numpy() - s_ref) / s_ref) < 1e-9        # high RELATIVE accuracy (Jacobi)
    assert np.max(np.abs(s2.cpu().numpy() - s_ref) / s_ref) < 1e-9
    assert rel_fro(((V1.T * s1) @ U1).cpu().numpy(), S) < 1e-12
    U3, s3, _ = ops.svd_jacobi(_dev(S), want_v=False, cluster=False)
    assert np.max(np.abs(s3.cpu().numpy() - s_ref) / s_ref) < 1e-9


def test_sketch_svd_qr_preconditioned(rb):
    from rla4mor_b200.rangefinder import sketch_svd
    rs = np.random.RandomState(5)
    for m, k, decay in ((64, 512, -6), (37, 300, -3), (48, 400, -18)):   # last one: numerically rank deficient
        S = rs.standard_normal((m, m)) @ (np.logspace(0, decay, m)[:, None] * rs.standard_normal((m, k)))
        U, s, W = sketch_svd(_dev(S), want_v=True, precondition=True)
        s_ref = np.linalg.svd(S, compute_uv=False)
        assert np.max(np.abs(s.cpu().numpy() - s_ref)) / s_ref[0] < 1e-12
        rec = (W.T * s) @ U
        assert rel_fro(rec.cpu().numpy(), S) < 1e-11
        r_num = int((s_ref > 1e-10 * s_ref[0]).sum())
        Un = U[:r_num].cpu().numpy()
        assert rel_fro(Un @ Un.T, np.eye(r_num)) < 1e-6


def test_sketch_svd_w_from_triangular_inverse(rb):
    """Well-conditioned R: only the rows of R are rotated (16 rows per block) and W = diag(s) Z R^-1;
    same factorisation as with accumulated rotations."""
    from rla4mor_b200.rangefinder import sketch_svd
    rs = np.random.RandomState(8)
    for m, k in ((64, 512), (96, 200), (256, 1024)):
        S = rs.standard_normal((m, k))
        s_ref = np.linalg.svd(S, compute_uv=False)
        got = {}
        for mode in ("solve", "accumulate", None):
            U, s, W = sketch_svd(_dev(S), want_v=True, precondition=True, w_from=mode)
            assert np.max(np.abs(s.cpu().numpy() - s_ref)) / s_ref[0] < 1e-13
            assert rel_fro(((W.T * s) @ U).cpu().numpy(), S) < 1e-12
            assert rel_fro((W @ W.T).cpu().numpy(), np.eye(m)) < 1e-11
            assert rel_fro((U @ U.T).cpu().numpy(), np.eye(m)) < 1e-11
            got[mode] = (U.cpu().numpy(), W.cpu().numpy())
        # singular vectors agree up to sign (distinct singular values)
        for a, b in zip(got["solve"], got["accumulate"]):
            sg = np.sign(np.sum(got["solve"][0] * got["accumulate"][0], axis=1))
            assert rel_fro(a * sg[:, None], b) < 1e-8


def test_pinv_R_triangular_inverse(rb):
    from rla4mor_b200 import reductor_ops as ops
    for r in (1, 5, 31, 32, 33, 64, 128, 129, 256, 257, 600, 1024, 1100):      # row kernel up to 1024, column kernel beyond
        R = np.triu(np.random.RandomState(r).standard_normal((r, r))) / np.sqrt(r) + 3.0 * np.eye(r)
        T = ops.pinv_R(_dev(R)).cpu().numpy()
        assert rel_fro(T, np.linalg.inv(R)) < 1e-12 and np.all(np.tril(T, -1) == 0.0)
        assert rel_fro(T @ R, np.eye(r)) < 1e-13
    Rr = np.random.RandomState(0).standard_normal((3, 5))
    assert rel_fro(ops.pinv_R(_dev(Rr)).cpu().numpy(), np.linalg.pinv(Rr)) < 1e-12


def test_residual_norm(rb):
    from rla4mor_b200 import reductor_ops as ops
    rs = np.random.RandomState(0)
    S = [rs.standard_normal((300, 12)) for _ in range(3)]
    b = [rs.standard_normal(300) for _ in range(2)]
    a = rs.standard_normal(12)
    got = float(ops.residual_norm([_dev(x) for x in S], [0.5, -1.0, 2.0], [_dev(x) for x in b], [1.0, 0.3], _dev(a)).cpu())
    ref = ro.residual_norm(S, [0.5, -1.0, 2.0], [x.reshape(-1, 1) for x in b], [1.0, 0.3], a)
    assert abs(got - ref) / ref < 1e-13


@pytest.mark.parametrize("kind", ["srht", "gauss"])
def test_sketched_reductor_pipeline(rb, kind):
    terms, n = _fem_terms(48)
    rs = np.random.RandomState(3)
    f = [rs.standard_normal(n)]
    k = 200
    space = rb.DeviceVectorSpace(n, id="STATE")
    ops_dev = [rb.MatrixOperator(A, source_id="STATE", range_id="STATE") for A in terms]
    if kind == "srht":
        emb = rb.SrhtEmbedding(source=space, options={"range_dim": k}, _seed=1)
        theta_apply = lambda V: eo.srht_apply(V, k, 1)
    else:
        emb = rb.GaussianEmbedding(source=space, options={"range_dim": k}, _seed=1)
        theta = eo.gaussian_random_matrix(k, n, 1)
        theta_apply = lambda V: eo.gaussian_apply(V, theta)
    red = rb.SketchedReductor(ops_dev, f, emb)
    U1, U2 = rs.standard_normal((5, n)), rs.standard_normal((4, n))
    # oracle side: same sequence of operations (reductor_oracle.py)
    srb = np.zeros((0, k)); S = [np.zeros((k, 0)) for _ in terms]; rbo = np.zeros((0, n))
    for U in (U1, U2):
        red.extend_basis(U)
        off = srb.shape[0]
        srb = np.vstack([srb, theta_apply(U)]); rbo = np.vstack([rbo, U])
        new = ro.sketch_affine_terms(theta_apply, terms, U)
        S = [np.hstack([s0, s1]) for s0, s1 in zip(S, new)]
        srb, R, T, S = ro.orthonormalize_sketch(srb, S, offset=off)
        rbo = ro.update_basis(rbo, T)
    srhs = [theta_apply(f[0].reshape(1, -1)).reshape(-1, 1)]
    assert rel_fro(red.srb.cpu().numpy(), srb) < 1e-9
    assert rel_fro(red.rb.cpu().numpy(), rbo) < 1e-9
    for got, ref in zip(red.sketched_operator_matrices(), S):
        assert rel_fro(got.cpu().numpy(), ref) < 1e-9
    rom = red.reduce()
    lhs, rhs = ro.galerkin_system(srb, S, srhs)
    for got, ref in zip(rom.lhs, lhs):
        assert rel_fro(got.cpu().numpy(), ref) < 1e-9
    assert rel_fro(rom.rhs[0].cpu().numpy(), rhs[0][:, 0]) < 1e-9
    th = [1.0, 0.5, 2.0, 0.1]
    a = rom.solve(th, [1.0])
    err = rom.estimate_error(a, th, [1.0])
    ref_err = ro.residual_norm(S, th, srhs, [1.0], a.cpu().numpy())
    assert abs(err - ref_err) / ref_err < 1e-7
    # the sketched residual tracks the true residual of the full model
    u_full = a.cpu().numpy() @ rbo
    true_res = sum(t * (A @ u_full) for t, A in zip(th, terms)) - f[0]
    assert 0.5 < err / np.linalg.norm(true_res) < 1.5


def test_range_finder_small(rb):
    from rla4mor_b200.rangefinder import sketched_range_finder
    rs = np.random.RandomState(0)
    m, n, k, r = 24, 30000, 256, 6
    U = (rs.standard_normal((m, r)) * np.logspace(0, -3, r)) @ rs.standard_normal((r, n)) + 1e-9 * rs.standard_normal((m, n))
    Ud = _dev(U)
    for kind in ("srht", "gauss"):
        res = sketched_range_finder(Ud, n, k, seed=1, kind=kind)
        s_true = np.linalg.svd(U, compute_uv=False)
        s = res["s"].cpu().numpy()
        assert np.all(np.abs(s[:r] / s_true[:r] - 1.0) < 0.35)           # (1 +- eps) embedding of the range
        assert s[r] / s[0] < 1e-6                                         # numerical rank revealed
        Q = res["Q"].cpu().numpy()
        assert rel_fro(Q @ Q.T, np.eye(Q.shape[0])) < 1e-10
        # oracle parity of the factorisation on the same sketch
        Qo, Ro = ro.gram_schmidt(res["sketch"].cpu().numpy())
        assert Q.shape == Qo.shape and rel_fro(res["R"].cpu().numpy(), Ro) < 1e-8


def test_sketched_reductor_at_fem_size(rb):
    """configs[3]-sized problem: n ~ 1e6, Q = 4 affine terms; checked through properties
    (the oracle sketches only a few vectors at this size)."""
    terms, n = _fem_terms(1000)                                           # n = 1e6, ~5 nnz/row
    k, m = 1000, 16
    space = rb.DeviceVectorSpace(n, id="STATE")
    ops_dev = [rb.MatrixOperator(A, source_id="STATE", range_id="STATE") for A in terms]
    emb = rb.SrhtEmbedding(source=space, options={"range_dim": k}, _seed=3)
    g = torch.Generator(device="cuda").manual_seed(0)
    U = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    f = [torch.randn(n, dtype=torch.float64, device="cuda", generator=g)]
    red = rb.SketchedReductor(ops_dev, f, emb)
    red.extend_basis(U)
    assert red.srb.shape == (m, k) and all(V.shape == (m, k) for V in red.s_lhs)
    # sketched basis orthonormal, sketch of the transformed basis equals the transformed sketch
    assert float(torch.linalg.norm(red.srb @ red.srb.T - torch.eye(m, device="cuda", dtype=torch.float64))) < 1e-10
    assert float(torch.linalg.norm(emb.apply(red.rb) - red.srb) / torch.linalg.norm(red.srb)) < 1e-10
    # one affine term against the oracle on two vectors
    from oracle import embeddings_oracle as eo
    v = red.rb[:2].cpu().numpy()
    ref = eo.srht_apply(np.asarray((terms[1] @ v.T).T), k, 3)
    assert rel_fro(red.s_lhs[1][:2].cpu().numpy(), ref) < 1e-11
    rom = red.reduce()
    th = [1.0, 0.7, 1.3, 0.2]
    a = rom.solve(th, [1.0])
    err = rom.estimate_error(a, th, [1.0])
    u_full = a @ red.rb
    true_res = sum(t * op.apply(space.from_numpy(u_full.reshape(1, -1))).data[0] for t, op in zip(th, ops_dev)) - f[0]
    assert 0.8 < err / float(torch.linalg.norm(true_res)) < 1.2           # k = 1000 embedding of one vector


def test_thin_qr_dgks_threshold(rb):
    """The range finder's QR re-iterates by the Daniel-Gragg-Kaufman-Stewart rule (1/sqrt(2)) instead
    of pyMOR's 0.9: orthogonality stays at rounding level on random, graded and nearly dependent rows,
    and the factors agree with the 0.9 rule's."""
    from rla4mor_b200.rangefinder import thin_qr
    rs = np.random.RandomState(21)
    blocks = [rs.standard_normal((256, 1024)),
              rs.standard_normal((96, 96)) @ (np.logspace(0, -8, 96)[:, None] * rs.standard_normal((96, 400)))]
    near = rs.standard_normal((64, 300))
    near[40] = near[3] + 1e-7 * near[40]                                  # forces a second pass under either rule
    blocks.append(near)
    for A in blocks:
        Q, R = thin_qr(_dev(A))
        Q9, R9 = thin_qr(_dev(A), reiteration_threshold=0.9)
        r = Q.shape[0]
        assert Q.shape == Q9.shape and r == A.shape[0]
        assert rel_fro((Q @ Q.T).cpu().numpy(), np.eye(r)) < 1e-12
        assert rel_fro((R.T @ Q).cpu().numpy(), A) < 1e-13
        assert rel_fro(R.cpu().numpy(), R9.cpu().numpy()) < 1e-9


@pytest.mark.parametrize("r,k", [(300, 700), (200, 512), (90, 256)])
def test_gram_schmidt_many_reiterations_ragged_groups(rb, r, k):
    """Random rows filling a large part of the space: pyMOR's rule re-iterates every row past
    i ~ 0.19 k, and the number of CTAs (75, 50, 45) is not a multiple of the leader-group size of the
    two-level sum of the partial corrections.  R and Q against the oracle's sequential loop."""
    from rla4mor_b200 import reductor_ops as ops
    A = np.random.RandomState(r * 7 + k).standard_normal((r, k))
    Qd, Rd = ops.gram_schmidt(_dev(A))
    Qo, Ro = ro.gram_schmidt(A)
    assert Qd.shape == Qo.shape == (r, k)
    assert rel_fro(Rd.cpu().numpy(), Ro) < 1e-12 and rel_fro(Qd.cpu().numpy(), Qo) < 1e-10
    assert rel_fro((Qd @ Qd.T).cpu().numpy(), np.eye(r)) < 1e-13
