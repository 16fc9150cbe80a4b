"""GPU parity of the embedding-operator API (rla/embeddings.py) against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle import embeddings_oracle as eo
from golden_util import emb_golden, rel_fro

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


def test_srht_embedding_apply_kinds(rb):
    n, k, seed = 5000, 64, 5
    src = rb.DeviceVectorSpace(n, id="STATE")
    emb = rb.SrhtEmbedding(source=src, options={"range_dim": k}, range_id="SK", _seed=seed)
    assert emb.range.dim == k and emb.range.id == "SK" and emb.source is src and emb.linear
    x = np.random.RandomState(0).standard_normal((3, n))
    ref = eo.srht_apply(x, k, seed)
    y_np = emb.apply(x)
    assert isinstance(y_np, np.ndarray) and rel_fro(y_np, ref) < TOL
    U = src.from_numpy(x)
    Y = emb.apply(U)
    assert Y in emb.range and len(Y) == 3 and rel_fro(Y.to_numpy(), ref) < TOL
    y_t = emb.apply(torch.from_numpy(x).cuda())
    assert y_t.is_cuda and rel_fro(y_t.cpu().numpy(), ref) < TOL
    with pytest.raises(AssertionError):
        emb.apply(rb.DeviceVectorSpace(n, id="OTHER").from_numpy(x))      # embeddings.py:168


def test_srht_embedding_explicit_matrix_bit_exact_and_adjoint(rb):
    z = emb_golden()
    for tag in ("rows_pow2", "rows_nonpow2", "rows_5000"):
        n, k, seed = (int(v) for v in z[tag + "__meta"])
        emb = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=seed)
        rows = emb._get_random_rows(z[tag + "__indices"])
        assert np.array_equal(rows, z[tag + "__rows"])                    # bit-identical to the reference
    n, k, seed = 100, 16, 7
    emb = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=seed)
    mat = emb.get_matrix()
    assert np.array_equal(mat, eo.srht_matrix(n, k, seed))
    assert np.array_equal(emb.as_source_array().to_numpy(), mat)
    assert np.array_equal(emb.as_range_array().to_numpy(), mat.T)
    v = np.random.RandomState(1).standard_normal((4, k))
    assert rel_fro(emb.apply_adjoint(v), eo.srht_apply_adjoint(v, n, k, seed)) < TOL
    # duplicates among the sampled rows (k > n forces them)
    emb2 = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(16), options={"range_dim": 40}, _seed=9)
    v2 = np.random.RandomState(2).standard_normal((3, 40))
    assert rel_fro(emb2.apply_adjoint(v2), eo.srht_apply_adjoint(v2, 16, 40, 9)) < TOL
    # the documented scale quirk: apply vs get_matrix differ by sqrt(2^d / n)
    x = np.random.RandomState(3).standard_normal((2, n))
    assert rel_fro(emb.apply(x) * np.sqrt(n / 128), x @ mat.T) < 1e-12


def test_srht_embedding_with_sqrt_product_and_seed_changes(rb):
    n, k = 300, 24
    Qm = np.random.RandomState(4).standard_normal((n, n))
    Q = rb.MatrixOperator(Qm, source_id="S", range_id="S")
    emb = rb.SrhtEmbedding(sqrt_product=Q, options={"range_dim": k}, _seed=11)
    x = np.random.RandomState(5).standard_normal((5, n))
    assert rel_fro(emb.apply(Q.source.from_numpy(x)).to_numpy(), eo.srht_apply(x, k, 11, Q=Qm)) < 1e-11
    assert rel_fro(emb.get_matrix(), eo.srht_matrix(n, k, 11, Q=Qm)) < 1e-12
    e2 = emb.with_(_seed=12)
    assert e2 is not emb and e2._seed == 12 and e2.range.dim == k
    assert rel_fro(e2.apply(Q.source.from_numpy(x)).to_numpy(), eo.srht_apply(x, k, 12, Q=Qm)) < 1e-11
    emb.set_seed(13)
    assert rel_fro(emb.apply(Q.source.from_numpy(x)).to_numpy(), eo.srht_apply(x, k, 13, Q=Qm)) < 1e-11
    with pytest.raises(AssertionError):
        rb.SrhtEmbedding(options={"range_dim": 3})
    with pytest.raises(AssertionError):
        rb.SrhtEmbedding(source=rb.DeviceVectorSpace(10), options={"epsilon": 0.1})


def test_gaussian_embedding_reference_rng(rb):
    z = emb_golden()
    for tag in ("gauss_small", "gauss_mid"):
        m, n, k, seed = (int(v) for v in z[tag + "__meta"])
        emb = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=seed)
        assert np.array_equal(emb._random_matrix, z[tag + "__theta"])     # same Theta as the reference
        assert np.array_equal(emb.get_random_matrix(), z[tag + "__theta"])
        assert rel_fro(emb.apply(z[tag + "__U"]), z[tag + "__Y"]) < TOL
    n, k = 2000, 300
    emb = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"epsilon": 0.5, "delta": 0.1, "oblivious_dim": 3}, _seed=3)
    assert emb.range.dim == eo.gaussian_compute_dim({"epsilon": 0.5, "delta": 0.1, "oblivious_dim": 3})
    x = np.random.RandomState(1).standard_normal((7, n))
    theta = eo.gaussian_random_matrix(emb.range.dim, n, 3)
    assert rel_fro(emb.apply(x), eo.gaussian_apply(x, theta)) < TOL
    v = np.random.RandomState(2).standard_normal((2, emb.range.dim))
    assert rel_fro(emb.apply_adjoint(v), v @ theta) < TOL
    emb.set_seed(4)
    assert np.array_equal(emb._random_matrix, eo.gaussian_random_matrix(emb.range.dim, n, 4))


def test_gaussian_embedding_on_the_fly(rb):
    n, k = 30000, 200
    for mode in ("philox", "philox_rademacher"):
        emb = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": mode}, _seed=77)
        assert emb._random_matrix is None                                 # never materialised by apply
        x = np.random.RandomState(0).standard_normal((9, n))
        y = emb.apply(x)
        theta = emb.get_random_matrix()                                   # export of what the kernel used
        assert rel_fro(y, eo.gaussian_apply(x, theta)) < TOL
        assert abs(np.linalg.norm(theta) ** 2 / n - 1.0) < 0.05           # E||Theta e_j||^2 = 1
    with pytest.raises(AssertionError):
        rb.GaussianEmbedding(source=rb.DeviceVectorSpace(8), options={"range_dim": 2, "rng": "bogus"}).apply(np.zeros((1, 8)))


def test_float32_blocks_through_the_embedding_classes(rb):
    """options['rng'] = 'philox_tf32' / 'philox_rademacher': float32 blocks (device, NumPy, pinned host)
    go through the tcgen05 path and match `U @ get_random_matrix().T` to the FP32 tolerance; FP64
    blocks use the same Theta to 1e-12."""
    import torch
    n, k = 20000, 150
    x32 = np.random.RandomState(4).standard_normal((70, n)).astype(np.float32)
    for mode in ("philox_tf32", "philox_rademacher"):
        emb = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": mode}, _seed=5)
        theta = emb.get_random_matrix()
        ref = eo.gaussian_apply(x32.astype(np.float64), theta)
        assert rel_fro(emb.apply(x32), ref) < 1e-5                                    # NumPy float32 block
        assert rel_fro(emb.apply(torch.from_numpy(x32).cuda()).cpu().numpy(), ref) < 1e-5
        host = torch.from_numpy(x32).pin_memory()
        assert rel_fro(emb.apply(host).numpy(), ref) < 1e-5                           # streamed by column slabs
        assert rel_fro(emb.apply(x32.astype(np.float64)), ref) < TOL                  # same Theta for FP64 blocks
    b = rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "max_block_size": 64, "rng": "philox_tf32"}, _seed=9)
    refb = eo.gaussian_apply(x32.astype(np.float64), b.get_random_matrix())
    assert rel_fro(b.apply(x32), refb) < 1e-5
    # 'philox' (unrounded normals) keeps float32 blocks on the FP64 path
    e0 = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": "philox"}, _seed=5)
    assert rel_fro(e0.apply(x32), eo.gaussian_apply(x32.astype(np.float64), e0.get_random_matrix())) < TOL


def test_block_gaussian_embedding(rb):
    z = emb_golden()
    for tag in ("block_a", "block_b"):
        m, n, k, seed, mbs = (int(v) for v in z[tag + "__meta"])
        emb = rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "max_block_size": mbs}, _seed=seed)
        assert emb.block_sizes == list(z[tag + "__sizes"]) and emb.n_blocks == len(z[tag + "__sizes"])
        assert np.array_equal(np.asarray(emb.block_seeds, dtype=np.int64), z[tag + "__seeds"])
        assert np.array_equal(emb.get_random_matrix(), z[tag + "__theta"])
        assert np.array_equal(emb.get_block(1), z[tag + "__theta"][mbs:2 * mbs] if emb.n_blocks > 1 else None)
        assert rel_fro(emb.apply(z[tag + "__U"]), z[tag + "__Y"]) < TOL
    with pytest.raises(AssertionError):
        rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(8), options={"range_dim": 4})
    n, k = 5000, 70
    emb = rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "max_block_size": 32, "rng": "philox"}, _seed=5)
    x = np.random.RandomState(0).standard_normal((4, n))
    assert rel_fro(emb.apply(x), eo.gaussian_apply(x, emb.get_random_matrix())) < TOL


def test_vectorized_and_identity_embeddings(rb):
    k1, nv, k = 40, 6, 16
    inner = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(k1 * nv), options={"range_dim": k}, _seed=8)
    src = rb.DeviceVectorSpace(k1)
    vec = rb.EmbeddingVectorized(src, nv, inner)
    assert vec.range.dim == k and vec.options["range_dim"] == k
    U = np.random.RandomState(0).standard_normal((nv, k1))
    theta = eo.gaussian_random_matrix(k, k1 * nv, 8)
    ref = eo.vectorized_apply(U, lambda x: eo.gaussian_apply(x, theta))
    assert rel_fro(vec.apply(src.from_numpy(U)).to_numpy(), ref) < TOL
    assert vec.apply_adjoint(None) is None                                # embeddings.py:360-361
    with pytest.raises(AssertionError):
        vec.apply(src.from_numpy(U[:3]))
    ident = rb.IdentityEmbedding(source=src)
    assert ident.range.dim == k1
    assert np.array_equal(ident.apply(src.from_numpy(U)).to_numpy(), U)
    assert np.array_equal(ident.get_matrix(), np.eye(k1))


def test_operator_composition(rb):
    import scipy.sparse as sp
    n, k = 4096, 50
    A = sp.random(n, n, density=2e-3, random_state=0, format="csr")
    op = rb.MatrixOperator(A, source_id="S", range_id="S")
    emb = rb.SrhtEmbedding(source=op.range, options={"range_dim": k}, _seed=1)
    x = np.random.RandomState(0).standard_normal((6, n))
    got = (emb @ op).apply(op.source.from_numpy(x)).to_numpy()
    ref = oracle.srht(np.asarray((A @ x.T).T), k, seed=1)
    assert rel_fro(got, ref) < 1e-12
