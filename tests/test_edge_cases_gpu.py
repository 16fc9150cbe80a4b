"""Edge cases the domain offers: empty blocks, single vectors, k = 1, n = 1, odd n, float32 /
complex input to the dense embedding, more than 4096 distinct SRHT indices (two passes),
non-contiguous device blocks."""
import numpy as np
import pytest
import torch

import oracle
from oracle import embeddings_oracle as eo
from golden_util import rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


def test_empty_and_degenerate_blocks(rb):
    n, k = 300, 7
    s = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=1)
    g = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=1)
    p = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": "philox"}, _seed=1)
    for emb in (s, g, p):
        y = emb.apply(np.zeros((0, n)))
        assert y.shape == (0, k)
        y1 = emb.apply(np.ones((1, n)))
        assert y1.shape == (1, k) and np.all(np.isfinite(y1))
        assert np.array_equal(emb.apply(np.zeros((2, n))), np.zeros((2, k)))
    assert rb.srht(np.zeros((0, 16)), 3, seed=0).shape == (0, 3)
    x1 = np.random.RandomState(0).standard_normal((4, 1))                       # n = 1: d = 0
    assert rel_fro(rb.srht(x1, 5, seed=2), oracle.srht(x1, 5, seed=2)) < 1e-14
    x = np.random.RandomState(0).standard_normal((3, 50))
    assert rel_fro(rb.srht(x, 1, seed=2), oracle.srht(x, 1, seed=2)) < 1e-13    # k = 1


def test_two_pass_srht_and_noncontiguous(rb):
    n, k = 2 ** 15, 9000                                                         # > 4096 distinct indices
    x = np.random.RandomState(1).standard_normal((3, n))
    assert rel_fro(rb.srht(x, k, seed=4), oracle.srht(x, k, seed=4)) < 1e-12
    big = torch.randn(8, 3 * 5000, dtype=torch.float64, device="cuda")
    view = big[::2, 5000:10000]                                                  # strided rows, offset columns
    assert rel_fro(rb.srht(view, 40, seed=1).cpu().numpy(), oracle.srht(view.cpu().numpy(), 40, seed=1)) < 1e-12
    g = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(5000), options={"range_dim": 33, "rng": "philox"}, _seed=2)
    assert rel_fro(g.apply(view).cpu().numpy(), eo.gaussian_apply(view.cpu().numpy(), g.get_random_matrix())) < 1e-12


def test_dense_embedding_odd_n_float32_complex(rb):
    n, k = 1001, 40                                                              # odd n: TMA-unfriendly stride
    x = np.random.RandomState(2).standard_normal((5, n))
    g = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=3)
    theta = eo.gaussian_random_matrix(k, n, 3)
    assert rel_fro(g.apply(x), eo.gaussian_apply(x, theta)) < 1e-12
    p = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": "philox_rademacher"}, _seed=3)
    assert rel_fro(p.apply(x), eo.gaussian_apply(x, p.get_random_matrix())) < 1e-12
    y32 = g.apply(x.astype(np.float32))                                          # widened on the host side
    assert rel_fro(y32, eo.gaussian_apply(x, theta)) < 1e-5
    # the raw C ABI also accepts the unaligned operands (plain-FMA fallback kernel)
    from rla4mor_b200 import dense
    th = torch.from_numpy(theta).cuda()
    xd = torch.from_numpy(x).cuda()
    from rla4mor_b200._lib import lib, check, stream_ptr
    out = torch.empty((5, k), dtype=torch.float64, device="cuda")
    ws = torch.empty(lib().rla_gemm_workspace_bytes(5, k, n), dtype=torch.uint8, device="cuda")
    check(lib().rla_gauss_apply_explicit_f64(th.data_ptr(), k, n, n, xd.data_ptr(), 5, n, out.data_ptr(), k,
                                             ws.data_ptr(), ws.numel(), stream_ptr()), "explicit odd n")
    assert rel_fro(out.cpu().numpy(), eo.gaussian_apply(x, theta)) < 1e-12


def test_srht_embedding_complex_and_float32(rb):
    n, k = 777, 20
    rs = np.random.RandomState(3)
    xc = rs.standard_normal((2, n)) + 1j * rs.standard_normal((2, n))
    s = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "dtype": complex}, _seed=1)
    y = s.apply(xc)
    assert y.dtype == np.complex128 and rel_fro(y, oracle.srht(xc, k, seed=1)) < 1e-12
    xf = rs.standard_normal((3, n)).astype(np.float32)
    yf = s.apply(xf)
    assert yf.dtype == np.float32 and rel_fro(yf, oracle.srht(xf.astype(np.float64), k, seed=1)) < 1e-5


def test_error_codes_from_the_abi(rb):
    from rla4mor_b200._lib import lib, RlaError, check
    import ctypes
    x = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
    plan = rb.get_plan(64, 8, 0, torch.float64, x.device)
    y = torch.empty(2, 8, dtype=torch.float64, device="cuda")
    # workspace too small -> RLA_ERR_WORKSPACE (-3), message available
    rc = lib().rla_srht_apply_f64(plan._handle, x.data_ptr(), 2, 64, 1.0, y.data_ptr(), 8, x.data_ptr(), 8, None)
    assert rc == -3 and b"workspace" in lib().rla_last_error()
    with pytest.raises(RlaError):
        check(lib().rla_embed_apply_rng_f64(0, 0, 1.0, 0, 4, 8, 64, x.data_ptr(), 2, 64, y.data_ptr(), 8, 0,
                                            x.data_ptr(), 1 << 20, None), "col0 not multiple of 16")


def test_single_vector_odd_n_on_the_fly_rng(rb):
    """A (1, odd n) block on the in-kernel RNG path (advisor finding, round 1): the padded copy
    must hand its real leading dimension to the TMA kernel; also through EmbeddingVectorized,
    which always sketches one long vector."""
    import torch
    from rla4mor_b200 import dense
    for n in (4097, 33, 1):
        x = torch.randn(1, n, dtype=torch.float64, device="cuda")
        for mode in ("philox", "philox_rademacher"):
            g = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": 24, "rng": mode}, _seed=9)
            y = g.apply(x)
            theta = torch.from_numpy(g.get_random_matrix()).cuda()
            ref = x @ theta.T
            assert float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref)) < 1e-12, (n, mode)
    k1, nv = 7, 5                                   # vec(U) has 35 entries: odd
    inner = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(k1 * nv), options={"range_dim": 6, "rng": "philox"}, _seed=3)
    vec = rb.EmbeddingVectorized(rb.DeviceVectorSpace(k1), nv, inner, options={})
    W = torch.randn(nv, k1, dtype=torch.float64, device="cuda")
    y = vec.apply(W)
    ref = W.T.reshape(1, -1) @ torch.from_numpy(inner.get_random_matrix()).cuda().T
    assert float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref)) < 1e-12
    b = rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(35), options={"range_dim": 10, "max_block_size": 4, "rng": "philox"}, _seed=3)
    yb = b.apply(W.T.reshape(1, -1))
    refb = W.T.reshape(1, -1) @ torch.from_numpy(b.get_random_matrix()).cuda().T
    assert float(torch.linalg.norm(yb - refb) / torch.linalg.norm(refb)) < 1e-12


def test_theta_row_sharded_front_end_matches_single_gpu(rb):
    """k split over (simulated) ranks: every rank's slice equals the corresponding columns of the
    single-GPU sketch bit for bit (same Philox counters / same MT19937 blocks)."""
    import torch
    from rla4mor_b200 import dense, sharding
    x = torch.randn(9, 3000, dtype=torch.float64, device="cuda")
    k, seed, world = 50, 5, 4
    full = dense.embed_apply_rng(seed, 0, 1.0 / np.sqrt(k), k, x)
    got = torch.cat([sharding.gaussian_theta_row_sharded(x, k, seed, r, world, gather=False)[0] for r in range(world)], dim=1)
    assert float(torch.linalg.norm(got - full) / torch.linalg.norm(full)) < 1e-14
    emb = rb.BlockGaussianEmbedding(source=rb.DeviceVectorSpace(3000), options={"range_dim": 50, "max_block_size": 8}, _seed=11)
    ref = emb.apply(x)
    parts = [sharding.block_gaussian_theta_row_sharded(emb, x, r, world, gather=False)[0] for r in range(world)]
    assert torch.equal(torch.cat(parts, dim=1), ref)
