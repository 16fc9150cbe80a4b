"""GPU parity of the dense-embedding GEMM (explicit Theta and on-the-fly Philox Theta)
against the oracle's (Theta @ U^T)^T on the same materialised Theta (1e-12 rel. Frobenius)."""
import numpy as np
import pytest
import torch

from oracle import embeddings_oracle as eo
from golden_util import emb_golden, rel_fro

pytestmark = pytest.mark.gpu
TOL64 = 1e-12


@pytest.fixture(scope="module")
def dn():
    import rla4mor_b200
    rla4mor_b200.lib()
    from rla4mor_b200 import dense
    return dense


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["gauss_small", "gauss_mid"])
def test_explicit_golden(dn, tag):
    z = emb_golden()
    y = dn.gauss_apply_explicit(_dev(z[tag + "__theta"]), _dev(z[tag + "__U"]))
    assert rel_fro(y.cpu().numpy(), z[tag + "__Y"]) < TOL64


@pytest.mark.parametrize("m,n,k", [(1, 16, 1), (3, 17, 5), (8, 64, 8), (64, 1000, 96), (65, 1024, 100),
                                   (130, 4099, 257), (200, 777, 33), (512, 20000, 300), (20, 50000, 2000)])
def test_explicit_vs_oracle(dn, m, n, k):
    rs = np.random.RandomState(m + n + k)
    theta = rs.standard_normal((k, n)) / np.sqrt(k)
    u = rs.standard_normal((m, n))
    y = dn.gauss_apply_explicit(_dev(theta), _dev(u))
    assert rel_fro(y.cpu().numpy(), eo.gaussian_apply(u, theta)) < TOL64


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("m,n,k", [(5, 100, 7), (64, 4096, 64), (130, 10007, 200), (512, 30000, 130)])
def test_rng_apply_matches_materialised_theta(dn, kind, m, n, k):
    u = np.random.RandomState(7).standard_normal((m, n))
    scale = 1.0 / np.sqrt(k)
    theta = dn.theta_materialize(1234, kind, scale, k, n).cpu().numpy()
    y = dn.embed_apply_rng(1234, kind, scale, k, _dev(u))
    assert rel_fro(y.cpu().numpy(), eo.gaussian_apply(u, theta)) < TOL64


TOL32 = 1e-5       # north_star: relative Frobenius error of the FP32 path


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("m,n,k", [(1, 37, 3), (5, 100, 7), (64, 4096, 64), (130, 10007, 200), (256, 65536 + 24, 300),
                                   (384, 30000, 130)])
def test_float32_blocks_on_tcgen05(dn, kind, m, n, k):
    """float32 blocks with a Theta exact in TF32 (csrc/gemm32.cu: tcgen05 kind::tf32, two-part split,
    FP64 accumulation of 64-term partial sums) against the FP64 product with the SAME materialised
    Theta; odd n, ragged m / k, single cluster (odd number of row tiles) and cluster pairs."""
    import torch
    u32 = np.random.RandomState(11).standard_normal((m, n)).astype(np.float32)
    scale = 1.0 / np.sqrt(k)
    theta = dn.theta_materialize(99, kind, scale, k, n).cpu().numpy()
    y = dn.embed_apply_rng(99, kind, scale, k, torch.from_numpy(u32).cuda())
    assert y.dtype == torch.float64
    ref = eo.gaussian_apply(u32.astype(np.float64), theta)
    assert rel_fro(y.cpu().numpy(), ref) < TOL32
    # the FP64 kernel on the widened block uses the same Theta
    y64 = dn.embed_apply_rng(99, kind, scale, k, torch.from_numpy(u32.astype(np.float64)).cuda())
    assert rel_fro(y64.cpu().numpy(), ref) < TOL64


def test_tf32_theta_is_exact_in_tf32(dn):
    t = dn.theta_materialize(3, 2, 1.0, 64, 4096).cpu().numpy().astype(np.float32)
    bits = t.view(np.uint32)
    assert np.all(bits & np.uint32(0x1FFF) == 0)                   # 13 low mantissa bits are zero
    t0 = dn.theta_materialize(3, 0, 1.0, 64, 4096).cpu().numpy()
    assert np.max(np.abs(t - t0) / np.maximum(np.abs(t0), 1e-30)) <= 2.0 ** -11 * 1.0001
    assert abs(t.std() - 1.0) < 5e-3


def test_float32_column_offsets_accumulate(dn):
    """Column slabs of a float32 block accumulate to the whole sketch (what streaming.py does)."""
    import torch
    m, n, k = 130, 3 * 4096, 96
    u = torch.from_numpy(np.random.RandomState(2).standard_normal((m, n)).astype(np.float32)).cuda()
    whole = dn.embed_apply_rng(7, 2, 0.5, k, u)
    acc = torch.zeros((m, k), dtype=torch.float64, device="cuda")
    for i, c0 in enumerate(range(0, n, 4096)):
        dn.embed_apply_rng(7, 2, 0.5, k, u[:, c0:c0 + 4096].contiguous(), col0=c0, out=acc, accumulate=i > 0)
    assert rel_fro(acc.cpu().numpy(), whole.cpu().numpy()) < 1e-6


def test_rng_statistics_and_determinism(dn):
    k, n = 256, 8192
    t0 = dn.theta_materialize(5, 0, 1.0, k, n).cpu().numpy()
    assert np.array_equal(t0, dn.theta_materialize(5, 0, 1.0, k, n).cpu().numpy())
    assert not np.array_equal(t0, dn.theta_materialize(6, 0, 1.0, k, n).cpu().numpy())
    assert abs(t0.mean()) < 5e-3 and abs(t0.std() - 1.0) < 5e-3
    assert abs((t0 ** 4).mean() - 3.0) < 0.05                      # Gaussian kurtosis
    assert abs(np.corrcoef(t0[0], t0[1])[0, 1]) < 0.05
    r = dn.theta_materialize(5, 1, 1.0, k, n).cpu().numpy()
    assert set(np.unique(r)) == {-1.0, 1.0} and abs(r.mean()) < 5e-3


def test_rng_blocks_tile_the_virtual_matrix(dn):
    # row / column offsets address sub-blocks of one virtual matrix (block-Gaussian rows,
    # row-sharded columns): sub-block results must add up to the full sketch
    m, n, k = 9, 6000, 50
    u = np.random.RandomState(3).standard_normal((m, n))
    full = dn.embed_apply_rng(9, 0, 0.5, k, _dev(u)).cpu().numpy()
    top = dn.embed_apply_rng(9, 0, 0.5, 20, _dev(u), row0=0).cpu().numpy()
    bot = dn.embed_apply_rng(9, 0, 0.5, 30, _dev(u), row0=20).cpu().numpy()
    assert rel_fro(np.hstack([top, bot]), full) < TOL64
    c = 2048                                                        # multiple of 16
    acc = dn.embed_apply_rng(9, 0, 0.5, k, _dev(u[:, :c]), col0=0)
    dn.embed_apply_rng(9, 0, 0.5, k, _dev(u[:, c:]), col0=c, out=acc, accumulate=True)
    assert rel_fro(acc.cpu().numpy(), full) < TOL64
    sub = dn.theta_materialize(9, 0, 0.5, 7, 11, row0=13, col0=5).cpu().numpy()
    whole = dn.theta_materialize(9, 0, 0.5, k, 64).cpu().numpy()
    assert np.array_equal(sub, whole[13:20, 5:16])


def test_norm_preservation_at_size(dn):
    m, n, k = 8, 2 ** 20, 2000
    u = torch.randn(m, n, dtype=torch.float64, device="cuda")
    y = dn.embed_apply_rng(0, 0, 1.0 / np.sqrt(k), k, u)
    ratio = (torch.linalg.norm(y, dim=1) / torch.linalg.norm(u, dim=1)).cpu().numpy()
    assert np.all(np.abs(ratio - 1.0) < 0.1)
    y2 = dn.embed_apply_rng(0, 0, 1.0 / np.sqrt(k), k, 2.0 * u)
    assert float(torch.linalg.norm(y2 - 2.0 * y) / torch.linalg.norm(y2)) < 1e-13
