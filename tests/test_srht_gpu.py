"""GPU parity of the fused SRHT kernel (through the C ABI) against the CPU oracle
and the reference-generated golden fixtures.  Tolerance: relative Frobenius error
1e-12 in FP64 (north_star); signs and indices are bit-exact by construction and
checked explicitly."""
import numpy as np
import pytest
import torch

import oracle
from golden_util import srht_cases, rel_fro

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


@pytest.mark.parametrize("case", list(srht_cases()), ids=lambda c: c["name"])
def test_srht_golden(rb, case):
    y = rb.srht(case["x"], case["k"], seed=case["seed"])
    assert isinstance(y, np.ndarray) and y.shape == case["y"].shape and y.dtype == case["y"].dtype
    assert rel_fro(y, case["y"]) < TOL64


@pytest.mark.parametrize("case", list(srht_cases()), ids=lambda c: c["name"])
def test_signs_indices_bit_exact(rb, case):
    r, s = rb.draw_signs_and_indices(case["n"], case["k"], case["seed"])
    assert np.array_equal(r.astype(np.int8), case["signs"])
    assert np.array_equal(s.astype(np.int64), case["sampling"])


@pytest.mark.parametrize("m,n,k,seed", [
    (1, 2, 1, 0), (3, 5, 7, 1), (2, 4095, 33, 2), (2, 4096, 600, 3), (3, 4097, 1000, 4),
    (5, 12289, 129, 5), (2, 2 ** 15, 1500, 6), (7, 2 ** 16 + 3, 4000, 7), (1, 2 ** 18, 4500, 8),
    (130, 1000, 64, 9), (2, 2 ** 20, 10000, 10),
])
def test_srht_vs_oracle(rb, m, n, k, seed):
    x = np.random.RandomState(100 + seed).standard_normal((m, n))
    y = rb.srht(x, k, seed=seed)
    ref = oracle.srht(x, k, seed=seed)
    assert rel_fro(y, ref) < TOL64


def test_srht_torch_input_stays_on_device_and_input_untouched(rb):
    x = torch.randn(4, 9000, dtype=torch.float64, device="cuda")
    x0 = x.clone()
    y = rb.srht(x, 50, seed=3)
    assert y.is_cuda and y.shape == (4, 50)
    assert torch.equal(x, x0)                      # srht never mutates x (srht.py:156)
    ref = oracle.srht(x0.cpu().numpy(), 50, seed=3)
    assert rel_fro(y.cpu().numpy(), ref) < TOL64


def test_srht_strided_rows_and_unaligned(rb):
    big = torch.randn(6, 5001, dtype=torch.float64, device="cuda")
    x = big[:, 1:4098]                             # odd offset: scalar-load path
    y = rb.srht(x, 77, seed=12)
    ref = oracle.srht(x.cpu().numpy(), 77, seed=12)
    assert rel_fro(y.cpu().numpy(), ref) < TOL64


def test_srht_float32(rb):
    x = np.random.RandomState(5).standard_normal((3, 20000)).astype(np.float32)
    y = rb.srht(x, 300, seed=2)
    assert y.dtype == np.float32
    ref = oracle.srht(x.astype(np.float64), 300, seed=2)
    assert rel_fro(y, ref) < TOL32


def test_srht_deterministic_and_seed_none(rb):
    x = np.random.RandomState(1).standard_normal((2, 3000))
    assert np.array_equal(rb.srht(x, 40, seed=5), rb.srht(x, 40, seed=5))
    a, b = rb.srht(x, 40), rb.srht(x, 40)          # seed=None: fresh entropy each call
    assert not np.array_equal(a, b)


def test_srht_linearity_and_norm_preservation_at_size(rb):
    # size-independent properties on a block the oracle would take too long on
    m, n, k = 4, 2 ** 22, 4000
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    z = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    yx, yz = rb.srht(x, k, seed=0), rb.srht(z, k, seed=0)
    yl = rb.srht(2.0 * x - 3.0 * z, k, seed=0)
    assert float(torch.linalg.norm(yl - (2.0 * yx - 3.0 * yz)) / torch.linalg.norm(yl)) < 1e-12
    ratio = (torch.linalg.norm(yx, dim=1) / torch.linalg.norm(x, dim=1)).cpu().numpy()
    assert np.all(np.abs(ratio - 1.0) < 0.1)       # E||Theta x||^2 = ||x||^2, k = 4000
    # closed form of one output entry, evaluated independently on the device in FP64
    r, s = rb.draw_signs_and_indices(n, k, 0)
    j = torch.arange(n, device="cuda")
    for i in (0, 1234, 3999):
        v = torch.bitwise_and(j, int(s[i]))
        par = torch.zeros_like(v)
        for b in range(22):
            par ^= (v >> b) & 1
        sgn = (1 - 2 * par).to(torch.float64) * torch.from_numpy(r).to("cuda").to(torch.float64)
        ref = (x * sgn).sum(dim=1) / np.sqrt(k)
        assert float(torch.linalg.norm(yx[:, i] - ref) / torch.linalg.norm(ref)) < 1e-11


def test_complex_input(rb):
    rs = np.random.RandomState(3)
    x = rs.standard_normal((2, 777)) + 1j * rs.standard_normal((2, 777))
    y = rb.srht(x, 20, seed=1)
    assert y.dtype == np.complex128
    assert rel_fro(y, oracle.srht(x, 20, seed=1)) < TOL64


def test_bad_arguments(rb):
    with pytest.raises(AssertionError):
        rb.srht(np.zeros((2, 2, 2)), 3, seed=0)


def test_warp_specialised_kernel_vs_oracle(rb, monkeypatch):
    """Small problems take the single-role kernel by default; force the warp-specialised one
    (the kernel the full-size blocks run on) and compare with the oracle: ragged last tile,
    several passes (k > 4096 distinct indices), unaligned rows, float32."""
    monkeypatch.setenv("RLA_SRHT_VARIANT", "0")
    for m, n, k, seed in [(3, 2 ** 16 + 3, 4000, 7), (2, 2 ** 18, 9000, 8), (5, 2 ** 14, 100, 9), (1, 70000, 1, 10)]:
        x = np.random.RandomState(200 + seed).standard_normal((m, n))
        assert rel_fro(rb.srht(x, k, seed=seed), oracle.srht(x, k, seed=seed)) < TOL64
    big = torch.randn(3, 2 ** 15 + 9, dtype=torch.float64, device="cuda")
    xs = big[:, 1:2 ** 15 + 2]                     # odd offset: staged (scalar) tile loads
    assert rel_fro(rb.srht(xs, 300, seed=4).cpu().numpy(), oracle.srht(xs.cpu().numpy(), 300, seed=4)) < TOL64
    x32 = np.random.RandomState(11).standard_normal((3, 2 ** 16)).astype(np.float32)
    assert rel_fro(rb.srht(x32, 500, seed=2), oracle.srht(x32.astype(np.float64), 500, seed=2)) < TOL32


def test_warp_specialised_kernel_matches_single_role_kernel(rb, monkeypatch):
    x = torch.randn(6, 2 ** 19, dtype=torch.float64, device="cuda")
    monkeypatch.setenv("RLA_SRHT_VARIANT", "1")
    y1 = rb.srht(x, 4000, seed=0)
    monkeypatch.setenv("RLA_SRHT_VARIANT", "0")
    y0 = rb.srht(x, 4000, seed=0)
    assert float(torch.linalg.norm(y0 - y1) / torch.linalg.norm(y1)) < 1e-13
