"""GPU parity of fht_oop / fht_ip (rla/srht.py:99-134) through the C ABI."""
import numpy as np
import pytest
import torch

import oracle
from golden_util import fht_cases, rel_fro

pytestmark = pytest.mark.gpu
TOL64 = 1e-12


@pytest.fixture(scope="module")
def rb():
    import rla4mor_b200
    rla4mor_b200.lib()
    return rla4mor_b200


@pytest.mark.parametrize("case", list(fht_cases()), ids=lambda c: c["name"])
def test_fht_golden(rb, case):
    a = case["a"]
    out = rb.fht_oop(a)
    assert out.shape == a.shape and out.dtype == a.dtype
    assert rel_fro(out, case["oop"]) < TOL64
    b = a.copy()
    assert rb.fht_ip(b) is None
    assert rel_fro(b, case["ip"]) < TOL64


@pytest.mark.parametrize("m,d", [(1, 1), (5, 3), (3, 11), (2, 12), (3, 13), (2, 17), (1, 22), (2, 23), (1, 25), (70, 6)])
def test_fht_vs_oracle(rb, m, d):
    a = np.random.RandomState(d).standard_normal((m, 2 ** d))
    assert rel_fro(rb.fht_oop(a), oracle.fht_oop(a)) < TOL64


def test_fht_device_tensor_inplace_and_involution(rb):
    a = torch.randn(3, 2 ** 14, dtype=torch.float64, device="cuda")
    a0 = a.clone()
    b = rb.fht_oop(a)
    assert torch.equal(a, a0) and b.is_cuda
    rb.fht_ip(a)
    assert torch.equal(a, b)
    rb.fht_ip(a)                                   # H_norm is an involution
    assert float(torch.linalg.norm(a - a0) / torch.linalg.norm(a0)) < 1e-14


def test_fht_identity_is_sylvester_matrix(rb):
    n = 64
    H = rb.fht_oop(np.eye(n)) * np.sqrt(n)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    par = np.vectorize(lambda v: bin(v).count("1") & 1)(i & j)
    assert np.array_equal(H, 1.0 - 2.0 * par)      # exact +-1, natural order


def test_fht_float32_and_errors(rb):
    a = np.random.RandomState(0).standard_normal((2, 4096)).astype(np.float32)
    out = rb.fht_oop(a)
    assert out.dtype == np.float32
    assert rel_fro(out, oracle.fht_oop(a.astype(np.float64))) < 1e-5
    with pytest.raises(AssertionError):
        rb.fht_oop(np.zeros((2, 12)))
    with pytest.raises(AssertionError):
        rb.fht_ip(np.zeros((2, 2, 4)))
