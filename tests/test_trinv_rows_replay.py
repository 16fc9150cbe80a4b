"""CPU replay of trinv_rows_kernel (csrc/factor.cu): row i of T = R^-1 from T R = I, right-looking --
the running right-hand side acc_j lives in the lanes (column j in lane j mod 32, register j div 32),
step l finalises t_l = acc_l / R_ll and subtracts t_l R[l, j] from the columns j > l, 32-column
chunks in order, starting at the chunk and lane of the diagonal."""
import numpy as np
import pytest


def trinv_row(R, i, nc):
    r = R.shape[0]
    acc = np.zeros((nc, 32))
    tv = np.zeros((nc, 32))
    acc[i // 32, i % 32] = 1.0
    ci = i >> 5
    for c0 in range(nc):
        if c0 < ci or 32 * c0 >= r:
            continue
        l0 = (i & 31) if c0 == ci else 0
        for ll in range(l0, min(32, r - 32 * c0)):
            l = 32 * c0 + ll
            t = acc[c0, ll] * (1.0 / R[l, l])
            tv[c0, ll] = t
            for c in range(c0, nc):
                for lane in range(32):
                    j = lane + 32 * c
                    if j > l and j < r:
                        acc[c, lane] -= t * R[l, j]
    return tv.reshape(-1)[:r]


@pytest.mark.parametrize("r", [1, 5, 31, 32, 33, 70, 128, 129])
def test_row_oriented_triangular_inverse(r):
    R = np.triu(np.random.RandomState(r).standard_normal((r, r))) / np.sqrt(r) + 3.0 * np.eye(r)
    nc = 4 if r <= 128 else 8
    T = np.vstack([trinv_row(R, i, nc) for i in range(r)])
    assert np.all(np.tril(T, -1) == 0.0)
    assert np.linalg.norm(T @ R - np.eye(r)) < 1e-13 * r
    assert np.linalg.norm(T - np.linalg.inv(R)) / np.linalg.norm(T) < 1e-13
