"""CPU restatement of /root/reference/rla/srht.py -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows.  The butterfly itself runs
either in `oracle/fwht_oracle.c` (gcc + OpenMP over rows, one row per thread --
the same parallel decomposition as the reference's numba `prange`,
srht.py:93-96) or, when the C library is not built, in a NumPy reshape
formulation.  Both apply the stages in the reference's order h = 1, 2, 4, ...
(srht.py:27-35) and divide by 2**(d/2) afterwards (srht.py:36), so for real
float64 input they reproduce the reference bit for bit (checked in
tests/test_oracle_golden.py against fixtures produced by the reference itself).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def _clib():
    """Load oracle/_build/libfwht_oracle.so if it has been built (optional)."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "libfwht_oracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.oracle_fwht_rows_f64.argtypes = [ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_int64, ctypes.c_int]
            lib.oracle_fwht_rows_f64.restype = None
            lib.oracle_max_threads.restype = ctypes.c_int
            _CLIB = lib
        else:
            _CLIB = False
    return _CLIB


def oracle_threads():
    lib = _clib()
    return int(lib.oracle_max_threads()) if lib else 1


def _butterflies_numpy(a):
    """Unnormalised in-place radix-2 WHT along the last axis (srht.py:27-35)."""
    m, n = a.shape
    h = 1
    while h < n:
        v = a.reshape(m, n // (2 * h), 2, h)
        x = v[:, :, 0, :].copy()
        y = v[:, :, 1, :]
        v[:, :, 0, :] = x + y
        v[:, :, 1, :] = x - y
        h *= 2


def fht_ip(a, nthreads=0):
    """In-place normalised FWHT (srht.py:99-118).

    1-D: `_fht_1d` (srht.py:113-114); 2-D: every row independently
    (`_fht_2d_parallel`, srht.py:93-96,117-118).  The `shape[1] == 1` branch of
    the reference (srht.py:115-116) transforms the length-1 row `a[0]`, which is
    the identity; a (m, 1) array is therefore left untouched here as well.
    """
    d = np.log2(a.shape[-1])
    assert d % 1 == 0           # srht.py:110-111
    assert a.ndim <= 2          # srht.py:112
    if a.shape[-1] == 1:
        return
    if np.iscomplexobj(a):      # c16 signature, srht.py:14: real and imag separately
        re = np.ascontiguousarray(a.real)
        im = np.ascontiguousarray(a.imag)
        fht_ip(re, nthreads)
        fht_ip(im, nthreads)
        a[...] = re + 1j * im
        return
    assert a.dtype == np.float64 and a.flags.c_contiguous
    v = a.reshape(1, -1) if a.ndim == 1 else a
    lib = _clib()
    if lib:
        lib.oracle_fwht_rows_f64(v.ctypes.data, v.shape[0], v.shape[1], int(nthreads))
    else:
        _butterflies_numpy(v)
    v /= 2 ** (int(d) / 2)      # srht.py:36  (a /= 2**(d/2))


def fht_oop(a, nthreads=1):
    """Out-of-place normalised FWHT (srht.py:121-134, the `ffht is None` branch)."""
    d = np.log2(a.shape[-1])
    assert d % 1 == 0           # srht.py:122-123
    assert a.ndim <= 2          # srht.py:124
    result = np.array(a, copy=True, order="C")   # srht.py:132
    fht_ip(result)              # srht.py:133
    return result


def rademacher_signs(n, seed):
    """srht.py:162 / embeddings.py:201 -- int64 array of +-1."""
    return np.random.RandomState(seed).choice([-1, 1], (n), True)


def sampling_indices(n, k, seed):
    """srht.py:161,163 / embeddings.py:198,202 -- k draws from range(2**d), with
    replacement, from a *second* RandomState built from the same seed."""
    d = int(np.ceil(np.log2(n)))
    # RandomState.choice(range(N)) == RandomState.choice(N) (same stream: both
    # reduce to randint(0, N, size)); np.arange avoids a 2**24-element Python range.
    return np.random.RandomState(seed).choice(2 ** d, k, True)


def srht(x, k, seed=None, nthreads=4):
    """SRHT of every row of x (srht.py:136-177)."""
    assert x.ndim <= 2                              # :155
    y = x.copy()                                    # :156
    if x.ndim == 1:
        y = y.reshape(1, -1)                        # :157-158
    n = y.shape[1]                                  # :160
    d = int(np.ceil(np.log2(n)))                    # :161
    rademacher = np.random.RandomState(seed).choice([-1, 1], (n), True)      # :162
    sampling = np.random.RandomState(seed).choice(2 ** d, k, True)           # :163
    y = rademacher * y                              # :165
    y = np.append(y, np.zeros((y.shape[0], 2 ** d - n)), axis=1)             # :167
    y = fht_oop(y, nthreads)                        # :169
    y = np.sqrt((2 ** d) / k) * y[:, sampling]      # :171
    if x.ndim == 1:
        y = y.reshape(-1)                           # :174-175
    return y


def srht_with(x, k, signs, sampling):
    """`srht` with the signs / row indices supplied (same arithmetic as
    srht.py:165-171); used to check a device sketch against the oracle on the
    *same materialised* signs and indices."""
    y = np.array(x, copy=True)
    one_d = y.ndim == 1
    if one_d:
        y = y.reshape(1, -1)
    n = y.shape[1]
    d = int(np.ceil(np.log2(n)))
    y = signs * y
    y = np.append(y, np.zeros((y.shape[0], 2 ** d - n)), axis=1)
    y = fht_oop(y)
    y = np.sqrt((2 ** d) / k) * y[:, sampling]
    return y.reshape(-1) if one_d else y


def srht_closed_form(x, k, seed):
    """Brute-force closed form, small sizes only:
    out[c, i] = k**-0.5 * sum_j (-1)**popcount(s_i & j) * r_j * x[c, j]."""
    x2 = np.atleast_2d(np.asarray(x))
    n = x2.shape[1]
    r = rademacher_signs(n, seed)
    s = sampling_indices(n, k, seed)
    j = np.arange(n)
    out = np.empty((x2.shape[0], k), dtype=np.result_type(x2.dtype, np.float64))
    for i in range(k):
        par = np.array([bin(int(s[i]) & int(t)).count("1") & 1 for t in j])
        out[:, i] = (x2 * (r * (1 - 2 * par))).sum(axis=1) / np.sqrt(k)
    return out if np.ndim(x) == 2 else out[0]
