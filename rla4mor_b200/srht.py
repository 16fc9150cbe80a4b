"""B200 drop-in for the reference's `rla/srht.py`: `srht`, `fht_oop`, `fht_ip`.

Same names, argument meaning and assertion behaviour as rla/srht.py:99-177.
Inputs may be NumPy arrays (copied to the GPU, result returned as NumPy -- the
reference's calling convention) or CUDA `torch.Tensor`s (result stays on the
device).  Layout is the reference's: `(m, n)`, one vector per row (srht.py:142).

The Rademacher signs and the row indices are drawn on the host by NumPy's
legacy `RandomState`, literally as srht.py:162-163 draws them, so they are
bit-identical to the reference's; they are cached per (seed, n, k) because the
reference's per-call redraw costs seconds at n = 2**24.
All arithmetic runs in the CUDA library (`csrc/srht.cu`, `csrc/fwht.cu`);
there is no CPU fallback.
"""
import ctypes
import weakref
from collections import OrderedDict

import numpy as np

from ._lib import c_void_p, check, lib, require_cuda, stream_ptr


def draw_signs_and_indices(n, k, seed):
    """The two draws of srht.py:161-163 (two RandomStates from the same seed).

    `RandomState.choice(range(2**d), k)` and `choice(2**d, k)` consume the
    stream identically (both are `randint(0, 2**d, k)`); the integer form avoids
    building a 2**24-element list."""
    d = int(np.ceil(np.log2(n)))
    rademacher = np.random.RandomState(seed).choice([-1, 1], (n), True)
    sampling = np.random.RandomState(seed).choice(2 ** d, k, True)
    return rademacher, sampling


class SrhtPlan:
    """Device-side descriptor of one SRHT operator (signs, indices, n, k)."""

    def __init__(self, n, k, signs, sampling, dtype, device):
        torch = require_cuda()
        self.n, self.k = int(n), int(k)
        self.d = int(np.ceil(np.log2(n)))
        self.dtype = dtype
        self.device = device
        signs8 = np.ascontiguousarray(signs, dtype=np.int8)
        idx64 = np.ascontiguousarray(sampling, dtype=np.int64)
        assert signs8.shape == (self.n,) and idx64.shape == (self.k,)
        self.signs_host, self.idx_host = signs8, idx64
        handle = c_void_p()
        elem = 8 if dtype == torch.float64 else 4
        check(lib().rla_srht_plan_create(ctypes.byref(handle), signs8.ctypes.data, self.n,
                                         idx64.ctypes.data, self.k, elem), "rla_srht_plan_create")
        self._handle = handle
        self._finalizer = weakref.finalize(self, lib().rla_srht_plan_destroy, handle)
        nbytes = lib().rla_srht_plan_device_bytes(handle)
        with torch.cuda.device(device):
            self.image = torch.empty(nbytes, dtype=torch.uint8, device=device)
            check(lib().rla_srht_plan_upload(handle, self.image.data_ptr(), stream_ptr()), "rla_srht_plan_upload")
        self._signs_dev = None
        self._idx_dev = None

    @property
    def signs_dev(self):
        if self._signs_dev is None:
            import torch
            self._signs_dev = torch.from_numpy(self.signs_host).to(self.device)
        return self._signs_dev

    @property
    def idx_dev(self):
        if self._idx_dev is None:
            import torch
            self._idx_dev = torch.from_numpy(self.idx_host).to(self.device)
        return self._idx_dev

    @property
    def passes(self):
        return lib().rla_srht_plan_passes(self._handle)

    def workspace(self, m):
        """Scratch for one call from torch's caching allocator (stream-safe, see dense._workspace)."""
        import torch
        need = lib().rla_srht_workspace_bytes(self._handle, m)
        return torch.empty(max(need, 16), dtype=torch.uint8, device=self.device)

    def apply(self, x, scale=None, out=None):
        """x: CUDA tensor (m, n) with unit inner stride -> (m, k) sketch."""
        import torch
        assert x.is_cuda and x.dim() == 2 and x.shape[1] == self.n and x.dtype == self.dtype
        if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < self.n):
            x = x.contiguous()
        m = x.shape[0]
        if out is None:
            out = torch.empty((m, self.k), dtype=self.dtype, device=x.device)
        if scale is None:
            scale = 1.0 / np.sqrt(self.k) if self.k else 1.0
        if m == 0 or self.k == 0:
            return out
        with torch.cuda.device(x.device):
            ws = self.workspace(m)
            fn = lib().rla_srht_apply_f64 if self.dtype == torch.float64 else lib().rla_srht_apply_f32
            ldx = x.stride(0) if m > 1 else max(x.stride(0), self.n)
            check(fn(self._handle, x.data_ptr(), m, ldx, float(scale), out.data_ptr(), out.stride(0),
                     ws.data_ptr(), ws.numel(), stream_ptr()), "rla_srht_apply")
        return out


_PLAN_CACHE = OrderedDict()
_PLAN_CACHE_MAX = 8


def get_plan(n, k, seed, dtype, device):
    """Cached plan for an integer seed; `seed=None` draws fresh OS entropy per
    call, as the reference does (srht.py:146-147)."""
    import torch
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(n), int(k), None if seed is None else int(seed), dtype, device.index)
    if seed is not None and key in _PLAN_CACHE:
        _PLAN_CACHE.move_to_end(key)
        return _PLAN_CACHE[key]
    signs, sampling = draw_signs_and_indices(n, k, seed)
    plan = SrhtPlan(n, k, signs, sampling, dtype, device)
    if seed is not None:
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
            _PLAN_CACHE.popitem(last=False)
    return plan


def _to_device(x):
    """(tensor on the GPU, was_numpy)."""
    torch = require_cuda()
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).cuda(), True
    assert isinstance(x, torch.Tensor), "expected a numpy array or a torch tensor"
    assert x.is_cuda, "torch inputs must live on the GPU (there is no CPU path)"
    return x, False


def _split_complex(t):
    import torch
    return torch.cat([t.real, t.imag], dim=0).contiguous()


def srht(x, k, seed=None, nthreads=4):
    """SRHT of every row of x -- signature and semantics of rla/srht.py:136-177.

    x : (n,) or (m, n), float64 / complex128 (float32 / complex64 also accepted).
    Returns sqrt(2**d / k) * (H_norm (r * x_pad))[:, s]  with d = ceil(log2 n).
    `nthreads` is accepted for signature compatibility and ignored.
    """
    import torch
    assert x.ndim <= 2                                  # srht.py:155
    xt, was_numpy = _to_device(x)
    one_d = xt.dim() == 1
    if one_d:
        xt = xt.reshape(1, -1)                          # srht.py:157-158
    m, n = xt.shape
    cplx = xt.is_complex()
    real_dtype = {torch.float64: torch.float64, torch.complex128: torch.float64,
                  torch.float32: torch.float32, torch.complex64: torch.float32}.get(xt.dtype)
    assert real_dtype is not None, f"unsupported dtype {xt.dtype}"
    plan = get_plan(n, k, seed, real_dtype, xt.device)
    if cplx:                                            # real and imaginary parts separately (srht.py:126-127)
        y2 = plan.apply(_split_complex(xt))
        y = torch.complex(y2[:m], y2[m:])
    else:
        y = plan.apply(xt)
    if one_d:
        y = y.reshape(-1)                               # srht.py:174-175
    return y.cpu().numpy() if was_numpy else y


def _fwht_device(a2, post_scale, out=None):
    """a2: CUDA tensor (m, 2**d) real, contiguous rows."""
    import torch
    m, n = a2.shape
    if out is None:
        out = torch.empty_like(a2)
    if m == 0:
        return out
    fn = lib().rla_fwht_f64 if a2.dtype == torch.float64 else lib().rla_fwht_f32
    with torch.cuda.device(a2.device):
        check(fn(a2.data_ptr(), m, n, a2.stride(0), out.data_ptr(), out.stride(0), float(post_scale),
                 stream_ptr()), "rla_fwht")
    return out


def _fht(a, inplace):
    import torch
    d = np.log2(a.shape[-1])
    assert d % 1 == 0                                   # srht.py:110-111 / 122-123
    assert a.ndim <= 2                                  # srht.py:112 / 124
    at, was_numpy = _to_device(a)
    if inplace and not was_numpy:
        assert at.is_contiguous(), "fht_ip needs a contiguous tensor"
    shape = at.shape
    a2 = at.reshape(1, -1) if at.dim() == 1 else at
    if a2.stride(-1) != 1:
        a2 = a2.contiguous()
    post = 1.0 / (2 ** (int(d) / 2))                    # srht.py:36
    if at.is_complex():
        m = a2.shape[0]
        r2 = _fwht_device(_split_complex(a2), post)
        res = torch.complex(r2[:m], r2[m:])
    elif inplace and not was_numpy:
        res = _fwht_device(a2, post, out=a2)
    else:
        res = _fwht_device(a2, post)
    res = res.reshape(shape)
    if inplace:
        if was_numpy:
            a[...] = res.cpu().numpy()
        elif res.data_ptr() != at.data_ptr():
            at.copy_(res)
        return None
    return res.cpu().numpy() if was_numpy else res


def fht_oop(a, nthreads=1):
    """Out-of-place normalised FWHT along the last axis (rla/srht.py:121-134)."""
    return _fht(a, inplace=False)


def fht_ip(a):
    """In-place normalised FWHT along the last axis (rla/srht.py:99-118)."""
    return _fht(a, inplace=True)


# The reference also defines four numba loop-order variants of the in-place transform
# (rla/srht.py:14-96).  `_fht_2d` and `_fht_2d_sequential` are unreachable from its public
# functions; all four compute the same normalised in-place
# FWHT along the last axis, so they are aliases of `fht_ip` here.
_fht_1d = fht_ip
_fht_2d = fht_ip
_fht_2d_sequential = fht_ip
_fht_2d_parallel = fht_ip
