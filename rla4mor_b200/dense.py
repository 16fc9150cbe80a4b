"""Device-level wrappers of the dense-embedding kernels (csrc/gemm.cu, csrc/rng.cuh).

Y(m, k) = U(m, n) Theta(k, n)^T in FP64 on the tensor pipe, with Theta explicit
(`gauss_apply_explicit`, the reference's `(Theta @ (QU)^T)^T`, rla/embeddings.py:250-254)
or generated on the fly from the counter-based RNG (`embed_apply_rng`), plus the export
of what that generator produces (`theta_materialize`, parity tooling).
All inputs are CUDA tensors; there is no CPU path.
"""
import torch

from ._lib import check, lib, stream_ptr

KIND_NORMAL = 0
KIND_RADEMACHER = 1
KIND_NORMAL_TF32 = 2      # the normals of kind 0 rounded to TF32: exact operands of tcgen05.mma kind::tf32

def _workspace(nbytes, device):
    """Scratch for one call, taken from torch's caching allocator: it is tied to the current
    stream, so concurrent calls on different streams never share a buffer."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _rows(x):
    """A 2-D CUDA tensor with unit inner stride (copy only if needed)."""
    assert x.is_cuda and x.dim() == 2
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def _ld(x):
    """Leading dimension handed to the kernels.  A single row has no meaningful stride of its own
    (torch reports anything for a size-1 dimension): use the stride when it can be one (a row of
    a padded buffer, as `_tma_friendly` builds for odd n), else the row length."""
    if x.shape[0] > 1:
        return max(x.stride(0), x.shape[1])
    return x.stride(0) if x.stride(0) >= x.shape[1] else x.shape[1]


def _tma_friendly(x):
    """TMA needs a 16-byte aligned base and an even row stride; copy into a padded
    buffer otherwise (odd n)."""
    if x.data_ptr() % 16 == 0 and _ld(x) % 2 == 0:
        return x
    m, n = x.shape
    buf = torch.zeros((m, n + (n & 1)), dtype=x.dtype, device=x.device)
    buf[:, :n] = x
    return buf[:, :n]


def gauss_apply_explicit(theta, u, out=None):
    """(m, n) x (k, n)^T -> (m, k), float64."""
    theta, u = _tma_friendly(_rows(theta)), _tma_friendly(_rows(u))
    assert theta.dtype == torch.float64 and u.dtype == torch.float64
    k, n = theta.shape
    m = u.shape[0]
    assert u.shape[1] == n
    if out is None:
        out = torch.empty((m, k), dtype=torch.float64, device=u.device)
    if m == 0 or k == 0:
        return out
    with torch.cuda.device(u.device):
        ws = _workspace(lib().rla_gemm_workspace_bytes(m, k, n), u.device)
        check(lib().rla_gauss_apply_explicit_f64(theta.data_ptr(), k, n, _ld(theta), u.data_ptr(), m, _ld(u),
                                                 out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(),
                                                 stream_ptr()), "rla_gauss_apply_explicit_f64")
    return out


def _tma_friendly_f32(x):
    """float32 rows for TMA: 16-byte aligned base, row stride a multiple of 4 elements."""
    if x.data_ptr() % 16 == 0 and _ld(x) % 4 == 0:
        return x
    m, n = x.shape
    buf = torch.zeros((m, (n + 3) // 4 * 4), dtype=x.dtype, device=x.device)
    buf[:, :n] = x
    return buf[:, :n]


def embed_apply_rng(seed, kind, scale, k, u, row0=0, col0=0, out=None, accumulate=False):
    """Sketch u (m, n) with rows [row0, row0 + k) and columns [col0, col0 + n) of the
    virtual matrix scale * g(seed, row, col); the matrix is never materialised.

    float64 blocks run on the FP64 tensor pipe (csrc/gemm.cu).  float32 blocks with a Theta that
    is exact in TF32 (KIND_RADEMACHER, KIND_NORMAL_TF32) run on the generation-5 tensor cores
    (csrc/gemm32.cu: tcgen05 kind::tf32, two-part split of the block, FP64 accumulation of the
    64-term partial sums); the sketch comes back as float64 either way."""
    if u.dtype == torch.float32 and kind in (KIND_RADEMACHER, KIND_NORMAL_TF32) and col0 % 32 == 0:
        u = _tma_friendly_f32(_rows(u))
        m, n = u.shape
        if out is None:
            assert not accumulate
            out = torch.empty((m, k), dtype=torch.float64, device=u.device)
        if m == 0 or k == 0:
            return out
        with torch.cuda.device(u.device):
            ws = _workspace(lib().rla_gemm32_workspace_bytes(m, k, n), u.device)
            check(lib().rla_embed_apply_rng_f32(int(seed) & (2 ** 64 - 1), int(kind), float(scale), int(row0), int(k),
                                                int(col0), n, u.data_ptr(), m, _ld(u), out.data_ptr(), out.stride(0),
                                                1 if accumulate else 0, ws.data_ptr(), ws.numel(), stream_ptr()),
                  "rla_embed_apply_rng_f32")
        return out
    if u.dtype == torch.float32:
        u = u.to(torch.float64)
    u = _tma_friendly(_rows(u))
    assert u.dtype == torch.float64
    m, n = u.shape
    if out is None:
        assert not accumulate
        out = torch.empty((m, k), dtype=torch.float64, device=u.device)
    if m == 0 or k == 0:
        return out
    with torch.cuda.device(u.device):
        ws = _workspace(lib().rla_gemm_workspace_bytes(m, k, n), u.device)
        check(lib().rla_embed_apply_rng_f64(int(seed) & (2 ** 64 - 1), int(kind), float(scale), int(row0), int(k),
                                            int(col0), n, u.data_ptr(), m, _ld(u), out.data_ptr(), out.stride(0),
                                            1 if accumulate else 0, ws.data_ptr(), ws.numel(), stream_ptr()),
              "rla_embed_apply_rng_f64")
    return out


def theta_materialize(seed, kind, scale, rows, cols, row0=0, col0=0, device=None):
    """The (rows, cols) block of the virtual matrix, bit-identical to what
    `embed_apply_rng` multiplies by."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((rows, cols), dtype=torch.float64, device=device)
    if rows == 0 or cols == 0:
        return out
    with torch.cuda.device(device):
        check(lib().rla_theta_materialize_f64(int(seed) & (2 ** 64 - 1), int(kind), float(scale), int(row0), rows,
                                              int(col0), cols, out.data_ptr(), out.stride(0), stream_ptr()),
              "rla_theta_materialize_f64")
    return out
