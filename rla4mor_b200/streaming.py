"""Sketching blocks that live in HOST memory: the (m, n) block is cut into groups of
vectors (rows), each group is copied host->device on a copy stream into one of two
device staging buffers while the previous group is being sketched, and the (rows, k)
results are written into one device tensor (optionally copied back to the host).

This is the column blocking of the reference's `project_block`
(utilities/utilities.py:114-121) turned into a copy/compute pipeline; vectors are
independent, so no result depends on the grouping.
"""
import numpy as np
import torch

from ._lib import check, lib, require_cuda


def pinned_like(shape, dtype=torch.float64):
    require_cuda()
    return torch.empty(shape, dtype=dtype, pin_memory=True)


def apply_streamed(apply_fn, U_host, k, rows_per_chunk=None, out=None, device=None, return_host=False):
    """apply_fn: CUDA (rows, n) tensor -> CUDA (rows, k) tensor (e.g. `embedding.apply`).
    U_host: CPU torch tensor (pinned for full PCIe speed) or NumPy array, (m, n).
    Returns the (m, k) sketch on the device (or a pinned host tensor if return_host)."""
    require_cuda()
    if isinstance(U_host, np.ndarray):
        U_host = torch.from_numpy(np.ascontiguousarray(U_host))
    assert not U_host.is_cuda and U_host.dim() == 2
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    m, n = U_host.shape
    if rows_per_chunk is None:
        rows_per_chunk = max(1, min(m, (1 << 30) // max(1, n * U_host.element_size())))   # ~1 GiB per chunk
    if out is None:
        out = torch.empty((m, k), dtype=U_host.dtype, device=device)
    if m == 0:
        return out
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    bufs = [torch.empty((rows_per_chunk, n), dtype=U_host.dtype, device=device) for _ in range(2)]
    # the staging buffers come from the caching allocator of the MAIN stream and may be blocks whose
    # last consumer is still running there: the first copies must not overtake that work
    copy.wait_stream(main)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    chunks = [(lo, min(lo + rows_per_chunk, m)) for lo in range(0, m, rows_per_chunk)]
    for i, (lo, hi) in enumerate(chunks):
        b = i & 1
        with torch.cuda.stream(copy):
            if i >= 2:
                copy.wait_event(consumed[b])            # the sketch of chunk i-2 has read this buffer
            bufs[b][:hi - lo].copy_(U_host[lo:hi], non_blocking=True)
            copied[b].record(copy)
        main.wait_event(copied[b])
        out[lo:hi].copy_(apply_fn(bufs[b][:hi - lo]))
        consumed[b].record(main)
    if return_host:
        host = torch.empty((m, k), dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        main.synchronize()
        return host
    return out


def apply_streamed_rng(seed, kind, scale, k, U_host, cols_per_slab=None, device=None, return_host=False):
    """On-the-fly (Philox) dense sketch of a HOST block, streamed by COLUMN slabs of the
    vector dimension: slab j is copied host->device with a pitched copy while slab j-1 is
    sketched against Theta[:, slab] (column offset into the virtual matrix) and accumulated
    into the (m, k) result.  Unlike row chunks this keeps the GEMM at full height (every
    Theta entry is still generated once per 128 vectors) -- a block of m vectors costs the
    same k*n generated entries however it is cut along n."""
    from . import dense
    require_cuda()
    if isinstance(U_host, np.ndarray):
        U_host = torch.from_numpy(np.ascontiguousarray(U_host))
    assert not U_host.is_cuda and U_host.dim() == 2 and U_host.dtype in (torch.float64, torch.float32)
    assert U_host.dtype == torch.float64 or kind in (dense.KIND_RADEMACHER, dense.KIND_NORMAL_TF32), \
        "float32 host blocks need a Theta that is exact in TF32 (csrc/gemm32.cu)"
    assert U_host.stride(1) == 1
    es = U_host.element_size()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    m, n = U_host.shape
    out = torch.zeros((m, k), dtype=torch.float64, device=device)
    if m == 0 or n == 0:
        return out
    if cols_per_slab is None:
        cols_per_slab = max(4096, ((1 << 30) // (es * m)) // 4096 * 4096)         # ~1 GiB per slab
    cols_per_slab = max(32, (min(cols_per_slab, n) + 31) // 32 * 32)
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    bufs = [torch.empty((m, cols_per_slab), dtype=U_host.dtype, device=device) for _ in range(2)]
    copy.wait_stream(main)                              # see apply_streamed
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    spitch = U_host.stride(0) * es
    src0 = U_host.data_ptr()
    slabs = [(c, min(c + cols_per_slab, n)) for c in range(0, n, cols_per_slab)]
    with torch.cuda.device(device):
        for i, (c0, c1) in enumerate(slabs):
            b = i & 1
            w = c1 - c0
            if i >= 2:
                copy.wait_event(consumed[b])
            check(lib().rla_copy2d_async(bufs[b].data_ptr(), cols_per_slab * es, src0 + c0 * es, spitch, w * es, m, 0,
                                         ctypes_stream(copy)), "rla_copy2d_async")
            copied[b].record(copy)
            main.wait_event(copied[b])
            dense.embed_apply_rng(seed, kind, scale, k, bufs[b][:, :w], col0=c0, out=out, accumulate=(i > 0))
            consumed[b].record(main)
    if return_host:
        host = torch.empty((m, k), dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        main.synchronize()
        return host
    return out


def ctypes_stream(stream):
    import ctypes
    return ctypes.c_void_p(stream.cuda_stream)
