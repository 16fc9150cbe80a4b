"""Peer-memory exchange of the row-sharded sketch (csrc/peer.cu).

`PeerSketchReducer(m, k, group)` owns, per rank, one cudaMalloc'ed buffer that every other
rank of the node maps through CUDA IPC (handles travel once through `all_gather_object`):

    [ flag row: 16 x uint64 | status int32 | pad ] [ partial A (m, k) f64 ] [ partial B (m, k) f64 ]

A row-sharded sketch then costs the local sketch kernel, which writes its partial STRAIGHT
into the current partial buffer (`partial()`), plus ONE kernel (`reduce()`): publish, wait
for the peers, rank-ordered signed sum over NVLink.  Results are bit-identical on all ranks.
torch.distributed is used for the one-time handle exchange and a host barrier only.
"""
import ctypes

import numpy as np

from ._lib import c_void_p, check, lib, require_cuda, stream_ptr, RlaError

_HEADER = 256      # bytes: 16 flags (128 B) + status + pad, keeps the partials 256-byte aligned
MAX_WORLD = 16


class _CudaView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap memory we cudaMalloc'ed."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerSketchReducer:
    def __init__(self, m, k, group=None, timeout_s=10.0):
        torch = require_cuda()
        import torch.distributed as dist
        assert dist.is_initialized(), "PeerSketchReducer needs an initialised process group"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert self.world <= MAX_WORLD, f"at most {MAX_WORLD} ranks (one node)"
        self.m, self.k = int(m), int(k)
        self.timeout_s = float(timeout_s)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._pbytes = -(-self.m * self.k * 8 // 256) * 256
        nbytes = _HEADER + 2 * max(self._pbytes, 256)
        own = c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        check(lib().rla_peer_buffer_create(nbytes, ctypes.byref(own), handle), "rla_peer_buffer_create")
        self._own = own.value
        self._opened = []
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._base = []
        for p in range(self.world):
            if p == self.rank:
                self._base.append(self._own)
                continue
            ptr = c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[p])
            check(lib().rla_peer_buffer_open(buf, ctypes.byref(ptr)), "rla_peer_buffer_open")
            self._opened.append(ptr.value)
            self._base.append(ptr.value)
        self.epoch = 0
        self._flag_ptrs = (c_void_p * self.world)(*self._base)
        self._part_ptrs = [(c_void_p * self.world)(*[b + _HEADER + par * max(self._pbytes, 256) for b in self._base])
                           for par in (0, 1)]
        self._status_ptr = self._own + 128
        self._partials = [torch.as_tensor(_CudaView(self._own + _HEADER + par * max(self._pbytes, 256),
                                                    (self.m, self.k), "<f8", self), device=self.device)
                          for par in (0, 1)] if self.m * self.k else [None, None]
        self._status = torch.as_tensor(_CudaView(self._status_ptr, (1,), "<i4", self), device=self.device)
        torch.cuda.synchronize()
        dist.barrier(group=group)              # every buffer is mapped and zeroed before the first epoch

    def partial(self):
        """The (m, k) buffer the NEXT reduce() will publish: pass it as `out=` to the sketch kernel."""
        return self._partials[(self.epoch + 1) & 1]

    def reduce(self, high=None, out=None):
        """Sum the partials written into `partial()` on all ranks (one kernel, current stream).
        high: optional (k,) int32 CUDA tensor of slab bits for the SRHT sign factors."""
        import torch
        self.epoch += 1
        if out is None:
            out = torch.empty((self.m, self.k), dtype=torch.float64, device=self.device)
        if self.m * self.k == 0:
            return out
        assert out.stride(1) == 1 and out.dtype == torch.float64
        with torch.cuda.device(self.device):
            check(lib().rla_peer_allreduce_f64(self._part_ptrs[self.epoch & 1], self._flag_ptrs, self.world, self.rank,
                                               self.epoch, self.m, self.k, self.k,
                                               None if high is None else high.data_ptr(),
                                               out.data_ptr(), out.stride(0), self._status_ptr,
                                               self.timeout_s, stream_ptr()), "rla_peer_allreduce_f64")
        return out

    def check_status(self):
        """Synchronises; raises if a reduce() timed out waiting for a peer."""
        if int(self._status.item()) != 0:
            raise RlaError("rla_peer_allreduce_f64: a peer did not publish its partial sketch in time")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        """Collective: every rank must call it (peers unmap before the owner frees)."""
        import torch
        if self._own is None:
            return
        torch.cuda.synchronize()
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)  # nobody unmaps while a peer may still read
        except Exception:
            pass
        self._partials = [None, None]
        self._status = None
        for p in self._opened:
            lib().rla_peer_buffer_close(c_void_p(p))
        self._opened = []
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)
        except Exception:
            pass
        lib().rla_peer_buffer_destroy(c_void_p(self._own))
        self._own = None


def srht_high_bits(sampling, n, world):
    """int32 slab bits s_i >> log2(slab) of the row-sharded SRHT (sharding.srht_slabs)."""
    from .sharding import srht_slabs
    slab = srht_slabs(n, world)[0]
    return (np.asarray(sampling, dtype=np.int64) // slab).astype(np.int32)
