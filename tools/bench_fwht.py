"""Developer timing of the full FWHT (fht_oop) and of the tall-skinny lincomb kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200 import reductor_ops as ops

def timeit(f, iters=5):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]

PEAK = 6549.8
for m, d in [(64, 24), (4096, 12), (4096, 13), (1024, 16), (8, 27), (256, 20), (3, 26)]:
    a = torch.randn(m, 2 ** d, dtype=torch.float64, device="cuda")
    out = torch.empty_like(a)
    from rla4mor_b200.srht import _fwht_device
    best, med = timeit(lambda: _fwht_device(a, 1.0 / 2 ** (d / 2), out=out))
    passes = 1 if d <= 13 else 1 + -(-(d - 13) // 11)
    rw = 2 * a.numel() * 8
    print(f"fwht ({m}, 2^{d}) f64: best {best:.3f} ms med {med:.3f} ms; {passes} pass(es): {passes * rw / best / 1e6:.0f} GB/s per pass "
          f"= {passes * rw / best / 1e6 / PEAK:.2f} of HBM; whole transform vs read+write-once {rw / best / 1e6 / PEAK:.2f}", flush=True)
    del a, out
for r, n in [(64, 10 ** 6), (128, 10 ** 6), (256, 10 ** 6), (256, 2 ** 23), (32, 2 ** 22), (16, 10 ** 7)]:
    X = torch.randn(r, n, dtype=torch.float64, device="cuda")
    C = torch.randn(r, r, dtype=torch.float64, device="cuda")
    best, med = timeit(lambda: ops.lincomb(C, X))
    byts, flops = 2 * r * n * 8, 2.0 * r * r * n
    print(f"lincomb ({r} x {r}) @ ({r} x {n}): best {best:.3f} ms; {byts / best / 1e6:.0f} GB/s ({byts / best / 1e6 / PEAK:.2f} of HBM), "
          f"{flops / best / 1e9:.1f} TFLOP/s; torch.matmul {timeit(lambda: C @ X)[0]:.3f} ms", flush=True)
    del X
