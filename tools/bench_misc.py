"""Developer timings: full FWHT, range finder, reductor pieces."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200 import reductor_ops as ops
from rla4mor_b200.rangefinder import sketched_range_finder

def timeit(f, iters=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

for m, d in [(64, 24), (4096, 12), (1024, 16), (8, 27)]:
    a = torch.randn(m, 2 ** d, dtype=torch.float64, device="cuda")
    t = timeit(lambda: rb.fht_oop(a))
    print(f"fht_oop ({m}, 2^{d}) f64: {t:.3f} ms  {2 * a.numel() * 8 / t / 1e6:.0f} GB/s (read+write once)")
    del a
U = torch.randn(256, 2 ** 23, dtype=torch.float64, device="cuda")
from rla4mor_b200.rangefinder import sketch_block
for kind in ("srht", "gauss"):
    print(f"sketch only {kind}: {timeit(lambda: sketch_block(U, 2 ** 23, 1024, 0, kind)):.2f} ms")
    t = timeit(lambda: sketched_range_finder(U, 2 ** 23, 1024, 0, kind), 2)
    t0 = timeit(lambda: sketched_range_finder(U, 2 ** 23, 1024, 0, kind, svd=False), 2)
    print(f"range finder 2^23 x 256, k=1024, {kind}: sketch+GS+SVD {t:.1f} ms, sketch+GS {t0:.1f} ms")
S = torch.randn(256, 1024, dtype=torch.float64, device="cuda")
print(f"gram_schmidt 256 x 1024: {timeit(lambda: ops.gram_schmidt(S)):.2f} ms; svd_jacobi: {timeit(lambda: ops.svd_jacobi(S), 2):.1f} ms")
t0 = time.time(); np.linalg.svd(S.cpu().numpy(), compute_uv=False); print(f"numpy svd (host) {1e3 * (time.time() - t0):.1f} ms")
