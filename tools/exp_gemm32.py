"""Developer check + timing of the float32 tcgen05 path (csrc/gemm32.cu)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import dense

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=512); ap.add_argument("--logn", type=int, default=20)
ap.add_argument("--k", type=int, default=2000); ap.add_argument("--kind", type=int, default=2)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
if a.check:
    for (m, n, k) in ((5, 100, 7), (130, 4096 + 36, 129), (256, 2 ** 15, 300), (64, 2 ** 18, 128)):
        torch.manual_seed(m)
        u = torch.randn(m, n, dtype=torch.float32, device="cuda")
        y = dense.embed_apply_rng(3, a.kind, 1.0 / np.sqrt(k), k, u)
        th = dense.theta_materialize(3, a.kind, 1.0 / np.sqrt(k), k, n)
        ref = u.to(torch.float64) @ th.T
        err = float((y - ref).norm() / ref.norm())
        y64 = dense.embed_apply_rng(3, a.kind, 1.0 / np.sqrt(k), k, u.to(torch.float64))
        err64 = float((y64 - ref).norm() / ref.norm())
        print(f"check m={m} n={n} k={k} kind={a.kind}: rel err f32 path {err:.2e}   (fp64 path {err64:.2e})", flush=True)
n = 2 ** a.logn
u = torch.randn(a.m, n, dtype=torch.float32, device="cuda")
f = lambda: dense.embed_apply_rng(0, a.kind, 1.0, a.k, u)
f(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
best = min(ts); fl = 2.0 * a.m * a.k * n
print(f"f32 tcgen05 kind={a.kind} m={a.m} n=2^{a.logn} k={a.k}: best {best:.2f} ms  {fl / best / 1e9:.1f} TFLOP/s", flush=True)
