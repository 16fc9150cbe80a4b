"""Developer timing of the dense sketch GEMM (not the contract bench)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import dense  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=512); ap.add_argument("--logn", type=int, default=20)
ap.add_argument("--k", type=int, default=2000); ap.add_argument("--mode", default="rng")
ap.add_argument("--kind", type=int, default=0); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
n = 2 ** a.logn
u = torch.randn(a.m, n, dtype=torch.float64, device="cuda")
if a.mode == "explicit":
    th = torch.randn(a.k, n, dtype=torch.float64, device="cuda")
    f = lambda: dense.gauss_apply_explicit(th, u)
else:
    f = lambda: dense.embed_apply_rng(0, a.kind, 1.0, a.k, u)
f(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
best = min(ts); fl = 2.0 * a.m * a.k * n
print(f"{a.mode} kind={a.kind} m={a.m} n=2^{a.logn} k={a.k}: best {best:.2f} ms  {fl / best / 1e9:.2f} TFLOP/s  "
      f"({fl / best / 1e9 / 37.16:.3f} of 37.16 DMMA peak)  waves={os.environ.get('RLA_GEMM_WAVES', '8')}", flush=True)
