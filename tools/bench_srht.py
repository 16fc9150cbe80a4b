"""Developer timing of the SRHT kernel (not the contract bench; see bench.py)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=128)
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--k", type=int, default=4000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="f64")
    a = ap.parse_args()
    n = a.n or 2 ** a.logn
    dt = torch.float64 if a.dtype == "f64" else torch.float32
    x = torch.randn(a.m, n, dtype=dt, device="cuda")
    t0 = time.time()
    plan = rb.get_plan(n, a.k, 0, dt, x.device)
    t_plan = time.time() - t0
    y = plan.apply(x)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.apply(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    es = x.element_size()
    byts = a.m * n * es + a.m * a.k * es + n + 4 * a.k
    best, med = min(ts), sorted(ts)[len(ts) // 2]
    print(f"m={a.m} n={n} k={a.k} {a.dtype} plan={t_plan:.2f}s  best={best:.3f} ms med={med:.3f} ms  "
          f"{byts / best / 1e6:.1f} GB/s (best) {byts / med / 1e6:.1f} GB/s (med)  frac_of_6549.8={byts / med / 1e6 / 6549.8:.3f}  "
          f"log2L={os.environ.get('RLA_SRHT_LOG2L', 'auto')}", flush=True)


if __name__ == "__main__":
    main()
