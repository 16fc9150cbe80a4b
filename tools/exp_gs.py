"""Developer experiment: Gram-Schmidt (pyMOR semantics) of (r, k) blocks, many-CTA kernel."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import reductor_ops as ops


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rs = np.random.RandomState(0)
for r, k in ((256, 1024), (128, 1024), (64, 1000), (128, 2048), (200, 512), (100, 4000)):
    A = torch.from_numpy(rs.standard_normal((r, k))).cuda()
    ms = timed(lambda: ops.gram_schmidt(A))
    Q, R = ops.gram_schmidt(A)
    orth = float((Q @ Q.T - torch.eye(Q.shape[0], dtype=torch.float64, device="cuda")).norm())
    rec = float((R.T @ Q - A).norm() / A.norm())
    print(f"{r} x {k}: {ms:.3f} ms (wrapper incl. clone / flag check)  |QQ^T - I| = {orth:.1e}  |R^T Q - A| / |A| = {rec:.1e}", flush=True)
