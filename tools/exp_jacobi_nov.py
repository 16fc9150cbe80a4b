"""Experiment: block Jacobi on the 256 x 256 triangular factor with / without accumulated rotations."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import reductor_ops as ops
rs = np.random.RandomState(0)
S = torch.from_numpy(rs.standard_normal((256, 1024))).cuda()
Q, R = ops.gram_schmidt(S)
M = R.contiguous()
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for wv in (True, False):
    ms = t(lambda: ops.svd_jacobi(M, want_v=wv))
    print("B env", os.environ.get("RLA_JACOBI_B"), "want_v", wv, "%.3f ms" % ms, ops.svd_jacobi.last_info.tolist(), flush=True)
print("pinv_R %.3f ms" % t(lambda: ops.pinv_R(R)), "gemm_nn 256^3 %.3f ms" % t(lambda: ops.gemm_nn(M, M)),
      "gemm_nn 256x256x1024 %.3f ms" % t(lambda: ops.gemm_nn(M, Q)), "gs %.3f" % t(lambda: ops.gram_schmidt(S)))
