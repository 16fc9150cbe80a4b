"""Developer experiment: the three Jacobi SVD kernels on the shapes of the range finder
(triangular factor 256 x 256 with / without accumulated rotations, the 256 x 1024 sketch itself).

    python tools/exp_jacobi_cluster.py
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import reductor_ops as ops


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rs = np.random.RandomState(0)
S = rs.standard_normal((256, 1024))
R = np.linalg.qr(S.T)[1]                                    # 256 x 256 upper triangular
for name, M in (("R 256x256", R), ("S 256x1024", S), ("S 64x1000", rs.standard_normal((64, 1000)))):
    Md = torch.from_numpy(np.ascontiguousarray(M)).cuda()
    s_ref = np.linalg.svd(M, compute_uv=False)
    for want_v in (False, True):
        for forced in ("0", "1", "2", "4", "8", "16"):
            os.environ["RLA_JACOBI_CLUSTER"] = forced
            k, m = M.shape[1], M.shape[0]
            from rla4mor_b200 import lib
            C = lib().rla_svd_jacobi_cluster_size(k, m, 1 if want_v else 0)
            if forced not in ("0", "1") and C != int(forced):
                continue
            ms = timed(lambda: ops.svd_jacobi(Md, want_v=want_v))
            _, s, _ = ops.svd_jacobi(Md, want_v=want_v)
            err = float(np.max(np.abs(s.cpu().numpy() - s_ref)) / s_ref[0])
            print(f"{name} want_v={int(want_v)} RLA_JACOBI_CLUSTER={forced} C={C}: {ms:.3f} ms (wrapper incl. clone/sort), "
                  f"info={ops.svd_jacobi.last_info.tolist()} phases(kcyc: load, steps, waitA, push, barB)={getattr(ops.svd_jacobi, 'last_phases', None)} err={err:.1e}", flush=True)
