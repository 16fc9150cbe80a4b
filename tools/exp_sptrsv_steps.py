"""Experiment: per-step durations of one U (and L) triangular solve, timed with CUDA events in
the real back-to-back sequence (ncu's per-launch times are cold-cache and serialised)."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200._lib import check, lib, stream_ptr
from rla4mor_b200.factorization import InverseLuOperator

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = 64
ex = np.ones(nx)
T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
A = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx)) + sp.eye(nx * nx)).tocsc()
inv = InverseLuOperator(rb.MatrixOperator(A, source_id="S", range_id="S"), symetric=True)
fL, fU = inv._device_lu._factors(False)[:2]
n = nx * nx
X0 = torch.randn(n, m, dtype=torch.float64, device="cuda")
for name, f in (("L", fL), ("U", fU)):
    def one(s, X):
        check(lib().rla_sptrsv_solve_f64(f.rowptr.data_ptr(), f.col.data_ptr(), f.val.data_ptr(),
                                         None if f.diag is None else f.diag.data_ptr(), f.split.data_ptr(),
                                         f.grp_start.data_ptr(), f.grp_rows.data_ptr(), f.dinv_ptr.data_ptr(), f.dinv.data_ptr(),
                                         f.step_lo[s:].ctypes.data, f.step_hi[s:].ctypes.data, f.step_kind[s:].ctypes.data, 1,
                                         X.data_ptr(), m, X.stride(0), stream_ptr()), "solve")
    best = None
    for rep in range(3):
        X = X0.clone()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(f.nsteps + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for s in range(f.nsteps):
            one(s, X)
            ev[s + 1].record()
        torch.cuda.synchronize()
        t = np.array([ev[s].elapsed_time(ev[s + 1]) * 1e3 for s in range(f.nsteps)])
        best = t if best is None else np.minimum(best, t)
    kinds = f.step_kind
    cnt = f.step_hi - f.step_lo
    print(name, "steps", f.nsteps, "total %.0f us" % best.sum())
    for k in (0, 1, 2):
        sel = kinds == k
        print("  kind", k, "n", int(sel.sum()), "total %.0f us" % best[sel].sum(), "avg %.1f" % (best[sel].mean() if sel.any() else 0))
    # entries per step
    rp = f.rowptr.cpu().numpy(); spl = f.split.cpu().numpy()
    out = []
    for s in range(f.nsteps):
        if kinds[s] == 2:
            out.append((s, 2, int(cnt[s]), 0, best[s]))
        else:
            lo, hi = int(f.step_lo[s]), int(f.step_hi[s])
            ne = int((spl[lo:hi] - rp[lo:hi]).sum())
            out.append((s, int(kinds[s]), hi - lo, ne, best[s]))
    for o in out[:40] + out[-40:]:
        print("   step %4d kind %d rows/groups %6d ext entries %9d  %.1f us  %s" % (o + (("%.2f TB/s" % (o[3] * 524 / o[4] / 1e6)) if o[3] else "",)))
