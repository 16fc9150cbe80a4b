"""BASELINE.json configs[3]: SketchedReductor on a thermal-block-like FEM problem, n = 1e6, Q = 4
affine terms, WITH the inverse product R^-1 in front of every sketch (Theta R^-1 A_q U):
CSR SpMM, sparse-LU solve (two sparse triangular solves on the device), sketch, and the whole
extend_basis; the CPU reference (SciPy SuperLU solve, what the reference's InverseLuOperator
calls) timed beside it on a subset of the right-hand sides.

    python tools/bench_c4.py [nx] [m]        # default 1000 64
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import rla4mor_b200 as rb
from rla4mor_b200.factorization import InverseLuOperator


def timeit(f, iters=3):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


def fem_terms(nx, Q=4):
    n = nx * nx
    ex = np.ones(nx)
    T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    L = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))).tocsr()
    idx = np.arange(n); blk = (idx // nx >= nx // 2) * 2 + (idx % nx >= nx // 2)
    terms = [(sp.diags((blk == q % 4).astype(float) + 0.01) @ L @ sp.diags((blk == q % 4).astype(float) + 0.01)
              + 1e-3 * sp.eye(n)).tocsr() for q in range(Q)]
    return terms, (L + sp.eye(n)).tocsc(), n


def run_c4(nx=1000, m=64, k=1000, hbm_peak=6549.8, with_cpu=True):
    """One dict for bench.py's `secondary` list (and for `python tools/bench_c4.py`)."""
    terms, R, n = fem_terms(nx)
    lib = rb.lib()
    res = {"workload": "sketched_reductor_c4", "unit": "ms",
           "config": {"workload": "sketched_reductor_c4", "baseline_config": "configs[3]", "n": n, "Q": len(terms), "m": m,
                      "k": k, "problem": f"P1-like 5-point FEM blocks on a {nx} x {nx} grid (SciPy-assembled; pyMOR is not in the image), "
                                         "inner product R = L + I factored by SuperLU on the host as in the reference",
                      "l2": "working set >> L2 (U block 0.5 GB, factors 0.9 GB)"}}
    space = rb.DeviceVectorSpace(n, id="S")
    ops_dev = [rb.MatrixOperator(A, source_id="S", range_id="S") for A in terms]
    t0 = time.time()
    rinv = InverseLuOperator(rb.MatrixOperator(R, source_id="S", range_id="S"), symetric=True)
    res["host_splu_s"] = time.time() - t0
    lu = rinv._device_lu
    t0 = time.time(); fL, fU = lu._factors(False)[:2]; res["host_plan_s"] = time.time() - t0
    res["L"] = {"nnz": fL.nnz, "row_levels": fL.nlevels, "launches": fL.nsteps, "group_levels": fL.group_levels, "groups": fL.ngroups}
    res["U"] = {"nnz": fU.nnz, "row_levels": fU.nlevels, "launches": fU.nsteps, "group_levels": fU.group_levels, "groups": fU.ngroups}
    U = torch.randn(m, n, dtype=torch.float64, device="cuda")
    Uva = space.from_numpy(U)
    res["spmm_ms"] = timeit(lambda: ops_dev[0].apply(Uva))
    nnz_a = terms[0].nnz
    res["spmm_algorithmic_GBs"] = (nnz_a * 12 + 2 * n * m * 8) / res["spmm_ms"] / 1e6
    V1 = ops_dev[0].apply(Uva)
    l0 = lib.rla_launch_count()
    res["lu_solve_ms"] = timeit(lambda: rinv.apply(V1))
    res["lu_solve_launches"] = int((lib.rla_launch_count() - l0) // 4)
    ldx = m + (m & 1)
    X = torch.randn(n, ldx, dtype=torch.float64, device="cuda")
    res["L_solve_ms"] = timeit(lambda: fL.solve_inplace(X.clone(), m))
    res["U_solve_ms"] = timeit(lambda: fU.solve_inplace(X.clone(), m))
    res["clone_ms"] = timeit(lambda: X.clone())
    # algorithmic traffic of one solve: every off-diagonal entry reads 8*m bytes of X and 12 bytes of the factor
    byts = (fL.nnz + fU.nnz) * (8 * m + 12) + 4 * n * m * 8
    # extend_basis applies R^-1 to the images of ALL Q affine terms in one solve of Q m right-hand sides
    Qn = len(terms)
    Vb = space.from_numpy(torch.cat([V1.data] * Qn, dim=0))
    res["lu_solve_batched_ms"] = timeit(lambda: rinv.apply(Vb))
    res["lu_solve_batched_rhs"] = Qn * m
    del Vb
    byts_b = (fL.nnz + fU.nnz) * (8 * Qn * m + 12) + 4 * n * Qn * m * 8
    res["roofline"] = {"bound": "hbm", "kernel": f"sptrsv kernels (the ONE LU solve of Q m = {Qn * m} right-hand sides that extend_basis issues)",
                       "unit": "GB/s", "achieved": byts_b / res["lu_solve_batched_ms"] / 1e6, "peak": hbm_peak,
                       "frac": byts_b / res["lu_solve_batched_ms"] / 1e6 / hbm_peak, "traffic": None,
                       "algorithmic": "(nnz(L) + nnz(U)) * (8 Q m + 12) + 4 n Q m 8 bytes per solve",
                       "single_solve_of_m_rhs": {"ms": res["lu_solve_ms"], "achieved": byts / res["lu_solve_ms"] / 1e6,
                                                 "frac": byts / res["lu_solve_ms"] / 1e6 / hbm_peak}}
    # parity against SuperLU on a few right-hand sides, and the CPU time of the reference's call
    sub = min(m, 8)
    Vh = V1.data[:sub].cpu().numpy()
    t0 = time.time(); ref = rinv.factorization.solve(Vh.T).T; cpu_s = time.time() - t0
    got = rinv.apply(space.from_numpy(V1.data[:sub])).to_numpy()
    res["parity_rel_fro_vs_superlu"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    if with_cpu:
        res["cpu_baseline"] = {"value": cpu_s * m / sub * 1e3, "unit": "ms per 64-rhs LU solve", "cores": 1, "kind": "reference",
                               "sample": f"SciPy SuperLU solve (what InverseLuOperator.apply calls, utilities/factorization.py:118-124) "
                                         f"of {sub} right-hand sides: {cpu_s:.2f} s; scaled to {m}", "host_cpus": os.cpu_count()}
    for kind, opt in (("srht", {"range_dim": k}), ("gauss", {"range_dim": k, "rng": "philox"})):
        emb = (rb.SrhtEmbedding if kind == "srht" else rb.GaussianEmbedding)(source=space, options=opt, _seed=0)
        res[f"sketch_{kind}_ms"] = timeit(lambda: emb.apply(U))
        for name, inv in (("identity_product", None), ("lu_inverse_product", rinv)):
            def ext():
                red = rb.SketchedReductor(ops_dev, [torch.ones(n, dtype=torch.float64, device="cuda")], emb,
                                          inverse_product=inv)
                red.extend_basis(U)
                return red
            l0 = lib.rla_launch_count()
            res[f"extend_basis_{kind}_{name}_ms"] = timeit(ext, 2)
            res[f"extend_basis_{kind}_{name}_launches"] = int((lib.rla_launch_count() - l0) // 3)
    # the headline of this config: one extend_basis(U) of m vectors with the SRHT and the LU inverse product
    res["value"] = res["extend_basis_srht_lu_inverse_product_ms"]
    res["ms_per_step"] = res["value"]
    res["gpu_launches"] = res["extend_basis_srht_lu_inverse_product_launches"]
    # the full-space basis update rb.lincomb(T.T) (mor/sketched_reductor.py:99-100) at this size
    C = torch.randn(m, m, dtype=torch.float64, device="cuda")
    from rla4mor_b200 import reductor_ops as rops
    res["lincomb_ms"] = timeit(lambda: rops.lincomb(C, U))
    res["lincomb_GBs"] = 2 * m * n * 8 / res["lincomb_ms"] / 1e6
    del U, X, V1
    torch.cuda.empty_cache()
    return res


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    print(json.dumps(run_c4(nx, m)))


if __name__ == "__main__":
    main()
