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

# ---- configs[3]: SketchedReductor on a thermal-block-like FEM problem, n = 1e6, Q = 4
import scipy.sparse as sp
def fem_terms(nx, Q=4):
    n = nx * nx
    ex = np.ones(nx)
    T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    L = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx))).tocsr()
    idx = np.arange(n); blk = (idx // nx >= nx // 2) * 2 + (idx % nx >= nx // 2)
    return [(sp.diags((blk == q % 4).astype(float) + 0.01) @ L @ sp.diags((blk == q % 4).astype(float) + 0.01) + 1e-3 * sp.eye(n)).tocsr() for q in range(Q)], n
terms, n = fem_terms(1000)
m, k = 64, 1000
space = rb.DeviceVectorSpace(n, id="S")
ops_dev = [rb.MatrixOperator(A, source_id="S", range_id="S") for A in terms]
U = torch.randn(m, n, dtype=torch.float64, device="cuda")
Uva = space.from_numpy(U)
nnz = terms[0].nnz
t_spmm = timeit(lambda: ops_dev[0].apply(Uva))
byts = nnz * 12 + 2 * m * n * 8
print(f"C4 SpMM A_q U (n=1e6, nnz={nnz}, m={m}): {t_spmm:.3f} ms  {byts / t_spmm / 1e6:.0f} GB/s algorithmic")
for kind, opt in (("srht", {"range_dim": k}), ("gauss", {"range_dim": k, "rng": "philox"})):
    emb = (rb.SrhtEmbedding if kind == "srht" else rb.GaussianEmbedding)(source=space, options=opt, _seed=0)
    t_sk = timeit(lambda: emb.apply(U))
    def ext():
        red = rb.SketchedReductor(ops_dev, [torch.ones(n, dtype=torch.float64, device="cuda")], emb)
        red.extend_basis(U)
        return red
    t_ext = timeit(ext, 2)
    print(f"C4 {kind}: sketch of {m} vectors {t_sk:.2f} ms; extend_basis (Theta U, 4 x Theta A_q U, rhs, Gram-Schmidt, basis update) {t_ext:.1f} ms")
