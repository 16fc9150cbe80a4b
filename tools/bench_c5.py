"""BASELINE.json configs[4]: sketched randomized range finder of a 2^23 x 256 float64 block,
k = 1024, ROW-sharded over the ranks (each rank holds an n/G slab of every vector), followed by
the small factorisation of the (m, k) sketch.

    python tools/bench_c5.py                                   # one GPU
    torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/bench_c5.py

`run_c5` is also what bench.py reports as its `rangefinder_c5` secondary result.  Phases are
timed on the device (CUDA events, max over ranks): the sketch with its exchange over NVLink
peer memory (csrc/peer.cu), the same with an NCCL all-reduce for comparison, and the thin
QR / SVD of the sketch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M, LOGN, K = 256, 23, 1024


def run_c5(rank, world, dev, steps=10, warmup=3, hbm_peak=None, dmma_peak=None):
    import torch
    import torch.distributed as dist
    import rla4mor_b200 as rb
    from rla4mor_b200 import sharding, rangefinder, reductor_ops as ops

    n = 2 ** LOGN
    lib = rb.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rmax(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def timed(step, steps=steps, warmup=warmup):
        for _ in range(warmup):
            step()
        barrier()
        l0 = lib.rla_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (lib.rla_launch_count() - l0) // steps
        barrier()
        return rmax(ms), launches

    slab, ranges = sharding.srht_slabs(n, world) if world & (world - 1) == 0 else (None, sharding.gaussian_slabs(n, world))
    lo, hi = ranges[rank]
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    U = torch.empty((M, hi - lo), dtype=torch.float64, device=dev)
    for r0 in range(0, M, 32):
        U[r0:r0 + 32].normal_(generator=gen)
    reducer = None
    if world > 1:
        from rla4mor_b200.peer import PeerSketchReducer
        reducer = PeerSketchReducer(M, K)
    res = {"workload": "rangefinder_c5", "config": {
        "workload": "rangefinder_c5", "baseline_config": "configs[4]", "m": M, "n": n, "k": K,
        "partition": f"rows x{world} (n split into {world} slabs" + (", one peer-memory exchange of the (m, k) partials)"
                                                                     if world > 1 else ")"),
        "scaling": "strong", "slab_bytes_per_gpu": int(M * (hi - lo) * 8)}}
    out = {}
    for kind in ("srht", "gauss"):
        if kind == "srht" and slab is None:
            continue
        sk = lambda red=reducer: rangefinder.sketch_block(U, n, K, 0, kind, rank, world, reducer=red, check=False)
        S = sk()
        t_peer, l_peer = timed(sk)
        entry = {"sketch_ms": t_peer, "launches_per_step": int(l_peer)}
        if world > 1:
            t_nccl, _ = timed(lambda: sk(None))
            S2 = sk(None)
            err = float(((S - S2).norm() / S2.norm()).item())
            # bit-identical on all ranks (rank-ordered sum)
            ref = S.clone()
            dist.broadcast(ref, src=0)
            same = torch.tensor([1.0 if torch.equal(ref, S) else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            entry.update(sketch_nccl_ms=t_nccl, exchange="NVLink peer-memory kernel (rla_peer_allreduce_f64)",
                         peer_vs_nccl_rel_diff=err, identical_on_all_ranks=bool(same.item() == 1.0))
            reducer.check_status()
        if kind == "srht":
            byts = M * n * 8 + M * K * 8 + n + 4 * K
            entry["input_gbs"] = M * n * 8 / t_peer / 1e6
            entry["roofline"] = {"bound": "hbm", "unit": "GB/s", "achieved": byts / t_peer / 1e6 / world,
                                 "peak": hbm_peak, "frac": (byts / t_peer / 1e6 / world / hbm_peak) if hbm_peak else None,
                                 "note": "per GPU: (algorithmic bytes / G) / (sketch + exchange time)"}
        else:
            fl = 2.0 * K * n * M
            entry["input_gbs"] = M * n * 8 / t_peer / 1e6
            entry["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "achieved": fl / t_peer / 1e9 / world,
                                 "peak": dmma_peak, "frac": (fl / t_peer / 1e9 / world / dmma_peak) if dmma_peak else None,
                                 "note": "per GPU: (2kmn / G) / (sketch + exchange time)"}
        out[kind] = entry
    # the small factorisations (replicated on every rank; identical input on every rank)
    t_qr, _ = timed(lambda: rangefinder.thin_qr(S), max(2, steps // 2), 1)
    t_svd, _ = timed(lambda: rangefinder.sketch_svd(S), max(2, steps // 2), 1)
    t_gs, _ = timed(lambda: ops.gram_schmidt(S), 2, 1)
    t_svd_direct, _ = timed(lambda: ops.svd_jacobi(S, want_v=True), 2, 1)
    info_direct = [int(v) for v in ops.svd_jacobi.last_info.tolist()] if hasattr(ops.svd_jacobi, "last_info") else None
    rangefinder.sketch_svd(S)
    info_pre = [int(v) for v in ops.svd_jacobi.last_info.tolist()] if hasattr(ops.svd_jacobi, "last_info") else None
    phases_pre = getattr(ops.svd_jacobi, "last_phases", None)
    t_svd_old, _ = timed(lambda: ops.svd_jacobi(S, want_v=True, block=False, cluster=False), 2, 1)
    t_svd_grid, _ = timed(lambda: ops.svd_jacobi(S, want_v=True, cluster=False), 2, 1)
    # the Jacobi iteration alone on the m x m triangular factor (what sketch_svd runs after the QR)
    Rf = rangefinder.thin_qr(S)[1]
    Rp = torch.zeros((Rf.shape[0], M + (M & 1)), dtype=torch.float64, device=dev)
    Rp[:, :M] = Rf
    t_jac_cluster, _ = timed(lambda: ops.svd_jacobi(Rp, want_v=False), max(2, steps // 2), 1)
    t_jac_grid, _ = timed(lambda: ops.svd_jacobi(Rp, want_v=False, cluster=False), 2, 1)
    out["factorisation"] = {"thin_qr_ms": t_qr, "svd_ms": t_svd, "gram_schmidt_pymor_semantics_ms": t_gs,
                            "jacobi_on_R_cluster_kernel_ms": t_jac_cluster,
                            "jacobi_on_R_grid_synchronised_kernel_ms": t_jac_grid,
                            "cluster_kernel": {"cluster_size": int(lib.rla_svd_jacobi_cluster_size(Rp.shape[1], Rp.shape[0], 0)),
                                               "phase_kilocycles(stage load, rotation steps, stage barrier, push, "
                                               "round barrier)": phases_pre},
                            "svd_direct_on_sketch_ms": t_svd_direct, "svd_direct_grid_synchronised_ms": t_svd_grid,
                            "svd_round_per_launch_ms": t_svd_old,
                            "jacobi_info_direct(sweeps,converged,timeout)": info_direct,
                            "jacobi_info_preconditioned": info_pre}
    # whole range finder step, SRHT when the world size allows it
    kind = "srht" if "srht" in out else "gauss"
    t_all, l_all = timed(lambda: rangefinder.sketched_range_finder(U, n, K, 0, kind, rank, world, reducer=reducer),
                         max(2, steps // 2), 1)
    # the reference's own variant of this step stops at the QR (gram_schmidt + T = pinv(R),
    # mor/sketched_reductor.py:94-95); reported beside the full sketch + QR + SVD step
    t_qr_only, _ = timed(lambda: rangefinder.sketched_range_finder(U, n, K, 0, kind, rank, world, reducer=reducer,
                                                                   svd=False), max(2, steps // 2), 1)
    res.update(value=M * n * 8 / t_all / 1e6, unit="GB/s", ms_per_step=t_all, cols_per_s=M / t_all * 1e3,
               gpu_launches=int(l_all), embedding=kind, phases=out,
               qr_only={"ms_per_step": t_qr_only, "value": M * n * 8 / t_qr_only / 1e6, "unit": "GB/s",
                        "what": "sketch + exchange + thin QR + T = R^-1 (no SVD)"})
    if world > 1 and slab is not None:
        # parity of the row-sharded path against the CPU oracle on a small block (the 2-GPU pytest is
        # skipped on 1-GPU boxes): the oracle is used as the CHECKER only, outside every timed region
        try:
            import oracle
            from rla4mor_b200.peer import PeerSketchReducer
            ms_, ns_, ks_ = 4, 2 ** 14 + 40, 96
            xs = np.random.RandomState(99).standard_normal((ms_, ns_))
            _, rg = sharding.srht_slabs(ns_, world)
            a, b = rg[rank]
            xloc = torch.from_numpy(np.ascontiguousarray(xs[:, a:b])).to(dev)
            with PeerSketchReducer(ms_, ks_) as red_s:
                ys = sharding.srht_row_sharded(xloc, ns_, ks_, 5, rank, world, reducer=red_s).cpu().numpy()
            yn = sharding.srht_row_sharded(xloc, ns_, ks_, 5, rank, world).cpu().numpy()
            ref = oracle.srht(xs, ks_, seed=5)
            res["row_sharded_vs_oracle_rel"] = {
                "peer_exchange": float(np.linalg.norm(ys - ref) / np.linalg.norm(ref)),
                "nccl_allreduce": float(np.linalg.norm(yn - ref) / np.linalg.norm(ref)),
                "case": f"SRHT, {ms_} x {ns_} (ragged), k={ks_}, seed 5, {world} slabs; tolerance 1e-12"}
        except Exception as exc:
            res["row_sharded_vs_oracle_rel"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    if reducer is not None:
        reducer.check_status()
        reducer.close()
    del U
    torch.cuda.empty_cache()
    return res


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                        # keep NCCL's banner off stdout
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    res = run_c5(rank, world, dev, hbm_peak=peaks.get("hbm_gbs"), dmma_peak=37.0)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
