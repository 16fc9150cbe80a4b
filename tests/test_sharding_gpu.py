"""Multi-GPU paths on real devices (needs >= 2 GPUs; skipped otherwise): column-sharded
sketches need no collective; row-sharded sketches agree with the single-GPU sketch after one
NCCL all-reduce.  Also the host-block streaming front end on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from golden_util import rel_fro

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, k, seed, out):
    import rla4mor_b200 as rb
    from rla4mor_b200 import sharding, dense
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x = np.random.RandomState(0).standard_normal((6, n))
        # column-sharded: each rank sketches its own vectors, no collective
        lo, hi = sharding.column_shard(6, rank, world)
        yc = torch.zeros(6, k, dtype=torch.float64, device="cuda")
        if hi > lo:
            yc[lo:hi] = rb.srht(torch.from_numpy(x[lo:hi]).cuda(), k, seed=seed)
        sharding.all_reduce_sum(yc)                      # only to collect the pieces for the check
        # row-sharded SRHT
        slab, ranges = sharding.srht_slabs(n, world)
        a, b = ranges[rank]
        ys = sharding.srht_row_sharded(torch.from_numpy(np.ascontiguousarray(x[:, a:b])).cuda(), n, k, seed, rank, world)
        # row-sharded on-the-fly Gaussian
        ga, gb = sharding.gaussian_slabs(n, world)[rank]
        yg = sharding.gaussian_row_sharded(torch.from_numpy(np.ascontiguousarray(x[:, ga:gb])).cuda(), n, k, seed, rank, world)
        # the same two through the NVLink peer-memory exchange (csrc/peer.cu), three epochs
        from rla4mor_b200.peer import PeerSketchReducer
        red = PeerSketchReducer(6, k)
        xs = torch.from_numpy(np.ascontiguousarray(x[:, a:b])).cuda()
        xg = torch.from_numpy(np.ascontiguousarray(x[:, ga:gb])).cuda()
        ps = sharding.srht_row_sharded(xs, n, k, seed, rank, world, reducer=red).clone()
        pg = sharding.gaussian_row_sharded(xg, n, k, seed, rank, world, reducer=red).clone()
        ps2 = sharding.srht_row_sharded(xs, n, k, seed, rank, world, reducer=red).clone()
        red.check_status()
        same = torch.tensor([float(torch.equal(ps, ps2))], device="cuda")
        gathered = [torch.empty_like(ps) for _ in range(world)]
        dist.all_gather(gathered, ps)
        same *= float(all(torch.equal(g, ps) for g in gathered))     # bit-identical on every rank
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        red.close()
        # odd k (scalar path of the exchange kernel), m = 1
        red1 = PeerSketchReducer(1, k + 1)
        p1 = sharding.srht_row_sharded(xs[:1], n, k + 1, seed, rank, world, reducer=red1).clone()
        n1 = sharding.srht_row_sharded(xs[:1], n, k + 1, seed, rank, world)
        same *= float(torch.equal(p1, n1) or float(((p1 - n1).norm() / n1.norm()).item()) < 1e-14)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        red1.check_status()
        red1.close()
        if rank == 0:
            full = dense.embed_apply_rng(seed, 0, 1.0 / np.sqrt(k), k, torch.from_numpy(x).cuda())
            np.savez(out, yc=yc.cpu().numpy(), ys=ys.cpu().numpy(), yg=yg.cpu().numpy(), full=full.cpu().numpy(),
                     ps=ps.cpu().numpy(), pg=pg.cpu().numpy(), same=same.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [2 ** 15, 40000])
def test_two_gpu_sharding(tmp_path, n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    k, seed = 300, 5
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, k, seed, out), nprocs=2, join=True)
    z = np.load(out)
    x = np.random.RandomState(0).standard_normal((6, n))
    ref = oracle.srht(x, k, seed=seed)
    assert rel_fro(z["yc"], ref) < 1e-12
    assert rel_fro(z["ys"], ref) < 1e-12
    assert rel_fro(z["yg"], z["full"]) < 1e-12
    assert rel_fro(z["ps"], ref) < 1e-12                             # peer-memory exchange, SRHT slab signs in-kernel
    assert rel_fro(z["pg"], z["full"]) < 1e-12
    assert z["same"][0] == 1.0                                       # deterministic and identical on all ranks


def test_streaming_host_block():
    import rla4mor_b200 as rb
    from rla4mor_b200.streaming import apply_streamed
    n, k = 20000, 128
    x = np.random.RandomState(1).standard_normal((37, n))
    emb = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=2)
    host = torch.from_numpy(x).pin_memory()
    y = apply_streamed(emb.apply, host, k, rows_per_chunk=8, return_host=True)
    assert not y.is_cuda and rel_fro(y.numpy(), oracle.srht(x, k, seed=2)) < 1e-12
    y2 = apply_streamed(emb.apply, x, k, rows_per_chunk=5)
    assert y2.is_cuda and rel_fro(y2.cpu().numpy(), oracle.srht(x, k, seed=2)) < 1e-12


def test_embedding_apply_on_host_blocks():
    import rla4mor_b200 as rb
    from oracle import embeddings_oracle as eo
    n, k = 9000, 64
    x = np.random.RandomState(2).standard_normal((21, n))
    host = torch.from_numpy(x).pin_memory()
    g = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k, "rng": "philox"}, _seed=4)
    y = g.apply(host)                                   # column-slab streaming, result on the host
    assert not y.is_cuda and rel_fro(y.numpy(), eo.gaussian_apply(x, g.get_random_matrix())) < 1e-12
    from rla4mor_b200.streaming import apply_streamed_rng
    y2 = apply_streamed_rng(4, 0, 1.0 / np.sqrt(k), k, host, cols_per_slab=2048)      # several slabs
    assert rel_fro(y2.cpu().numpy(), y.numpy()) < 1e-13
    e = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=4)
    assert rel_fro(e.apply(host).numpy(), eo.gaussian_apply(x, eo.gaussian_random_matrix(k, n, 4))) < 1e-12
    s = rb.SrhtEmbedding(source=rb.DeviceVectorSpace(n), options={"range_dim": k}, _seed=4)
    assert rel_fro(s.apply(host).numpy(), oracle.srht(x, k, seed=4)) < 1e-12
