"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the
host-side plan builder validates its input, the embedding API keeps the reference's
constructor / option / error behaviour, the product never touches the oracle, and the
product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import rla4mor_b200 as rb
from oracle import embeddings_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rla_b200.h")).read()
    declared = set(re.findall(r"\b(rla_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = rb.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in rla_b200.h but not exported"
    assert set(rb.EXPORTED_SYMBOLS) == declared
    assert lib.rla_version() >= 100


def test_plan_builder_is_host_only_and_validates():
    lib = rb.lib()
    n, k = 5000, 300
    signs, idx = rb.draw_signs_and_indices(n, k, 3)
    s8 = np.ascontiguousarray(signs, dtype=np.int8)
    i64 = np.ascontiguousarray(idx, dtype=np.int64)
    h = ctypes.c_void_p()
    assert lib.rla_srht_plan_create(ctypes.byref(h), s8.ctypes.data, n, i64.ctypes.data, k, 8) == 0
    assert lib.rla_srht_plan_device_bytes(h) >= 2 * 64 * 8 + 4 * k
    assert lib.rla_srht_plan_passes(h) == 1
    assert lib.rla_srht_workspace_bytes(h, 10) > 0
    lib.rla_srht_plan_destroy(h)
    bad = i64.copy(); bad[5] = 2 ** 13                               # outside [0, 2^ceil(log2 n))
    assert lib.rla_srht_plan_create(ctypes.byref(h), s8.ctypes.data, n, bad.ctypes.data, k, 8) == -1
    assert b"outside" in lib.rla_last_error()
    assert lib.rla_srht_plan_create(ctypes.byref(h), s8.ctypes.data, n, i64.ctypes.data, k, 3) == -1
    # more than 4096 distinct indices -> two passes over x
    signs2, idx2 = rb.draw_signs_and_indices(2 ** 16, 6000, 1)
    s8b = np.ascontiguousarray(signs2, dtype=np.int8); i64b = np.ascontiguousarray(idx2, dtype=np.int64)
    assert lib.rla_srht_plan_create(ctypes.byref(h), s8b.ctypes.data, 2 ** 16, i64b.ctypes.data, 6000, 8) == 0
    assert lib.rla_srht_plan_passes(h) == 2
    lib.rla_srht_plan_destroy(h)


def test_workspace_queries_and_argument_checks_without_gpu():
    lib = rb.lib()
    assert lib.rla_gemm_workspace_bytes(512, 2000, 2 ** 22) >= 512 * 2000 * 8
    assert lib.rla_gemm_workspace_bytes(0, 5, 5) == 0
    assert lib.rla_srht_adjoint_workspace_bytes(3, 100) == 3 * 128 * 8
    assert lib.rla_fwht_f64(None, 2, 12, 12, None, 12, 1.0, None) == -1      # not a power of two (srht.py:110-111)
    assert b"power of two" in lib.rla_last_error()
    assert lib.rla_embed_apply_rng_f64(0, 7, 1.0, 0, 4, 0, 64, None, 1, 64, None, 4, 0, None, 0, None) == -1


def test_new_entry_points_validate_arguments_without_gpu():
    """Argument checks of the factorisation / triangular-solve / peer-exchange entry points
    happen on the host before any CUDA call."""
    import ctypes
    lib = rb.lib()
    assert lib.rla_sptrsv_solve_f64(None, None, None, None, None, None, None, None, None, None, None, None, 3, None, 4, 3, None) == -1   # odd ldx
    assert lib.rla_sptrsv_group_inverses_host(2, None, None, None, None, None, 0, None, None, None, None, None) == -1
    assert lib.rla_svd_jacobi_block_scratch_ints(256, 8, 30) == 30 + 32 + 8
    assert lib.rla_svd_jacobi_block_rows(1024, 256, 1) in (0, 2, 4, 8)          # 0 without a device
    assert lib.rla_svd_jacobi_cluster_size(256, 256, 0) in (0, 16)              # 0 without a device
    assert lib.rla_svd_jacobi_cluster_size(255, 256, 0) == 0                    # odd row length: not supported
    assert lib.rla_svd_jacobi_cluster_size(1024, 256, 1) == 0                   # 256 x 1280 does not fit 16 CTAs
    # odd ldx / negative sizes / null pointers -> RLA_ERR_INVALID (-1), no device touched
    assert lib.rla_sptrsv_transpose_in_f64(None, 3, 10, 10, None, None, 3, None) == -1
    assert lib.rla_sptrsv_permute_rows_f64(None, None, None, 5, 3, None) == -1
    assert lib.rla_trinv_upper_f64(None, 4, 2, None, 4, None) == -1
    assert lib.rla_peer_allreduce_f64(None, None, 2, 0, 1, 4, 4, 4, None, None, 4, None, 1.0, None) == -1
    one = (ctypes.c_void_p * 1)(8)
    out = ctypes.c_void_p(16)
    assert lib.rla_peer_allreduce_f64(one, one, 17, 0, 1, 4, 4, 4, None, out, 4, out, 1.0, None) == -1   # world > 16
    assert b"world size" in lib.rla_last_error()
    assert lib.rla_peer_allreduce_f64(one, one, 1, 0, 0, 4, 4, 4, None, out, 4, out, 1.0, None) == -1    # epoch 0
    # the plan builder rejects entries on the wrong side of the diagonal
    rowptr = np.array([0, 2, 3], dtype=np.int64); col = np.array([0, 1, 1], dtype=np.int32); val = np.ones(3)
    bufs = [np.empty(4, dtype=np.int64) for _ in range(12)]
    ns, ng, nl = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
    rc = lib.rla_sptrsv_plan_host(2, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, 1, 4096, 32, 256,
                                  bufs[0].ctypes.data, bufs[1].ctypes.data, bufs[2].ctypes.data, bufs[3].ctypes.data,
                                  bufs[4].ctypes.data, bufs[5].ctypes.data, bufs[6].ctypes.data, bufs[7].ctypes.data,
                                  bufs[8].ctypes.data, bufs[9].ctypes.data, ctypes.byref(ng), bufs[10].ctypes.data,
                                  bufs[11].ctypes.data, bufs[0].ctypes.data, bufs[1].ctypes.data, ctypes.byref(ns),
                                  ctypes.byref(nl))
    assert rc == -1 and b"wrong side" in lib.rla_last_error()


def test_draws_match_reference_expressions():
    n, k, seed = 1000, 50, 4
    r, s = rb.draw_signs_and_indices(n, k, seed)
    assert np.array_equal(r, np.random.RandomState(seed).choice([-1, 1], (n), True))
    assert np.array_equal(s, np.random.RandomState(seed).choice(range(2 ** 10), k, True))
    assert np.array_equal(r[:k], 2 * (s & 1) - 1)                    # correlated streams


def test_embedding_constructors_options_and_dims():
    src = rb.DeviceVectorSpace(1000, id="STATE")
    opt = {"epsilon": 0.5, "delta": 0.01, "oblivious_dim": 10}
    e = rb.SrhtEmbedding(source=src, options=opt, range_id="R", _seed=1)
    assert e.range.dim == eo.srht_compute_dim(opt, 1000) and e.range.id == "R" and e._seed == 1
    assert isinstance(e.options, rb.FrozenDict)
    with pytest.raises(TypeError):
        e.options["range_dim"] = 3
    cplx = dict(opt, dtype=complex)
    assert rb.SrhtEmbedding(source=src, options=cplx).range.dim == eo.srht_compute_dim(cplx, 1000)
    g = rb.GaussianEmbedding(source=src, options=opt, _seed=2)
    assert g.range.dim == eo.gaussian_compute_dim(opt)
    assert np.array_equal(g._random_matrix, eo.gaussian_random_matrix(g.range.dim, 1000, 2))   # eager, embeddings.py:230
    assert rb.GaussianEmbedding(source=src, options={"range_dim": 5})._seed is not None
    b = rb.BlockGaussianEmbedding(source=src, options={"range_dim": 70, "max_block_size": 32}, _seed=5)
    assert b.block_sizes == [32, 32, 6] and b.n_blocks == 3
    seeds, seed = eo.block_seeds(5, 3)
    assert np.array_equal(b.block_seeds, seeds) and b._seed == seed
    assert np.array_equal(b._get_random_block(2), eo.block_gaussian_block(70, 1000, 6, seeds[2]))
    i = rb.IdentityEmbedding(source=src)
    assert i.range.dim == 1000
    for bad in ({}, {"epsilon": 0.1}, {"epsilon": 0.1, "delta": 0.1}):
        with pytest.raises(AssertionError):
            rb.GaussianEmbedding(source=src, options=bad)
    with pytest.raises(AssertionError):
        rb.SrhtEmbedding()
    e2 = e.with_(_seed=9)
    assert type(e2) is type(e) and e2._seed == 9 and e2.options == e.options and e2 is not e
    comp = e @ rb.IdentityOperator(src)
    assert comp.source is src and comp.range == e.range


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rla4mor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "libfwht_oracle" not in text and "oracle._" not in text, f"{f} reaches into the oracle"


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_gpu():
    with pytest.raises(rb.RlaError):
        rb.srht(np.zeros((2, 8)), 3, seed=0)
    e = rb.GaussianEmbedding(source=rb.DeviceVectorSpace(8), options={"range_dim": 2}, _seed=0)
    with pytest.raises(rb.RlaError):
        e.apply(np.zeros((1, 8)))
    with pytest.raises(rb.RlaError):
        rb.fht_oop(np.zeros((2, 8)))
