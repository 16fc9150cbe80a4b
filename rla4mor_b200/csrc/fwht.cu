// fwht.cu -- full (all outputs) Walsh-Hadamard transform along rows, natural order.
//
// Replaces fht_oop / fht_ip (rla/srht.py:99-134) and backs the implicit SRHT adjoint
// (SrhtEmbedding.apply_adjoint, rla/embeddings.py:175-178).
//
// A row of length 2^d is transformed in passes over 8192-element tiles (13 index bits per
// pass; round 1 used 4096-element tiles and needed 12 + 10 + 2 bits = three passes at d = 24):
//   pass 1  : bits [0, min(d,13))            tile = 8192 contiguous elements
//             (for d < 13 a tile packs 2^(13-d) rows);
//   pass >=2: 11 further bits each, the tile carrying 2 contiguous low bits (32-byte
//             sectors) plus 11 strided bits           =>  d = 24 in TWO passes.
// Every pass reads and writes each element once (HBM-bound); a tile reads all its elements
// before writing them, so passes after the first run in place on `out`.
//
// One tile = 128 threads x 64 registers, three register rounds with two shared-memory
// exchanges (element index e, 13 bits; l = bit 0):
//   round 1  registers (l, bits 8..12), thread = bits 1..7   <- coalesced 16-byte global loads;
//            stages 8..12 (l only rides along)
//   round 2  registers  bits 0..5,      thread = bits 6..12     stages 0..5
//   round 3  registers (l, bits 6, 7, 10..12), thread = bits 1..5, 8, 9 -> coalesced stores; stages 6, 7
// so the exchange that round 1 of this project spent on re-ordering for the stores now also
// carries butterfly stages.  All four shared-memory passes are 16-byte accesses (8-byte for
// float) and conflict free under the XOR swizzle  pos(e) = e ^ (((e >> 6) & SW) << 1).
#include "tile.cuh"
#include <algorithm>
#include <stdlib.h>

namespace rla {

struct FwhtPass {
    int d;        // log2 of the row length
    int cbits;    // contiguous low bits carried (not transformed) inside a tile
    int nb;       // bits transformed in this pass
    int lo;       // position of the first transformed bit in the row index
    int xb;       // 13 - cbits - nb: spare tile bits enumerating other (row, index) combos
    int64_t m;    // rows
    int64_t nq;   // m * 2^(d - cbits - nb): number of (row, other-bits) combinations
    int flags;    // bit 3: launched as clusters of neighbouring tiles (MODE 2)
};

// global element offset (row * ld + column) of tile element e of tile t; returns -1 when
// the element belongs to a row >= m (padding of the last tile)
__device__ __forceinline__ int64_t fwht_map(const FwhtPass &p, int64_t t, int e, int64_t ld) {
    const int c = e & ((1 << p.cbits) - 1);
    const int bf = (e >> p.cbits) & ((1 << p.nb) - 1);
    const int ex = e >> (p.cbits + p.nb);
    const int64_t q = (t << p.xb) + ex;
    if (q >= p.nq) return -1;
    const int O = p.d - p.cbits - p.nb;           // other bits per row
    const int64_t row = q >> O;
    const int64_t o = q & ((int64_t(1) << O) - 1);
    const int nlow = p.lo - p.cbits;              // other bits below the transformed range
    const int64_t col = c | ((o & ((int64_t(1) << nlow) - 1)) << p.cbits) | ((int64_t)bf << p.lo) |
                        ((o >> nlow) << (p.lo + p.nb));
    return row * ld + col;
}

// butterflies on the register bits selected by `mask` (bit i of mask <-> register bit i)
template <typename T>
__device__ __forceinline__ void butterflies64_masked(T (&v)[64], int mask) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if (mask & (1 << b)) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                if ((i & (1 << b)) == 0) {
                    T p = v[i], q = v[i | (1 << b)];
                    v[i] = p + q;
                    v[i | (1 << b)] = p - q;
                }
            }
        }
    }
}

// same with the mask known at compile time (no branches, no register shuffling at their joins)
template <int MASK, typename T>
__device__ __forceinline__ void butterflies64_static(T (&v)[64]) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if (MASK & (1 << b)) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                if ((i & (1 << b)) == 0) {
                    T p = v[i], q = v[i | (1 << b)];
                    v[i] = p + q;
                    v[i | (1 << b)] = p - q;
                }
            }
        }
    }
}

constexpr int FW_TL = 13;                 // tile bits
constexpr int FW_TILE = 1 << FW_TL;
constexpr int FW_THREADS = 128;

template <typename T> struct FwPair;
template <> struct FwPair<double> { using type = double2; static constexpr int SW = 7; };
template <> struct FwPair<float> { using type = float2; static constexpr int SW = 15; };

// One pass: persistent CTAs, each looping over tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// The loop is rotated so that there is ONE load site: the stores of tile t are interleaved,
// pair by pair, with the loads of the next tile into the registers they free, so the next
// tile's HBM latency overlaps this tile's stores and the other CTAs' butterflies.
// MODE 0  generic index map (fwht_map): ragged tiles, packed short rows, unaligned rows
// MODE 1  every tile is 8192 contiguous, 16-byte aligned elements of one row (first pass)
// MODE 2  strided pass whose tile is exactly  2^cbits contiguous x 2^nb strided  elements
//         (cbits + nb = 13): offsets are linear in the register index, no per-element map
// PAIR    (MODE 1 only) 256-thread CTAs of two 128-thread groups that transform two adjacent
//         tiles and then exchange them through shared memory for one more stage (bit 13):
//         14 bits in the first pass, so that d = 24 needs 14 + 10, i.e. two passes, the second
//         one with 64-byte contiguous pieces (32-byte pieces run at half the bandwidth).
// CB: contiguous bits of a MODE 2 pass (3..6), known at compile time like the stage masks of the
// two fast modes (13 - CB strided stages; all 13 in MODE 1)
template <typename T, int MODE, bool PAIR, int CB>
__global__ void __launch_bounds__(PAIR ? 2 * FW_THREADS : FW_THREADS, PAIR ? 1 : 3)
fwht_pass_kernel(const T *__restrict__ in, int64_t ldi, T *__restrict__ out, int64_t ldo,
                 FwhtPass p, int64_t ntiles, T post_scale, int vec) {
    static_assert(!PAIR || MODE == 1, "paired tiles are contiguous");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Pair = typename FwPair<T>::type;
    constexpr int SW = FwPair<T>::SW;
    constexpr bool FAST = MODE != 0;
    const int grp = PAIR ? (int)(threadIdx.x >> 7) : 0;
    T *buf = reinterpret_cast<T *>(smem_raw) + grp * FW_TILE;
    T *pbuf = reinterpret_cast<T *>(smem_raw) + (grp ^ 1) * FW_TILE;     // partner's buffer (PAIR)
    int tid = threadIdx.x & (FW_THREADS - 1);
    auto group_sync = [&]() {
        if (PAIR) asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        else __syncthreads();
    };
    // stage mask over tile bits -> register-bit masks of the three rounds
    const int smask = ((1 << (p.nb > FW_TL ? FW_TL : p.nb)) - 1) << p.cbits;
    const int m1 = ((smask >> 8) & 31) << 1;                     // round 1: tile bits 8..12 (bit 0 rides along for 16-byte accesses)
    const int m2 = smask & 63;                                   // round 2: tile bits 0..5
    const int m3 = ((smask >> 6) & 3) << 1;                      // round 3: tile bits 6, 7
    // the fast modes know their stages at compile time
    constexpr int SMASK = MODE == 1 ? (1 << FW_TL) - 1 : (MODE == 2 ? (((1 << (FW_TL - CB)) - 1) << CB) : 0);
    constexpr int M1 = ((SMASK >> 8) & 31) << 1, M2 = SMASK & 63, M3 = ((SMASK >> 6) & 3) << 1;
    const int tiles_per_row_log2 = p.d - FW_TL;                  // MODE 1
    // MODE 2: offset of tile element e = (e & cmask) + ((e >> cbits) << lo), tile base from t
    const int cb = MODE == 2 ? CB : p.cbits, lo = p.lo;
    auto tile_base = [&](int64_t t, int64_t ld) -> int64_t {
        if (MODE == 1) {
            const int64_t row = t >> tiles_per_row_log2;
            return row * ld + ((t - (row << tiles_per_row_log2)) << FW_TL);
        }
        const int O = p.d - p.cbits - p.nb, nlow = p.lo - p.cbits;
        const int64_t row = t >> O, o = t & ((int64_t(1) << O) - 1);
        return row * ld + ((o & ((int64_t(1) << nlow) - 1)) << cb) + ((o >> nlow) << (lo + p.nb));
    };
    auto elem_off = [&](int e) -> int64_t {
        if (MODE == 1) return e;
        return (int64_t)(e & ((1 << cb) - 1)) + ((int64_t)(e >> cb) << lo);
    };
    const bool scale = post_scale != T(1);
    T v[64];
    const int64_t G = gridDim.x;
    // MODE 2 may run as thread-block clusters of neighbouring tiles that enter their memory phase
    // together (p.flags bit 3): the 64-byte pieces of neighbouring tiles are adjacent in memory, so
    // a cluster touches runs of cluster_size x 64 bytes at a time.  All CTAs then take the same
    // number of trips (a CTA without a tile still meets the barrier).
    const bool clustered = MODE == 2 && (p.flags & 8);
    const int64_t t_end = clustered ? ((ntiles + G - 1) / G) * G : ntiles;
    // PAIR: the loop runs over tile PAIRS; this group's tile is 2 * t + grp
    for (int64_t t = (int64_t)blockIdx.x - G; t < t_end; t += G) {
        const int64_t tn = t + G;
        asm volatile("" : "+r"(tid));                            // keeps the swizzled addresses out of the loop-invariant set
        // element offsets of this thread in the round-1 (load) and round-3 (store) layouts
        const int e1 = 2 * tid;                                                        // + 256 h
        const int e3 = 2 * (tid & 31) + 256 * ((tid >> 5) & 1) + 512 * ((tid >> 6) & 1);   // + 64 b6 + 128 b7 + 1024 hh
        const bool have_cur = t >= 0 && t < ntiles;
        if (have_cur) {
            if (FAST) butterflies64_static<M1>(v); else butterflies64_masked(v, m1);
            // exchange 1: natural (swizzled) order, pairs (l = 0, 1) as one access
            {
                const int g0 = (tid >> 5) & 3;
#pragma unroll
                for (int h = 0; h < 32; ++h) {
                    const int g = (g0 | (h << 2)) & SW;
                    Pair pr; pr.x = v[2 * h]; pr.y = v[2 * h + 1];
                    *reinterpret_cast<Pair *>(buf + ((e1 ^ (g << 1)) + 256 * h)) = pr;
                }
            }
            group_sync();
            {
                Pair *row = reinterpret_cast<Pair *>(buf + 64 * tid);
                const int g = tid & SW;
#pragma unroll
                for (int j = 0; j < 32; ++j) { const Pair pr = row[j ^ g]; v[2 * j] = pr.x; v[2 * j + 1] = pr.y; }
                if (FAST) butterflies64_static<M2>(v); else butterflies64_masked(v, m2);
#pragma unroll
                for (int j = 0; j < 32; ++j) { Pair pr; pr.x = v[2 * j]; pr.y = v[2 * j + 1]; row[j ^ g] = pr; }
            }
            group_sync();
            {
                const int gw = ((tid >> 5) & 3) << 2;            // bits 8, 9 of e -> bits 2, 3 of the swizzle
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int b67 = i & 3, hh = i >> 2;
                    const int g = (b67 | gw) & SW;
                    const Pair pr = *reinterpret_cast<const Pair *>(buf + ((e3 + 64 * b67 + 1024 * hh) ^ (g << 1)));
                    v[2 * i] = pr.x; v[2 * i + 1] = pr.y;
                }
                if (FAST) butterflies64_static<M3>(v); else butterflies64_masked(v, m3);
            }
            if (PAIR) {
                // stage 13: this thread and the same thread of the other group hold the same positions of
                // the two tiles; swap through shared memory ([i][tid] pairs: consecutive lanes, no conflicts)
                group_sync();                                    // round 3 of this group has read its buffer
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    Pair pr; pr.x = v[2 * i]; pr.y = v[2 * i + 1];
                    reinterpret_cast<Pair *>(buf)[i * FW_THREADS + tid] = pr;
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const Pair pr = reinterpret_cast<const Pair *>(pbuf)[i * FW_THREADS + tid];
                    if (grp == 0) { v[2 * i] = v[2 * i] + pr.x; v[2 * i + 1] = v[2 * i + 1] + pr.y; }
                    else { v[2 * i] = pr.x - v[2 * i]; v[2 * i + 1] = pr.y - v[2 * i + 1]; }
                }
            }
            if (scale) {
#pragma unroll
                for (int r = 0; r < 64; ++r) v[r] *= post_scale;
            }
        }
        const bool have_next = tn < ntiles;
        const int64_t tt = PAIR ? 2 * t + grp : t, ttn = PAIR ? 2 * tn + grp : tn;
        const int64_t ob = (FAST && have_cur) ? tile_base(tt, ldo) + elem_off(e3) : 0;
        const int64_t ib = (FAST && have_next) ? tile_base(ttn, ldi) + elem_off(e1) : 0;
        // MODE 2 needs cbits <= 6 so that 64, 256 and 1024 are whole multiples of the contiguous piece
        const int64_t s64 = FAST ? elem_off(64) : 0, s256 = FAST ? elem_off(256) : 0, s1024 = FAST ? elem_off(1024) : 0;
        auto store_pair = [&](int i) {
            const int e = e3 + 64 * (i & 3) + 1024 * (i >> 2);
            const int64_t g = FAST ? ob + (i & 3) * s64 + (i >> 2) * s1024 : fwht_map(p, t, e, ldo);
            if (g >= 0) {
                if (FAST || vec) {
                    Pair pr; pr.x = v[2 * i]; pr.y = v[2 * i + 1];
                    *reinterpret_cast<Pair *>(out + g) = pr;
                } else {
                    out[g] = v[2 * i]; out[g + 1] = v[2 * i + 1];
                }
            }
        };
        // every path defines v[2i], v[2i+1] (nothing of the old tile stays live)
        auto load_pair = [&](int i) {
            const int en = e1 + 256 * i;
            const int64_t g = FAST ? ib + i * s256 : fwht_map(p, tn, en, ldi);
            if (g < 0) {
                v[2 * i] = T(0); v[2 * i + 1] = T(0);
            } else if (FAST || vec) {
                Elem<T>::load2(in + g, v[2 * i], v[2 * i + 1]);
            } else {
                v[2 * i] = Elem<T>::load1(in + g);
                v[2 * i + 1] = Elem<T>::load1(in + g + 1);   // bit 0 is always contiguous
            }
        };
        // the three cases are uniform over the CTA: no per-pair branches inside the unrolled loops
        if (clustered) asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        if (have_cur && have_next) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { store_pair(i); load_pair(i); }
        } else if (have_cur) {
#pragma unroll
            for (int i = 0; i < 32; ++i) store_pair(i);
        } else if (have_next) {
#pragma unroll
            for (int i = 0; i < 32; ++i) load_pair(i);
        }
        __syncthreads();        // all shared-memory reads of this tile are done before the next exchange overwrites it
    }
}

// n == 1: out = post_scale * a
template <typename T>
__global__ void scale_copy_kernel(const T *__restrict__ in, int64_t ldi, T *__restrict__ out, int64_t ldo,
                                  int64_t m, T s) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) out[i * ldo] = in[i * ldi] * s;
}

// RLA_FWHT_PLAN: 0 = generic passes only (13 bits, then 11 with 32-byte pieces: the round-2 baseline),
// 1 (default) = paired first pass + strided passes with >= 64-byte pieces and linear offsets
static int fwht_plan_mode() {
    const char *e = getenv("RLA_FWHT_PLAN");
    return e ? atoi(e) : 1;
}

template <typename T, int MODE, bool PAIR, int CB = 0>
static int fwht_launch(const T *src, int64_t lds, T *out, int64_t ldo, const FwhtPass &p, int64_t ntiles, T post,
                       int vec, cudaStream_t st) {
    auto kern = fwht_pass_kernel<T, MODE, PAIR, CB>;
    const int smem = (PAIR ? 2 : 1) * FW_TILE * (int)sizeof(T);
    RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t units = PAIR ? ntiles / 2 : ntiles;
    int64_t grid = std::min<int64_t>(units, (int64_t)sm_count() * (PAIR ? 1 : 3));
    int csize = 0;
    // clusters of 2: the two 64-byte halves of every 128-byte line are touched together
    // (measured at (64, 2^24): strided pass 6.4 ms unclustered, 4.4 / 5.5 / 6.0 ms for 2 / 4 / 8)
    if (MODE == 2) { const char *e = getenv("RLA_FWHT_CLUSTER"); csize = e ? atoi(e) : 2; }
    if (MODE == 2 && (csize == 2 || csize == 4 || csize == 8) && grid >= csize) {
        grid -= grid % csize;
        FwhtPass pc = p;
        pc.flags |= 8;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(FW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        RLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, src, lds, out, ldo, pc, units, post, vec));
    } else {
        kern<<<(unsigned)grid, PAIR ? 2 * FW_THREADS : FW_THREADS, smem, st>>>(src, lds, out, ldo, p, units, post, vec);
    }
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

template <typename T>
static int fwht_run(const T *a, int64_t m, int64_t n, int64_t lda, T *out, int64_t ldo, T post_scale, void *stream) {
    RLA_REQUIRE(m >= 0 && n >= 1, "rla_fwht: bad sizes");
    RLA_REQUIRE((n & (n - 1)) == 0, "rla_fwht: n=%lld is not a power of two", (long long)n);   // srht.py:110-111
    RLA_REQUIRE(lda >= n && ldo >= n, "rla_fwht: leading dimension smaller than n");
    if (m == 0) return RLA_OK;
    RLA_REQUIRE(a && out, "rla_fwht: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int d = ceil_log2_i64(n);
    RLA_REQUIRE(d <= 40, "rla_fwht: n too large");
    if (d == 0) {
        scale_copy_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(a, lda, out, ldo, m, post_scale);
        count_launch();
        RLA_CUDA_CHECK(cudaGetLastError());
        return RLA_OK;
    }
    const int plan = fwht_plan_mode();
    int done = 0;
    bool first = true;
    while (done < d) {
        const T *src = first ? a : out;
        const int64_t lds = first ? lda : ldo;
        const int vec = (reinterpret_cast<uintptr_t>(src) % (2 * sizeof(T)) == 0) && (lds % 2 == 0) &&
                        (reinterpret_cast<uintptr_t>(out) % (2 * sizeof(T)) == 0) && (ldo % 2 == 0) && (n >= 2);
        FwhtPass p;
        p.d = d; p.m = m; p.lo = done;
        p.flags = 0;
        int mode = 0;
        bool pair = false;
        if (first) {
            p.cbits = 0;
            p.nb = std::min(d, FW_TL);
            if (p.nb == FW_TL && vec) {
                mode = 1;
                // one more stage by pairing adjacent tiles when the row has it and there is a later pass to save
                if (plan >= 1 && d >= FW_TL + 1 && sizeof(T) == 8) { pair = true; p.nb = FW_TL + 1; }
            }
        } else {
            const int rem = d - done;
            if (plan >= 1 && vec) {
                // strided pass: as many bits as fit beside >= 8 contiguous elements (64 bytes for FP64);
                // a full tile (cbits + nb = 13, cbits <= 6) takes the linear-offset kernel
                p.nb = std::min(rem, FW_TL - 3);
                p.cbits = std::min(6, FW_TL - p.nb);
                if (p.cbits + p.nb == FW_TL) mode = 2;
            } else {
                p.cbits = 2;
                p.nb = std::min(rem, FW_TL - p.cbits);
            }
        }
        const int tile_nb = pair ? FW_TL : p.nb;               // bits the tile itself carries
        p.xb = FW_TL - p.cbits - tile_nb;
        p.nq = m << (d - p.cbits - tile_nb);
        const int64_t ntiles = (p.nq + (int64_t(1) << p.xb) - 1) >> p.xb;
        const bool last = done + p.nb >= d;
        const T post = last ? post_scale : T(1);
        FwhtPass pk = p;
        pk.nb = tile_nb;                                       // the kernel's masks / maps see the tile's own bits
        int rc;
        if (pair) rc = fwht_launch<T, 1, true>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else if (mode == 1) rc = fwht_launch<T, 1, false>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else if (mode == 2 && p.cbits == 3) rc = fwht_launch<T, 2, false, 3>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else if (mode == 2 && p.cbits == 4) rc = fwht_launch<T, 2, false, 4>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else if (mode == 2 && p.cbits == 5) rc = fwht_launch<T, 2, false, 5>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else if (mode == 2) rc = fwht_launch<T, 2, false, 6>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        else rc = fwht_launch<T, 0, false>(src, lds, out, ldo, pk, ntiles, post, vec, st);
        if (rc != RLA_OK) return rc;
        if (getenv("RLA_FWHT_TIMING")) {                  // developer aid: per-pass wall time
            static cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms = 0.f;
            if (!first) { cudaEventElapsedTime(&ms, e0, e1); }
            fprintf(stderr, "fwht pass lo=%d nb=%d cbits=%d mode=%d pair=%d tiles=%lld  %s%.3f ms\n", p.lo, p.nb, p.cbits, mode,
                    (int)pair, (long long)ntiles, first ? "(first pass, since previous event) " : "", ms);
            cudaEventRecord(e0, st);
        }
        done += p.nb;
        first = false;
    }
    return RLA_OK;
}

// ---- implicit SRHT adjoint: scatter v into the 2^d grid, transform, sign and truncate
template <typename T>
__global__ void adjoint_scatter_kernel(const T *__restrict__ v, int64_t ldv, int64_t k,
                                       const int32_t *__restrict__ order, const int64_t *__restrict__ idx,
                                       T *__restrict__ z, int64_t ldz) {
    // one thread per distinct index value: `order` sorts the samples by index, so a thread
    // that starts a run of equal indices sums the run in sample order (deterministic)
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (j >= k) return;
    const int64_t s = idx[order[j]];
    if (j > 0 && idx[order[j - 1]] == s) return;
    T acc = T(0);
    for (int64_t t = j; t < k && idx[order[t]] == s; ++t) acc += v[row * ldv + order[t]];
    z[row * ldz + s] = acc;
}

template <typename T>
__global__ void adjoint_finish_kernel(const T *__restrict__ z, int64_t ldz, const int8_t *__restrict__ signs,
                                      int64_t n, T value, T *__restrict__ out, int64_t ldo) {
    const int64_t row = blockIdx.y;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const T x = z[row * ldz + j] * value;
        out[row * ldo + j] = signs[j] < 0 ? -x : x;
    }
}

}  // namespace rla

using namespace rla;

extern "C" int rla_fwht_f64(const double *a, int64_t m, int64_t n, int64_t lda, double *out, int64_t ldo,
                            double post_scale, void *stream) {
    return fwht_run<double>(a, m, n, lda, out, ldo, post_scale, stream);
}
extern "C" int rla_fwht_f32(const float *a, int64_t m, int64_t n, int64_t lda, float *out, int64_t ldo,
                            float post_scale, void *stream) {
    return fwht_run<float>(a, m, n, lda, out, ldo, post_scale, stream);
}

extern "C" size_t rla_srht_adjoint_workspace_bytes(int64_t m, int64_t n) {
    if (m <= 0 || n <= 0) return 0;
    const int d = ceil_log2_i64(n);
    return (size_t)m * ((size_t)1 << d) * sizeof(double);
}

// `order_dev`: int32 permutation of [0, k) that sorts idx ascending (stable); built by the caller
// once per embedding (host argsort of the k indices).
extern "C" int rla_srht_adjoint_f64(const int8_t *signs, int64_t n, const int64_t *idx, const int32_t *order_dev,
                                    int64_t k, const double *v, int64_t m, int64_t ldv, double value,
                                    double *out, int64_t ldo, void *ws, size_t ws_bytes, void *stream) {
    RLA_REQUIRE(n >= 1 && k >= 0 && m >= 0 && ldv >= k && ldo >= n, "rla_srht_adjoint_f64: bad sizes");
    if (m == 0) return RLA_OK;
    RLA_REQUIRE(signs && out && ws && (k == 0 || (idx && order_dev && v)), "rla_srht_adjoint_f64: null pointer");
    const int d = ceil_log2_i64(n);
    const int64_t np2 = int64_t(1) << d;
    const size_t need = (size_t)m * np2 * sizeof(double);
    if (ws_bytes < need) return fail(RLA_ERR_WORKSPACE, "rla_srht_adjoint_f64: workspace %zu < %zu", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    double *z = static_cast<double *>(ws);
    RLA_CUDA_CHECK(cudaMemsetAsync(z, 0, need, st));
    for (int64_t r0 = 0; r0 < m; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, m - r0);
        if (k > 0) {
            dim3 g((unsigned)((k + 255) / 256), (unsigned)nr);
            adjoint_scatter_kernel<double><<<g, 256, 0, st>>>(v + r0 * ldv, ldv, k, order_dev, idx, z + r0 * np2, np2);
            count_launch();
        }
    }
    int rc = fwht_run<double>(z, m, np2, np2, z, np2, 1.0, stream);
    if (rc != RLA_OK) return rc;
    for (int64_t r0 = 0; r0 < m; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, m - r0);
        dim3 g((unsigned)std::min<int64_t>((n + 255) / 256, 4096), (unsigned)nr);
        adjoint_finish_kernel<double><<<g, 256, 0, st>>>(z + r0 * np2, np2, signs, n, value, out + r0 * ldo, ldo);
        count_launch();
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
