// gemm32.cu -- dense embeddings of FLOAT32 blocks on the generation-5 tensor cores:
//   Y(m, k) = U(m, n) * Theta(k, n)^T,  U float32, Theta generated in the kernel, Y float64.
//
// Replaces GaussianEmbedding.apply / BlockGaussianEmbedding.apply (rla/embeddings.py:250-254,
// 425-434) for float32 blocks; tolerance of the path: 1e-5 relative Frobenius (north_star).
//
// FP64 has no tcgen05 kind, float32 has kind::tf32 -- whose operands carry 10 explicit mantissa
// bits.  A sketch sums n ~ 2^22 products, so three things keep it inside 1e-5:
//   * Theta is tf32-EXACT by definition (rng kinds 1 = Rademacher and 2 = normals rounded to
//     TF32, rng.cuh): no error from Theta at all, and Theta is the same matrix whatever the
//     dtype of the block (FP64 blocks use the same rounded values in gemm.cu);
//   * U is split in two: the tensor core truncates an FP32 operand to TF32, so the raw tile IS
//     the high part hi = trunc(u), and lo = u - hi (exact in FP32, 13 significant bits, truncated
//     again to 11) is a second tile: two MMAs per k-step, residual 2^-21 per product;
//   * the FP32 accumulators in TMEM are flushed into FP64 registers every 64 terms.
//
// One CTA per SM, 128 x 128 tile of Y over one chunk of n (split-n as in gemm.cu: partial
// tiles to a workspace, summed in chunk order by the reduce kernel), 20 warps:
//   warps 0..7   flush: tcgen05.ld of the finished accumulator buffer, FP64 accumulate (64 per thread)
//   warps 8..17  producers: generate the 128 x 32 Theta tile of the stage (Philox + Box-Muller,
//                rng.cuh) into the 128-byte-swizzled K-major layout; split the U tile TMA delivered
//   warp 18      TMA: U tile of the stage (cp.async.bulk.tensor, 128-byte swizzle, zero fill)
//   warp 19      MMA: one thread issues tcgen05.mma.kind::tf32 (M = N = 128, K = 8), two per k-step,
//                tcgen05.commit releases the stage / hands the accumulator buffer to the flush warps
// Registers: setmaxnreg moves registers inside the pool the CTA got at launch (640 threads x 96 =
// 61 440; five warps per SM sub-partition cap the launch at 96): 8 flush warps at 176 + 12 small
// warps at 40 = 60 416, and per sub-partition (warp w lives on w % 4) 2 x 176 + 3 x 40 warps.
// (A first version asked for more than its pool held: the flush warps waited for ever.)
// The kernel is bound by the generator (1024 Philox blocks per stage against 512 cycles of MMA):
// hence ten producer warps.  Measured dead end: letting the (mostly idle) flush warps do the split
// of the U tile, with four TMEM buffers so that no flush is ever waited for: 146 vs 157 TFLOP/s.
// Two accumulator buffers in TMEM (2 x 128 columns): MMAs of chunk c + 1 run while chunk c is flushed.
#include "common.cuh"
#include "rng.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <algorithm>
#include <stdlib.h>

namespace rla {

constexpr int TK = 32;                 // floats per row per stage (128 bytes = swizzle span)
constexpr int TSTAGES = 4;
constexpr int TBM = 128, TBN = 128;
constexpr int TFLUSH = 8, TPROD = 10;  // warps
constexpr int TTHREADS = (TFLUSH + TPROD + 2) * 32;
constexpr int TCHUNK = 2;              // stages per accumulator chunk: 64 terms in FP32, then FP64
constexpr int T_TILE_BYTES = TBM * TK * 4;            // 16 KB
constexpr int T_STAGE_BYTES = 3 * T_TILE_BYTES;       // U raw | U lo | Theta
constexpr int T_TMEM_COLS = 256;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void t_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void t_tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(x), "r"(y) : "memory");
}
// K-major operand tile in the canonical 128-byte-swizzle layout (what TMA SWIZZLE_128B writes):
// rows of 128 bytes, groups of 8 rows 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
           (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// D (FP32, TMEM) (+)= A (tf32, smem) * B (tf32, smem)^T
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// the same arrive on the barrier at this offset in BOTH CTAs of a pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(s32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t t_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void t_dsmem_push(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t t_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void t_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct Gemm32Args {
    int64_t m, k, n;          // Y is m x k, reduction over n
    int64_t kper;             // 32-wide k blocks per chunk
    int64_t nchunks;
    int mtiles, ntiles;
    double *ws;               // [nchunks][m][k] partial sketches
    uint64_t seed;
    int64_t row0, col0;       // offsets of this block inside the virtual Theta
};

// CL = 2: the CTAs of a cluster pair work on the same Theta tile (same sketch rows, neighbouring
// row tiles of U); each generates half of its rows and pushes them into the peer's tile with DSMEM
// bulk copies that signal the peer's full barrier (complete_tx), and a stage is released on BOTH
// CTAs by each MMA warp (multicast commit): the generator, which bounds this kernel, runs once
// per 256 rows of U.
template <int KIND, int CL>
__global__ void __launch_bounds__(TTHREADS, 1)
sketch_gemm_tf32_kernel(const __grid_constant__ CUtensorMap mapU, const Gemm32Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_a[TSTAGES], full[TSTAGES], empty[TSTAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // blockIdx.x = rank + CL * (ntile + ntiles * (mgroup + mgroups * chunk))
    const uint32_t rank = (CL > 1) ? t_ctarank() : 0u;
    int64_t b = blockIdx.x / CL;
    const int ntile = (int)(b % a.ntiles); b /= a.ntiles;
    const int mgrp = a.mtiles / CL;
    const int mtile = (int)(b % mgrp) * CL + (int)rank; b /= mgrp;
    const int64_t chunk = b;
    const int64_t kb0 = chunk * a.kper;
    const int64_t nk32 = (a.n + TK - 1) / TK;
    const int64_t kb1 = (kb0 + a.kper < nk32) ? kb0 + a.kper : nk32;
    const int iters = (int)(kb1 - kb0);
    const int m0 = mtile * TBM, n0 = ntile * TBN;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES; ++s) {
            t_mbar_init(&full_a[s], 1);
            t_mbar_init(&full[s], TPROD + (CL > 1 ? 1 : 0));       // + the expect_tx of the peer's half of Theta
            t_mbar_init(&empty[s], CL);                            // the MMA warps of every CTA of the pair
        }
        t_mbar_init(&tmem_full[0], 1); t_mbar_init(&tmem_full[1], 1);
        t_mbar_init(&tmem_empty[0], TFLUSH); t_mbar_init(&tmem_empty[1], TFLUSH);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(T_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    if (CL > 1) t_cluster_sync(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < TFLUSH) {
        // ------------------------------------------------------------ flush warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
        const int q = warp & 3, h = warp >> 2;            // TMEM lane quarter (= warp % 4), column half
        double acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0;
        const int nloc = (iters + TCHUNK - 1) / TCHUNK;
        for (int c = 0; c < nloc; ++c) {
            const int buf = c & 1;
            t_mbar_wait(&tmem_full[buf], (uint32_t)(c >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TBN + 64 * h);
            uint32_t v[2][32];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]), "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]),
                      "=r"(v[g][8]), "=r"(v[g][9]), "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]), "=r"(v[g][15]),
                      "=r"(v[g][16]), "=r"(v[g][17]), "=r"(v[g][18]), "=r"(v[g][19]), "=r"(v[g][20]), "=r"(v[g][21]), "=r"(v[g][22]), "=r"(v[g][23]),
                      "=r"(v[g][24]), "=r"(v[g][25]), "=r"(v[g][26]), "=r"(v[g][27]), "=r"(v[g][28]), "=r"(v[g][29]), "=r"(v[g][30]), "=r"(v[g][31])
                    : "r"(taddr + 32 * g) : "memory");
                if (g == 0) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] += (double)__uint_as_float(v[0][j]);
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the values are in registers: the buffer can take the MMAs of chunk c + 2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) t_mbar_arrive(&tmem_empty[buf]);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[32 + j] += (double)__uint_as_float(v[1][j]);
        }
        // partial tile -> workspace [chunk][m][k]: thread = one row of Y, 64 consecutive columns
        const int64_t row = m0 + 32 * q + lane;
        if (row < a.m) {
            double *wsp = a.ws + (chunk * a.m + row) * a.k;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int64_t col = n0 + 64 * h + j;
                if (col < a.k) wsp[col] = acc[j];
            }
        }
    } else if (warp < TFLUSH + TPROD) {
        // ------------------------------------------------------------ producers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int p = tid - TFLUSH * 32;
        const int pw = warp - TFLUSH;                      // producer warp
        constexpr int NPT = TPROD * 32;                    // producer threads: a multiple of 8
        // Theta tile: 128 rows x 8 chunks of four floats; this CTA generates ROWS = 128 / CL of them
        // (rows [rank * ROWS, + ROWS)): item = p + NPT j is chunk c = p & 7 of row item >> 3 of that
        // slice, so warp pw owns the 4 rows 4 pw + (NPT / 8) j .. + 3 per j: 512 contiguous bytes
        constexpr int ROWS = TBN / CL;
        constexpr int NJ = (ROWS * 8 + NPT - 1) / NPT;
        const int c = p & 7;
        uint32_t peer_b = 0, peer_full = 0;
        if (CL > 1) {
            peer_b = t_mapa(s32(smem) + 2 * T_TILE_BYTES, rank ^ 1u);
            peer_full = t_mapa(s32(&full[0]), rank ^ 1u);
        }
        uint64_t qv = ((uint64_t)(a.col0 + kb0 * TK) >> 2) + (uint64_t)c;
        for (int it = 0; it < iters; ++it, qv += TK / 4) {
            const int s = it % TSTAGES;
            if (it >= TSTAGES) t_mbar_wait(&empty[s], (uint32_t)((it / TSTAGES) - 1) & 1u);
            unsigned char *st = smem + s * T_STAGE_BYTES;
            if (CL > 1 && p == 0) t_mbar_expect_tx(&full[s], (CL - 1) * ROWS * 128);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int rl = (p + NPT * j) >> 3;                                  // row inside the slice
                const int r = (int)rank * ROWS + rl;                                // row inside the tile
                if (rl < ROWS && n0 + r < a.k) {
                    float f[4];
                    theta4f<KIND>(a.seed, (uint32_t)(a.row0 + n0 + r), qv, f);
                    *reinterpret_cast<float4 *>(st + 2 * T_TILE_BYTES + r * 128 + ((c ^ (r & 7)) << 4)) =
                        make_float4(f[0], f[1], f[2], f[3]);
                }
            }
            if (CL > 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int rl0 = 4 * pw + (NPT / 8) * j;                     // first of this warp's 4 rows
                        if (rl0 < ROWS) {
                            const uint32_t off = (uint32_t)(s * T_STAGE_BYTES + ((int)rank * ROWS + rl0) * 128);
                            t_dsmem_push(peer_b + off, s32(smem) + 2 * T_TILE_BYTES + off, 512u, peer_full + (uint32_t)s * 8u);
                        }
                    }
                }
            }
            // split the U tile: raw stays (the tensor core truncates it to its TF32 high part),
            // lo = u - trunc(u) goes to the second tile at the same (swizzled) offsets
            t_mbar_wait(&full_a[s], (uint32_t)(it / TSTAGES) & 1u);
#pragma unroll
            for (int j = 0; j < (T_TILE_BYTES / 16 + NPT - 1) / NPT; ++j) {
                const int off = (p + NPT * j) * 16;
                if (off < T_TILE_BYTES) {
                    const float4 u = *reinterpret_cast<const float4 *>(st + off);
                    float4 lo;
                    lo.x = u.x - __uint_as_float(__float_as_uint(u.x) & 0xffffe000u);
                    lo.y = u.y - __uint_as_float(__float_as_uint(u.y) & 0xffffe000u);
                    lo.z = u.z - __uint_as_float(__float_as_uint(u.z) & 0xffffe000u);
                    lo.w = u.w - __uint_as_float(__float_as_uint(u.w) & 0xffffe000u);
                    *reinterpret_cast<float4 *>(st + T_TILE_BYTES + off) = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) t_mbar_arrive(&full[s]);
        }
    } else if (warp == TFLUSH + TPROD) {
        // ------------------------------------------------------------ TMA
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % TSTAGES;
                if (it >= TSTAGES) t_mbar_wait(&empty[s], (uint32_t)((it / TSTAGES) - 1) & 1u);
                t_mbar_expect_tx(&full_a[s], T_TILE_BYTES);
                t_tma_load_2d(smem + s * T_STAGE_BYTES, &mapU, &full_a[s], (int)((kb0 + it) * TK), m0);
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issue
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            // kind::tf32, FP32 accumulate, A and B K-major, M = N = 128 (cute::UMMA::InstrDescriptor)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t sbase = s32(smem);
            for (int it = 0; it < iters; ++it) {
                const int s = it % TSTAGES;
                const int c = it / TCHUNK, buf = c & 1;
                const bool first = (it % TCHUNK) == 0;
                if (first && c >= 2) {
                    t_mbar_wait(&tmem_empty[buf], (uint32_t)((c >> 1) - 1) & 1u);
                    tc_fence_after();
                }
                t_mbar_wait(&full[s], (uint32_t)(it / TSTAGES) & 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * TBN);
                const uint32_t st = sbase + (uint32_t)(s * T_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < TK / 8; ++k) {
                    const uint64_t da = umma_desc(st + k * 32), dl = umma_desc(st + T_TILE_BYTES + k * 32);
                    const uint64_t db = umma_desc(st + 2 * T_TILE_BYTES + k * 32);
                    umma_tf32(d, da, db, idesc, (first && k == 0) ? 0u : 1u);
                    umma_tf32(d, dl, db, idesc, 1u);
                }
                if (CL > 1) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);   // stage free when these MMAs are done
                if ((it % TCHUNK) == TCHUNK - 1 || it == iters - 1) umma_commit(&tmem_full[buf]);
            }
        }
    }
    tc_fence_before();
    // no CTA of a pair may exit while its peer can still push into its shared memory or signal its barriers
    if (CL > 1) t_cluster_sync(); else __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T_TMEM_COLS) : "memory");
    }
}

// y[c, i] = (accumulate ? y[c, i] : 0) + scale * sum_chunks ws[chunk][c][i]
__global__ void gemm32_reduce_kernel(const double *__restrict__ ws, int64_t nchunks, int64_t m, int64_t k, double scale,
                                     double *__restrict__ y, int64_t ldy, int accumulate) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= m * k) return;
    const int64_t row = idx / k, col = idx % k;
    double s = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) s += ws[c * m * k + idx];
    double *dst = y + row * ldy + col;
    *dst = accumulate ? (*dst + scale * s) : scale * s;
}

static PFN_cuTensorMapEncodeTiled get_encode32() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

struct Gemm32Plan {
    int mtiles, ntiles;
    int64_t kper, nchunks;
};

static Gemm32Plan plan_gemm32(int64_t m, int64_t k, int64_t n) {
    Gemm32Plan p;
    p.mtiles = (int)((m + TBM - 1) / TBM);
    p.ntiles = (int)((k + TBN - 1) / TBN);
    const int64_t nk32 = (n + TK - 1) / TK;
    const int64_t tiles = (int64_t)p.mtiles * p.ntiles;
    const int64_t sms = sm_count();
    const int64_t target = sms * 8;
    // every chunk a whole number of accumulator chunks and at least 64 k-blocks unless the problem is tiny
    const int64_t cmax = std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(1, nk32 / 64), (target + tiles - 1) / tiles));
    const int64_t cmin = std::max<int64_t>(1, cmax / 4);
    int64_t chunks = cmax;
    double best = -1.0;
    for (int64_t c = cmin; c <= cmax; ++c) {
        int64_t kper = (nk32 + c - 1) / c;
        kper = (kper + TCHUNK - 1) / TCHUNK * TCHUNK;
        const int64_t nc = (nk32 + kper - 1) / kper;
        const int64_t waves = (nc * tiles + sms - 1) / sms;
        const double eff = (double)nk32 * tiles / ((double)waves * sms * kper);
        if (eff > best + 2e-3) { best = eff; chunks = c; }
    }
    p.kper = (nk32 + chunks - 1) / chunks;
    p.kper = (p.kper + TCHUNK - 1) / TCHUNK * TCHUNK;
    p.nchunks = (nk32 + p.kper - 1) / p.kper;
    return p;
}

}  // namespace rla

using namespace rla;

extern "C" size_t rla_gemm32_workspace_bytes(int64_t m, int64_t k, int64_t n) {
    if (m <= 0 || k <= 0 || n <= 0) return 0;
    const Gemm32Plan p = plan_gemm32(m, k, n);
    return (size_t)p.nchunks * m * k * sizeof(double);
}

extern "C" int rla_embed_apply_rng_f32(uint64_t seed, int kind, double scale, int64_t row0, int64_t k_blk, int64_t col0,
                                       int64_t n, const float *u, int64_t m, int64_t ldu, double *y, int64_t ldy,
                                       int accumulate, void *ws, size_t ws_bytes, void *stream) {
    RLA_REQUIRE(kind == 1 || kind == 2, "rla_embed_apply_rng_f32: kind must be 1 (rademacher) or 2 (normal rounded to TF32): "
                                        "Theta has to be exact in TF32");
    RLA_REQUIRE(k_blk >= 0 && n >= 1 && m >= 0 && ldu >= n && ldy >= k_blk && row0 >= 0 && col0 >= 0,
                "rla_embed_apply_rng_f32: bad sizes");
    RLA_REQUIRE(col0 % 32 == 0, "rla_embed_apply_rng_f32: col0 must be a multiple of 32");
    RLA_REQUIRE(row0 + k_blk <= (int64_t(1) << 32), "rla_embed_apply_rng_f32: more than 2^32 sketch rows");
    if (m == 0 || k_blk == 0) return RLA_OK;
    RLA_REQUIRE(u && y && ws, "rla_embed_apply_rng_f32: null pointer");
    RLA_REQUIRE(reinterpret_cast<uintptr_t>(u) % 16 == 0 && ldu % 4 == 0,
                "rla_embed_apply_rng_f32: u must be 16-byte aligned with a leading dimension that is a multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const Gemm32Plan p = plan_gemm32(m, k_blk, n);
    const size_t need = (size_t)p.nchunks * m * k_blk * sizeof(double);
    if (ws_bytes < need) return fail(RLA_ERR_WORKSPACE, "rla_embed_apply_rng_f32: workspace %zu < %zu bytes", ws_bytes, need);
    PFN_cuTensorMapEncodeTiled enc = get_encode32();
    if (!enc) return fail(RLA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap mu;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    cuuint64_t gstr[1] = {(cuuint64_t)ldu * 4};
    cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)TBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mu, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(u), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RLA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    Gemm32Args a;
    a.m = m; a.k = k_blk; a.n = n; a.kper = p.kper; a.nchunks = p.nchunks; a.mtiles = p.mtiles; a.ntiles = p.ntiles;
    a.ws = static_cast<double *>(ws); a.seed = seed; a.row0 = row0; a.col0 = col0;
    const int64_t grid = (int64_t)p.mtiles * p.ntiles * p.nchunks;
    RLA_REQUIRE(grid < (int64_t(1) << 31), "rla_embed_apply_rng_f32: grid too large");
    const int smem = TSTAGES * T_STAGE_BYTES + 1024;
    static const int cl_env = getenv("RLA_GEMM32_CL") ? atoi(getenv("RLA_GEMM32_CL")) : 0;
    const int cl = (p.mtiles % 2 == 0 && cl_env != 1) ? 2 : 1;
    const void *fn = kind == 1 ? (cl == 2 ? (const void *)sketch_gemm_tf32_kernel<1, 2> : (const void *)sketch_gemm_tf32_kernel<1, 1>)
                               : (cl == 2 ? (const void *)sketch_gemm_tf32_kernel<2, 2> : (const void *)sketch_gemm_tf32_kernel<2, 1>);
    RLA_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(TTHREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl > 1 ? 1 : 0;
    void *kargs[] = {(void *)&mu, (void *)&a};
    RLA_CUDA_CHECK(cudaLaunchKernelExC(&cfg, fn, kargs));
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    const int64_t tot = m * k_blk;
    gemm32_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a.ws, p.nchunks, m, k_blk, scale, y, ldy, accumulate);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
