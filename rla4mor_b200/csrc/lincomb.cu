// lincomb.cu -- out(m, n) = V(m, k) * X(k, n), all row-major, for SMALL m and k and LONG n:
//   * rb.lincomb(T.T), the full-space basis update after the sketched orthonormalisation
//     (mor/sketched_reductor.py:99-100): V = T^T (r' x r), X = rb (r x n);
//   * S_q <- S_q T and srb.lincomb(T.T) (:97, :104-108), V @ get_matrix() of the adjoints
//     (rla/embeddings.py:175-178).
//
// X is read ONCE per 128 rows of the result and `out` written once (the round-1 path transposed X
// into scratch and ran the split-n sketch GEMM over it: ~6 passes over k*n*8 bytes).  Bound:
// HBM for k below ~50 (8*(k + m)*n bytes), the FP64 pipe above (2*m*k*n flops; DFMA and DMMA
// share one datapath on this part, tools/micro/fp64_dual_probe.cu, so a register-tiled DFMA
// kernel has the same ceiling as a tensor-core one and needs no fragment shuffling of the
// n-contiguous operand).
//
// CTA = 256 threads, tile (TR*16) x 128 of `out`; thread (ty, tx) owns rows ty*TR..+TR-1 and the
// eight columns {2*tx + 32*j, +1}, j = 0..3 (16-byte shared-memory reads, conflict free).  The
// reduction dimension runs in chunks of 16 through two shared-memory stages: X[kc.., n tile] by
// cp.async (16 bytes per thread and request), V^T[kc.., rows] transposed on the way in.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace rla {

constexpr int LC_BN = 128;     // columns of the CTA tile
constexpr int LC_BK = 16;      // reduction chunk
constexpr int LC_THREADS = 256;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TR>
__global__ void __launch_bounds__(LC_THREADS, TR == 2 ? 2 : 1)
lincomb_kernel(const double *__restrict__ v, int64_t ldv, const double *__restrict__ x, int64_t ldx,
               double *__restrict__ out, int64_t ldo, int64_t m, int64_t k, int64_t n, int vec_ok, int mtiles) {
    constexpr int BM = TR * 16;
    constexpr int BMP = BM + 2;                        // padded: the transposing stores hit 8 bank pairs, not 1
    extern __shared__ __align__(16) unsigned char lc_smem[];
    double (*sx)[LC_BK][LC_BN] = reinterpret_cast<double (*)[LC_BK][LC_BN]>(lc_smem);
    double (*sv)[LC_BK][BMP] = reinterpret_cast<double (*)[LC_BK][BMP]>(lc_smem + 2 * LC_BK * LC_BN * sizeof(double));
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    // m tiles of one n tile are adjacent in launch order: X[:, n tile] is re-read from L2, not HBM
    const int64_t n0 = (int64_t)(blockIdx.x / mtiles) * LC_BN;
    const int64_t m0 = (int64_t)(blockIdx.x % mtiles) * BM;
    const int nchunks = (int)((k + LC_BK - 1) / LC_BK);

    double acc[TR][8];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    // stage loaders: X chunk = 16 rows x 128 columns = 1024 16-byte requests, 4 per thread
    auto load_x = [&](int st, int c) {
        const int64_t k0 = (int64_t)c * LC_BK;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int id = tid + q * LC_THREADS;       // 0..1023
            const int kk = id >> 6, cc = (id & 63) * 2;
            double *dst = &sx[st][kk][cc];
            const int64_t row = k0 + kk, col = n0 + cc;
            if (row < k && vec_ok && col + 1 < n) {
                cp_async16(dst, x + row * ldx + col);
            } else {
                dst[0] = (row < k && col < n) ? x[row * ldx + col] : 0.0;
                dst[1] = (row < k && col + 1 < n) ? x[row * ldx + col + 1] : 0.0;
            }
        }
    };
    // V^T chunk: sv[kk][row] = V[m0 + row][k0 + kk]  (BM * 16 values, TR per thread); V is tiny
    // and L2-resident, so plain loads (reads along k are contiguous per row)
    auto load_v = [&](int st, int c) {
        const int64_t k0 = (int64_t)c * LC_BK;
#pragma unroll
        for (int q = 0; q < TR; ++q) {
            const int id = tid + q * LC_THREADS;       // 0 .. BM*16-1
            const int row = id >> 4, kk = id & 15;
            const int64_t gr = m0 + row, gk = k0 + kk;
            sv[st][kk][row] = (gr < m && gk < k) ? v[gr * ldv + gk] : 0.0;
        }
    };

    load_x(0, 0);
    load_v(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            load_x(st ^ 1, c + 1);
            load_v(st ^ 1, c + 1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < LC_BK; ++kk) {
            double a[TR], b[8];
#pragma unroll
            for (int i = 0; i < TR; i += 2) {
                const double2 t = *reinterpret_cast<const double2 *>(&sv[st][kk][ty * TR + i]);
                a[i] = t.x; a[i + 1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 t = *reinterpret_cast<const double2 *>(&sx[st][kk][2 * tx + 32 * j]);
                b[2 * j] = t.x; b[2 * j + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
        const int64_t row = m0 + ty * TR + i;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + 2 * tx + 32 * j;
            double *dst = out + row * ldo + col;
            if (vec_ok && col + 1 < n) {
                *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][2 * j], acc[i][2 * j + 1]);
            } else {
                if (col < n) dst[0] = acc[i][2 * j];
                if (col + 1 < n) dst[1] = acc[i][2 * j + 1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tensor-pipe version (default): the vector FP64 pipe of this part issues 32 DFMA lanes per
// clock and SM, the DMMA path 64 (measured: tools/micro/fp64_dual_probe.cu, and the register-
// tiled kernel above saturates at 19 TFLOP/s), so everything but tiny k is bound by the DFMA
// rate unless it runs on mma.sync.m8n8k4.f64.
//   out tile (WM*WTM) x (WN*32), 8 warps, warp tile WTM x 32 (WTM = 64 or 32);
//   A fragment  a = V[i0 + g][k0 + t]   from the V chunk  [BM][16] (row stride 20 doubles)
//   B fragment  b = X[k0 + t][j0 + g]   from the X chunk  [16][BN] (row stride BN + 4)
//   D fragment  out[i0 + g][j0 + 2t .. +1]  -> one 16-byte store per lane
// (g = lane / 4, t = lane % 4).  Row strides = 4 (mod 16) doubles make every fragment load one
// conflict-free ld.shared.f64 per lane without any swizzle.  Chunks of 16 along k stream through
// a 3-stage cp.async ring; rows / columns outside the problem are zero-filled by the copy
// (src-size operand), so ragged m, k, n need no predicates in the math loop.
constexpr int LM_BK = 16;
constexpr int LM_STAGES = 3;
constexpr int LM_LDV = LM_BK + 4;

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void lc_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int WTM, int WM>
__global__ void __launch_bounds__(256, 1)
lincomb_mma_kernel(const double *__restrict__ v, int64_t ldv, const double *__restrict__ x, int64_t ldx,
                   double *__restrict__ out, int64_t ldo, int64_t m, int64_t k, int64_t n, int mtiles) {
    constexpr int WN = 8 / WM, BM = WM * WTM, BN = WN * 32, LDX = BN + 4;
    constexpr int MI = WTM / 8;                           // 8-row groups per warp tile
    constexpr int XS = LM_BK * LDX, VS = BM * LM_LDV;     // doubles per stage
    extern __shared__ __align__(16) unsigned char lm_smem[];
    double *sx = reinterpret_cast<double *>(lm_smem);
    double *sv = sx + LM_STAGES * XS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int64_t n0 = (int64_t)(blockIdx.x / mtiles) * BN;
    const int64_t m0 = (int64_t)(blockIdx.x % mtiles) * BM;
    const int nchunks = (int)((k + LM_BK - 1) / LM_BK);

    auto load_stage = [&](int st, int c) {
        const int64_t k0 = (int64_t)c * LM_BK;
        double *dx = sx + st * XS, *dv = sv + st * VS;
        // X chunk: 16 rows x BN/2 16-byte pieces
        for (int id = tid; id < LM_BK * (BN / 2); id += 256) {
            const int kk = id / (BN / 2), cc = (id % (BN / 2)) * 2;
            const int64_t row = k0 + kk, col = n0 + cc;
            const int bytes = (row < k && col < n) ? (col + 1 < n ? 16 : 8) : 0;
            cp_async16_zfill(dx + kk * LDX + cc, bytes ? x + row * ldx + col : x, bytes);
        }
        // V chunk: BM rows x 8 pieces
        for (int id = tid; id < BM * (LM_BK / 2); id += 256) {
            const int r = id / (LM_BK / 2), cc = (id % (LM_BK / 2)) * 2;
            const int64_t row = m0 + r, col = k0 + cc;
            const int bytes = (row < m && col < k) ? (col + 1 < k ? 16 : 8) : 0;
            cp_async16_zfill(dv + r * LM_LDV + cc, bytes ? v + row * ldv + col : v, bytes);
        }
    };

    double acc[MI][4][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < LM_STAGES - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<LM_STAGES - 2>();
        __syncthreads();                                  // chunk c has landed; everyone is done with chunk c - 1
        if (c + LM_STAGES - 1 < nchunks) load_stage((c + LM_STAGES - 1) % LM_STAGES, c + LM_STAGES - 1);
        cp_async_commit();
        const double *px = sx + (c % LM_STAGES) * XS + wn * 32 + g;
        const double *pv = sv + (c % LM_STAGES) * VS + (wm * WTM + g) * LM_LDV + t;
#pragma unroll
        for (int ks = 0; ks < LM_BK / 4; ++ks) {
            double a[MI], b[4];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = pv[i * 8 * LM_LDV + ks * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = px[(ks * 4 + t) * LDX + j * 8];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) lc_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = m0 + wm * WTM + 8 * i + g;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + wn * 32 + 8 * j + 2 * t;
            if (col + 1 < n) *reinterpret_cast<double2 *>(out + row * ldo + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (col < n) out[row * ldo + col] = acc[i][j][0];
        }
    }
}

template <int WTM, int WM>
static int lincomb_mma_launch(const double *v, int64_t m, int64_t k, int64_t ldv, const double *x, int64_t n, int64_t ldx,
                              double *out, int64_t ldo, cudaStream_t st) {
    constexpr int WN = 8 / WM, BM = WM * WTM, BN = WN * 32;
    const int64_t gx = (n + BN - 1) / BN, gy = (m + BM - 1) / BM;
    RLA_REQUIRE(gx * gy < (int64_t(1) << 31), "lincomb: problem too large");
    const int smem = (int)(LM_STAGES * (LM_BK * (BN + 4) + BM * LM_LDV) * sizeof(double));
    auto kern = lincomb_mma_kernel<WTM, WM>;
    RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)(gx * gy), 256, smem, st>>>(v, ldv, x, ldx, out, ldo, m, k, n, (int)gy);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

int lincomb_launch(const double *v, int64_t m, int64_t k, int64_t ldv, const double *x, int64_t n, int64_t ldx,
                   double *out, int64_t ldo, cudaStream_t st) {
    const int vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ldx % 2 == 0) &&
                       (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (ldo % 2 == 0);
    const bool v_ok = (reinterpret_cast<uintptr_t>(v) % 16 == 0) && (ldv % 2 == 0);
    static int force_fma = -1;
    if (force_fma < 0) { const char *e = getenv("RLA_LINCOMB_FMA"); force_fma = e ? atoi(e) : 0; }
    if (vec_ok && v_ok && !force_fma) {
        // tensor pipe: 128 x 128 tiles for tall coefficient blocks, 64 x 256 / 32 x 256 below
        if (m > 64) return lincomb_mma_launch<64, 2>(v, m, k, ldv, x, n, ldx, out, ldo, st);
        if (m > 32) return lincomb_mma_launch<64, 1>(v, m, k, ldv, x, n, ldx, out, ldo, st);
        return lincomb_mma_launch<32, 1>(v, m, k, ldv, x, n, ldx, out, ldo, st);
    }
    const int64_t gx = (n + LC_BN - 1) / LC_BN;
    // rows per thread: the smallest tile that covers m in one pass, 128 rows per pass beyond
    const int tr = m <= 32 ? 2 : (m <= 64 ? 4 : 8);
    const int64_t gy = (m + tr * 16 - 1) / (tr * 16);
    RLA_REQUIRE(gx * gy < (int64_t(1) << 31), "lincomb: problem too large");
    const unsigned grid = (unsigned)(gx * gy);
    const int smem = (int)(2 * LC_BK * LC_BN * sizeof(double) + 2 * LC_BK * (tr * 16 + 2) * sizeof(double));
    // more than the 48 KB default of dynamic shared memory for the two larger tiles
    if (tr == 2) {
        lincomb_kernel<2><<<grid, LC_THREADS, smem, st>>>(v, ldv, x, ldx, out, ldo, m, k, n, vec_ok, (int)gy);
    } else if (tr == 4) {
        RLA_CUDA_CHECK(cudaFuncSetAttribute(lincomb_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        lincomb_kernel<4><<<grid, LC_THREADS, smem, st>>>(v, ldv, x, ldx, out, ldo, m, k, n, vec_ok, (int)gy);
    } else {
        RLA_CUDA_CHECK(cudaFuncSetAttribute(lincomb_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        lincomb_kernel<8><<<grid, LC_THREADS, smem, st>>>(v, ldv, x, ldx, out, ldo, m, k, n, vec_ok, (int)gy);
    }
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

}  // namespace rla
