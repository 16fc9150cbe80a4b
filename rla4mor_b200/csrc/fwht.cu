// fwht.cu -- full (all outputs) Walsh-Hadamard transform along rows, natural order.
//
// Replaces fht_oop / fht_ip (rla/srht.py:99-134) and backs the implicit SRHT adjoint
// (SrhtEmbedding.apply_adjoint, rla/embeddings.py:175-178).
//
// A row of length 2^d is transformed in passes; each pass runs the 4096-element
// register/shared-memory tile of tile.cuh on up to 12 index bits:
//   pass 1  : bits [0, min(d,12))            tile = 4096 contiguous elements
//             (for d < 12 a tile packs 2^(12-d) rows);
//   pass >=2: 10 further bits each, the tile carrying 2 contiguous low bits (32-byte
//             sectors) plus 10 strided bits.
// Every pass reads and writes each element once (HBM-bound, d/12 .. d/10 passes); a
// tile reads all its elements before writing them, so passes after the first run in
// place on `out`.
#include "tile.cuh"
#include <algorithm>

namespace rla {

struct FwhtPass {
    int d;        // log2 of the row length
    int cbits;    // contiguous low bits carried (not transformed) inside a tile
    int nb;       // bits transformed in this pass
    int lo;       // position of the first transformed bit in the row index
    int xb;       // 12 - cbits - nb: spare tile bits enumerating other (row, index) combos
    int64_t m;    // rows
    int64_t nq;   // m * 2^(d - cbits - nb): number of (row, other-bits) combinations
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

// One pass: persistent 64-thread CTAs (one tile group each, four per SM), each looping over
// tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The loop is rotated so that there is ONE
// load site: the stores of tile t (read back from shared memory in the coalesced round-1
// layout) are interleaved, pair by pair, with the loads of the CTA's next tile into the
// registers they free, so the next tile's HBM latency overlaps this tile's stores and the
// other CTAs' butterflies.
// CONTIG: every tile is 4096 contiguous, 16-byte aligned elements of one row (first pass of
// rows of length >= 4096): the index arithmetic of fwht_map disappears.
template <typename T, bool CONTIG>
__global__ void __launch_bounds__(GROUP, 4) fwht_pass_kernel(const T *__restrict__ in, int64_t ldi,
                                                             T *__restrict__ out, int64_t ldo,
                                                             FwhtPass p, int64_t ntiles, T post_scale, int vec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);
    constexpr int M = Elem<T>::MASK;
    int tg = threadIdx.x;
    // stage mask over tile bits -> register-bit masks of the two rounds
    const int smask = ((1 << p.nb) - 1) << p.cbits;
    const int m1 = (smask & 1) | (((smask >> 7) & 31) << 1);    // round 1: tile bits 0, 7..11
    const int m2 = (smask >> 1) & 63;                            // round 2: tile bits 1..6
    const int tiles_per_row_log2 = p.d - TILE_LOG2;              // CONTIG only
    auto contig_off = [&](int64_t t, int64_t ld) -> int64_t {
        const int64_t row = t >> tiles_per_row_log2;
        return row * ld + ((t - (row << tiles_per_row_log2)) << TILE_LOG2) + 2 * tg;
    };
    T v[64];
    const int64_t G = gridDim.x;
    for (int64_t t = (int64_t)blockIdx.x - G; t < ntiles; t += G) {
        const int64_t tn = t + G;
        if (t >= 0) {
            butterflies64_masked(v, m1);
            asm volatile("" : "+r"(tg));
#pragma unroll
            for (int r = 0; r < 64; ++r) buf[r * 64 + (tg ^ (r & M))] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 64; ++r) v[r] = buf[tg * 64 + (r ^ (tg & M))];
            butterflies64_masked(v, m2);
#pragma unroll
            for (int r = 0; r < 64; ++r) buf[tg * 64 + (r ^ (tg & M))] = v[r] * post_scale;
            __syncthreads();
        }
        const bool have_next = tn < ntiles;
        const int64_t ob = (CONTIG && t >= 0) ? contig_off(t, ldo) : 0;
        const int64_t ib = (CONTIG && have_next) ? contig_off(tn, ldi) : 0;
#pragma unroll
        for (int h = 0; h < 32; ++h) {
            const int e = 128 * h + 2 * tg;
            if (t >= 0) {
                // back to the round-1 layout for coalesced 16-byte stores
                const int64_t g = CONTIG ? ob + 128 * h : fwht_map(p, t, e, ldo);
                if (g >= 0) {
                    const T a = buf[(2 * h) * 64 + (tg ^ ((2 * h) & M))];
                    const T b = buf[(2 * h + 1) * 64 + (tg ^ ((2 * h + 1) & M))];
                    if (CONTIG || vec) {
                        if (sizeof(T) == 8) *reinterpret_cast<double2 *>(out + g) = make_double2((double)a, (double)b);
                        else *reinterpret_cast<float2 *>(out + g) = make_float2((float)a, (float)b);
                    } else {
                        out[g] = a; out[g + 1] = b;
                    }
                }
            }
            // every path defines v[2h], v[2h+1] (nothing of the old tile stays live)
            const int64_t g = !have_next ? -1 : (CONTIG ? ib + 128 * h : fwht_map(p, tn, e, ldi));
            if (g < 0) {
                v[2 * h] = T(0); v[2 * h + 1] = T(0);
            } else if (CONTIG || vec) {
                Elem<T>::load2(in + g, v[2 * h], v[2 * h + 1]);
            } else {
                v[2 * h] = Elem<T>::load1(in + g);
                v[2 * h + 1] = Elem<T>::load1(in + g + 1);   // bit 0 is always contiguous
            }
        }
        __syncthreads();        // the store phase has finished reading buf before it is overwritten
    }
}

// n == 1: out = post_scale * a
template <typename T>
__global__ void scale_copy_kernel(const T *__restrict__ in, int64_t ldi, T *__restrict__ out, int64_t ldo,
                                  int64_t m, T s) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) out[i * ldo] = in[i * ldi] * s;
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
    const int smem = TILE * sizeof(T);
    RLA_CUDA_CHECK(cudaFuncSetAttribute(fwht_pass_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    RLA_CUDA_CHECK(cudaFuncSetAttribute(fwht_pass_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int done = 0;
    bool first = true;
    while (done < d) {
        FwhtPass p;
        p.d = d; p.m = m;
        p.cbits = first ? 0 : 2;
        p.nb = std::min(d - done, TILE_LOG2 - p.cbits);
        p.lo = done;
        p.xb = TILE_LOG2 - p.cbits - p.nb;
        p.nq = m << (d - p.cbits - p.nb);
        const int64_t ntiles = (p.nq + (int64_t(1) << p.xb) - 1) >> p.xb;
        const T *src = first ? a : out;
        const int64_t lds = first ? lda : ldo;
        const bool last = done + p.nb >= d;
        const int vec = (reinterpret_cast<uintptr_t>(src) % (2 * sizeof(T)) == 0) && (lds % 2 == 0) &&
                        (reinterpret_cast<uintptr_t>(out) % (2 * sizeof(T)) == 0) && (ldo % 2 == 0) && (n >= 2);
        const int64_t grid = std::min<int64_t>(ntiles, (int64_t)sm_count() * 4);
        const bool contig = first && p.nb == TILE_LOG2 && vec;
        if (contig)
            fwht_pass_kernel<T, true><<<(unsigned)grid, GROUP, smem, st>>>(src, lds, out, ldo, p, ntiles, last ? post_scale : T(1), vec);
        else
            fwht_pass_kernel<T, false><<<(unsigned)grid, GROUP, smem, st>>>(src, lds, out, ldo, p, ntiles, last ? post_scale : T(1), vec);
        count_launch();
        RLA_CUDA_CHECK(cudaGetLastError());
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
