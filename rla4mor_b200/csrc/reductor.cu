// reductor.cu -- the small operations around the sketch that SketchedReductor needs
// (mor/sketched_reductor.py): CSR SpMM in front of the sketch (:69-70), Gram-Schmidt of
// the sketched basis (:94), one-sided Jacobi SVD of a k x m sketch, sketched residual
// norm (:216-219), and V * Theta for the adjoint / basis update (:99-100).
// These are latency- or HBM-bound helpers next to the two headline kernels (srht.cu,
// gemm.cu); they are written for correctness and determinism first.
#include "common.cuh"
#include <algorithm>

namespace rla {

// ------------------------------------------------------------------- CSR SpMM
// out[c, i] = sum_j A[i, j] u[c, j].  Thread = one matrix row i for VB vectors c at once,
// so the CSR arrays are streamed m / VB times and out is written coalesced along i.
template <int VB>
__global__ void spmm_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                const double *__restrict__ val, int64_t n_rows, const double *__restrict__ u,
                                int64_t m, int64_t ldu, double *__restrict__ out, int64_t ldo) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.y * VB;
    if (i >= n_rows) return;
    double acc[VB];
#pragma unroll
    for (int v = 0; v < VB; ++v) acc[v] = 0.0;
    const int64_t e1 = rowptr[i + 1];
    for (int64_t e = rowptr[i]; e < e1; ++e) {
        const double a = val[e];
        const int64_t j = col[e];
#pragma unroll
        for (int v = 0; v < VB; ++v)
            if (c0 + v < m) acc[v] = fma(a, __ldg(u + (c0 + v) * ldu + j), acc[v]);
    }
#pragma unroll
    for (int v = 0; v < VB; ++v)
        if (c0 + v < m) out[(c0 + v) * ldo + i] = acc[v];
}

// ------------------------------------------------------------ tiled transpose
__global__ void transpose_kernel(const double *__restrict__ in, int64_t rows, int64_t cols, int64_t ldi,
                                 double *__restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t r = r0 + y, c = c0 + threadIdx.x;
        tile[y][threadIdx.x] = (r < rows && c < cols) ? in[r * ldi + c] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t c = c0 + y, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * ldo + r] = tile[threadIdx.x][y];
    }
}

// --------------------------------------------------------- block reductions
__device__ __forceinline__ double block_sum(double v, double *red) {
    // deterministic: fixed shuffle tree inside a warp, then warp partials in lane order
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// ------------------------------------------------------------- Gram-Schmidt
// pyMOR's gram_schmidt(A, offset, return_R=True) on the rows of A (r x k): modified
// Gram-Schmidt, re-iterated while the norm drops below reiteration_threshold * old norm;
// R[j, i] += <A_j, A_i>, R[i, i] = final norm; a row whose norm falls below
// rtol * initial (or atol initially) is flagged as removed and skipped afterwards.
// One CTA of W warps; the row being orthogonalised lives in registers (EPT <= 16 elements
// per thread), the next basis row is prefetched while the current dot product is reduced
// (warp shuffles + one barrier per step), so a step costs a few hundred cycles.
template <int EPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
gram_schmidt_kernel(double *__restrict__ A, int64_t r, int64_t k, int64_t lda, int64_t offset,
                    double *__restrict__ R, int32_t *__restrict__ flags, double atol, double rtol, double thr) {
    __shared__ double red[2][32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    int par = 0;
    // deterministic block sum: shuffle tree in the warp, warp partials added in warp order;
    // the partial buffer alternates so one barrier per reduction suffices
    auto bsum = [&](double v) -> double {
#pragma unroll
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[par][warp] = v;
        __syncthreads();
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[par][w];
        par ^= 1;
        return s;
    };
    for (int64_t i = tid; i < r * r; i += nthr) R[i] = ((i / r) == (i % r)) ? 1.0 : 0.0;
    for (int64_t i = tid; i < r; i += nthr) flags[i] = 0;
    __syncthreads();
    for (int64_t i = offset; i < r; ++i) {
        double x[EPT];
        double ss = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int64_t c = tid + (int64_t)e * nthr;
            x[e] = c < k ? A[i * lda + c] : 0.0;
            ss = fma(x[e], x[e], ss);
        }
        const double initial = sqrt(bsum(ss));
        if (initial <= atol) {
            if (tid == 0) flags[i] = 1;
            __syncthreads();
            continue;
        }
        double norm = initial;
        bool removed = false;
        if (i > 0) {
            while (true) {
                double aj[EPT], an[EPT];
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int64_t c = tid + (int64_t)e * nthr;
                    an[e] = c < k ? A[c] : 0.0;                     // row 0
                }
                for (int64_t j = 0; j < i; ++j) {
#pragma unroll
                    for (int e = 0; e < EPT; ++e) aj[e] = an[e];
                    if (j + 1 < i) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) {
                            const int64_t c = tid + (int64_t)e * nthr;
                            an[e] = c < k ? A[(j + 1) * lda + c] : 0.0;
                        }
                    }
                    if (flags[j]) continue;
                    double d = 0.0;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) d = fma(aj[e], x[e], d);
                    const double p = bsum(d);
#pragma unroll
                    for (int e = 0; e < EPT; ++e) x[e] = fma(-p, aj[e], x[e]);
                    if (tid == 0) R[j * r + i] += p;
                }
                ss = 0.0;
#pragma unroll
                for (int e = 0; e < EPT; ++e) ss = fma(x[e], x[e], ss);
                const double old = norm;
                norm = sqrt(bsum(ss));
                if (norm <= rtol * initial) { removed = true; break; }
                if (!(norm < thr * old)) break;
            }
        }
        if (removed) {
            if (tid == 0) flags[i] = 1;
        } else {
            const double inv = 1.0 / norm;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int64_t c = tid + (int64_t)e * nthr;
                if (c < k) A[i * lda + c] = x[e] * inv;
            }
            if (tid == 0) R[i * r + i] = norm;
        }
        __threadfence_block();
        __syncthreads();
    }
}

// ------------------------------------------------- one-sided Jacobi (Hestenes) SVD
// The k x m sketch is held as m rows of length k (row p = column p of the k x m matrix).
// One launch = one round of a round-robin ordering: m/2 disjoint pairs, one CTA each.
__global__ void __launch_bounds__(256)
jacobi_round_kernel(double *__restrict__ A, int64_t k, int64_t lda, double *__restrict__ V, int64_t m,
                    const int32_t *__restrict__ pairs, int npairs, double tol, int32_t *__restrict__ rotated) {
    __shared__ double red[8];
    const int pr = blockIdx.x;
    if (pr >= npairs) return;
    int p = pairs[2 * pr], q = pairs[2 * pr + 1];
    if (p < 0 || q < 0 || p >= m || q >= m) return;
    if (p > q) { const int t = p; p = q; q = t; }
    double *ap = A + p * lda, *aq = A + q * lda;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const double x = ap[i], y = aq[i];
        a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
    }
    a = block_sum(a, red); b = block_sum(b, red); g = block_sum(g, red);
    if (fabs(g) <= tol * sqrt(a * b) || g == 0.0) return;
    const double zeta = (b - a) / (2.0 * g);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const double x = ap[i], y = aq[i];
        ap[i] = c * x - s * y;
        aq[i] = s * x + c * y;
    }
    if (V) {
        double *vp = V + p * m, *vq = V + q * m;
        for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
            const double x = vp[i], y = vq[i];
            vp[i] = c * x - s * y;
            vq[i] = s * x + c * y;
        }
    }
    if (threadIdx.x == 0) atomicAdd(rotated, 1);
}

__global__ void row_norms_kernel(const double *__restrict__ A, int64_t k, int64_t lda, double *__restrict__ s) {
    __shared__ double red[8];
    const double *a = A + (int64_t)blockIdx.x * lda;
    double v = 0.0;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) v = fma(a[i], a[i], v);
    v = block_sum(v, red);
    if (threadIdx.x == 0) s[blockIdx.x] = sqrt(v);
}

__global__ void eye_kernel(double *V, int64_t m) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m * m) V[i] = ((i / m) == (i % m)) ? 1.0 : 0.0;
}

// ------------------------------------------------------- sketched residual norm
// || sum_q th[q] S_q a - sum_p tr[p] b_p ||_2, S_q: k x r row-major blocks
__global__ void __launch_bounds__(256, 1)
residual_norm_kernel(const double *__restrict__ S, int64_t Q, int64_t k, int64_t r, const double *__restrict__ th,
                     const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                     const double *__restrict__ tr, double *__restrict__ out) {
    __shared__ double red[8];
    double ss = 0.0;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        double res = 0.0;
        for (int64_t q = 0; q < Q; ++q) {
            const double *row = S + (q * k + i) * r;
            double d = 0.0;
            for (int64_t j = 0; j < r; ++j) d = fma(row[j], a[j], d);
            res = fma(th[q], d, res);
        }
        for (int64_t p = 0; p < P; ++p) res = fma(-tr[p], b[p * k + i], res);
        ss = fma(res, res, ss);
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) *out = sqrt(ss);
}

}  // namespace rla

using namespace rla;

extern "C" int rla_spmm_csr_f64(const int64_t *rowptr, const int32_t *col, const double *val, int64_t n_rows,
                                int64_t n_cols, const double *u, int64_t m, int64_t ldu, double *out, int64_t ldo,
                                void *stream) {
    RLA_REQUIRE(n_rows >= 0 && n_cols >= 0 && m >= 0 && ldu >= n_cols && ldo >= n_rows, "rla_spmm_csr_f64: bad sizes");
    if (m == 0 || n_rows == 0) return RLA_OK;
    RLA_REQUIRE(rowptr && u && out, "rla_spmm_csr_f64: null pointer");
    constexpr int VB = 8;
    for (int64_t c0 = 0; c0 < m; c0 += 65535LL * VB) {
        const int64_t mc = std::min<int64_t>(65535LL * VB, m - c0);
        dim3 grid((unsigned)((n_rows + 127) / 128), (unsigned)((mc + VB - 1) / VB));
        spmm_csr_kernel<VB><<<grid, 128, 0, (cudaStream_t)stream>>>(rowptr, col, val, n_rows, u + c0 * ldu, mc, ldu,
                                                                     out + c0 * ldo, ldo);
        count_launch();
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

namespace rla {
int lincomb_launch(const double *v, int64_t m, int64_t k, int64_t ldv, const double *x, int64_t n, int64_t ldx,
                   double *out, int64_t ldo, cudaStream_t st);
}

// kept for ABI stability: the tall-skinny kernel needs no scratch
extern "C" size_t rla_gemm_nn_workspace_bytes(int64_t m, int64_t k, int64_t n) {
    (void)m; (void)k; (void)n;
    return 0;
}

// out(m, n) = V(m, k) * Theta(k, n) with m, k small and n long: csrc/lincomb.cu (Theta read once
// per 128 rows of the result, no transposed copy)
extern "C" int rla_gemm_nn_f64(const double *v, int64_t m, int64_t k, int64_t ldv, const double *theta, int64_t n,
                               int64_t ldt, double *out, int64_t ldo, void *ws, size_t ws_bytes, void *stream) {
    (void)ws; (void)ws_bytes;
    RLA_REQUIRE(m >= 0 && k >= 1 && n >= 0 && ldv >= k && ldt >= n && ldo >= n, "rla_gemm_nn_f64: bad sizes");
    if (m == 0 || n == 0) return RLA_OK;
    RLA_REQUIRE(v && theta && out, "rla_gemm_nn_f64: null pointer");
    return lincomb_launch(v, m, k, ldv, theta, n, ldt, out, ldo, (cudaStream_t)stream);
}

extern "C" int rla_gram_schmidt_f64(double *a, int64_t r, int64_t k, int64_t lda, int64_t offset, double *R,
                                    int32_t *flags, double atol, double rtol, double thr, void *stream) {
    RLA_REQUIRE(r >= 0 && k >= 1 && lda >= k && offset >= 0, "rla_gram_schmidt_f64: bad sizes");
    if (r == 0) return RLA_OK;
    RLA_REQUIRE(a && R && flags, "rla_gram_schmidt_f64: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // about 8 elements per thread, at most 512 threads up to k = 8192 (registers hold the
    // row, the current and the prefetched basis row), 1024 threads beyond
    RLA_REQUIRE(k <= 16384, "rla_gram_schmidt_f64: sketch dimension k=%lld > 16384", (long long)k);
    int threads = (int)(((k + 7) / 8 + 31) / 32 * 32);
    threads = std::max(32, std::min(threads, k > 8192 ? 1024 : 512));
    const int64_t ept = (k + threads - 1) / threads;
    if (ept <= 4) gram_schmidt_kernel<4, 512><<<1, threads, 0, st>>>(a, r, k, lda, offset, R, flags, atol, rtol, thr);
    else if (ept <= 8) gram_schmidt_kernel<8, 512><<<1, threads, 0, st>>>(a, r, k, lda, offset, R, flags, atol, rtol, thr);
    else if (threads <= 512) gram_schmidt_kernel<16, 512><<<1, threads, 0, st>>>(a, r, k, lda, offset, R, flags, atol, rtol, thr);
    else gram_schmidt_kernel<16, 1024><<<1, threads, 0, st>>>(a, r, k, lda, offset, R, flags, atol, rtol, thr);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

// pairs_dev: (m_even - 1) rounds x (m_even / 2) pairs of int32 (round-robin schedule, -1 = bye),
// rot_dev: one int32 of scratch.  Both caller-owned (see reductor_ops.py).
extern "C" int rla_svd_jacobi_f64(double *a, int64_t k, int64_t m, int64_t lda, double *s, double *V,
                                  const int32_t *pairs_dev, int32_t *rot_dev, int max_sweeps, double tol,
                                  int *sweeps_done, void *stream) {
    RLA_REQUIRE(k >= 1 && m >= 1 && lda >= k && max_sweeps >= 1, "rla_svd_jacobi_f64: bad sizes");
    RLA_REQUIRE(a && s && pairs_dev && rot_dev, "rla_svd_jacobi_f64: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (V) {
        eye_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, st>>>(V, m);
        count_launch();
    }
    const int me = (int)(m + (m & 1));
    const int rounds = me - 1, npairs = me / 2;
    int sweep = 0;
    for (; sweep < max_sweeps && m > 1; ++sweep) {
        RLA_CUDA_CHECK(cudaMemsetAsync(rot_dev, 0, sizeof(int32_t), st));
        for (int rd = 0; rd < rounds; ++rd) {
            jacobi_round_kernel<<<npairs, 256, 0, st>>>(a, k, lda, V, m, pairs_dev + (size_t)rd * npairs * 2, npairs,
                                                        tol, rot_dev);
            count_launch();
        }
        int32_t rot = 0;
        RLA_CUDA_CHECK(cudaMemcpyAsync(&rot, rot_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        RLA_CUDA_CHECK(cudaStreamSynchronize(st));
        if (rot == 0) { ++sweep; break; }
    }
    row_norms_kernel<<<(unsigned)m, 256, 0, st>>>(a, k, lda, s);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    if (sweeps_done) *sweeps_done = sweep;
    return RLA_OK;
}

extern "C" int rla_residual_norm_f64(const double *S, int64_t Q, int64_t k, int64_t r, const double *th,
                                     const double *a, const double *b, int64_t P, const double *tr, double *out,
                                     void *stream) {
    RLA_REQUIRE(Q >= 0 && k >= 1 && r >= 0 && P >= 0, "rla_residual_norm_f64: bad sizes");
    RLA_REQUIRE(out && (Q == 0 || (S && th)) && (r == 0 || a) && (P == 0 || (b && tr)), "rla_residual_norm_f64: null pointer");
    residual_norm_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(S, Q, k, r, th, a, b, P, tr, out);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
