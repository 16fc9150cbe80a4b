// gemm_probe.cu -- developer probe (round 2): where does the DMMA pipe lose time in the
// sketch GEMM's consumer loop?  Each variant runs the 64 x 32 warp tile of gemm.cu
// (8 x 4 DMMA.8x8x4 per k4 step) with one more ingredient of the real kernel:
//   1  fragments in registers, loaded once (pure issue pattern: 8 distinct A, 4 distinct B)
//   2  + the 12 LDS.128 per 64 DMMAs of the double-buffered fragment loads
//   3  + four more warps generating Theta (Philox + Box-Muller) into shared memory
//   6  like 1, A.x paired with B.y (different register banks for A and B)
//   7  like 1, 16 accumulators reused (the issue-peak benchmark's pattern) but distinct A/B
// Timing only; results are meaningless.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rla4mor_b200/csrc/rng.cuh"

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double2 &v) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
}

constexpr int STAGE_BYTES = 256 * 128;   // 128 rows of A + 128 rows of B, 128 bytes each
constexpr int NSTAGE = 4;

template <int V, int NCW, int PV = 0>
__global__ void __launch_bounds__(NCW * 32 + (V == 3 ? 128 : 0), 1) probe_kernel(double *out, int iters, uint64_t seed) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < NSTAGE * STAGE_BYTES / 8; i += blockDim.x) reinterpret_cast<double *>(smem)[i] = 1e-3 * (i & 1023);
    __syncthreads();
    if (V == 3 && warp >= NCW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int p = tid - NCW * 32;
        for (int it = 0; it < iters / 2; ++it) {
            unsigned char *st = smem + (it % NSTAGE) * STAGE_BYTES;
            unsigned char *rowp = st + 128 * 128 + p * 128;
            const int sw = p & 7;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double v[4];
                if (PV == 0) rla::theta4<0>(seed, (uint32_t)p, (uint64_t)it * 4 + c, v);
                else if (PV == 1) {   // Philox only
                    const rla::PhiloxOut q = rla::philox4x32_10((uint32_t)(it * 4 + c), p, 0u, 0u, (uint32_t)seed, 7u);
                    v[0] = __hiloint2double(q.x, q.y); v[1] = __hiloint2double(q.y, q.z);
                    v[2] = __hiloint2double(q.z, q.w); v[3] = __hiloint2double(q.w, q.x);
                } else if (PV == 2) { // Box-Muller + F2F only
                    float n0, n1, n2, n3;
                    rla::box_muller((uint32_t)(it * 4 + c) * 2654435761u, p * 40503u + it, n0, n1);
                    rla::box_muller((uint32_t)(it * 4 + c) * 2246822519u, p * 50503u + it, n2, n3);
                    v[0] = n0; v[1] = n1; v[2] = n2; v[3] = n3;
                } else {              // full, widening by integer arithmetic instead of F2F
                    const rla::PhiloxOut q = rla::philox4x32_10((uint32_t)(it * 4 + c), p, 0u, 0u, (uint32_t)seed, 7u);
                    float n[4];
                    rla::box_muller(q.x, q.y, n[0], n[1]);
                    rla::box_muller(q.z, q.w, n[2], n[3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t f = __float_as_uint(n[j]);
                        const uint32_t mag = f & 0x7fffffffu;
                        const uint32_t hi = (f & 0x80000000u) | (mag ? (mag >> 3) + 0x38000000u : 0u);
                        v[j] = __hiloint2double((int)hi, (int)(f << 29));
                    }
                }
                *reinterpret_cast<double2 *>(rowp + (((2 * c) ^ sw) << 4)) = make_double2(v[0], v[1]);
                *reinterpret_cast<double2 *>(rowp + (((2 * c + 1) ^ sw) << 4)) = make_double2(v[2], v[3]);
            }
        }
        return;
    }
    if (V == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / 4) & 1, wn = warp % 4;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    double2 a8[2][8], b8[2][4];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t row_off = (uint32_t)g * 128u;
    const uint32_t off0 = row_off + ((uint32_t)((2 * t) ^ g) << 4);
    const uint32_t off1 = row_off + ((uint32_t)((2 * t + 1) ^ g) << 4);
    const uint32_t a_warp = (uint32_t)(wm * 64) * 128u, b_warp = 128u * 128u + (uint32_t)(wn * 32) * 128u;
#define LOAD_FRAGS(BUF, STAGE, OFF)                                                              \
    {                                                                                            \
        const uint32_t sa_ = smem_base + (STAGE) * STAGE_BYTES + a_warp + (OFF);                 \
        const uint32_t sb_ = smem_base + (STAGE) * STAGE_BYTES + b_warp + (OFF);                 \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) lds_f64x2(sa_ + i * 1024, a8[BUF][i]);     \
        _Pragma("unroll") for (int j = 0; j < 4; ++j) lds_f64x2(sb_ + j * 1024, b8[BUF][j]);     \
    }
#define MMA_HALF(BUF)                                                                            \
    {                                                                                            \
        _Pragma("unroll") for (int i = 0; i < 8; ++i)                                            \
            _Pragma("unroll") for (int j = 0; j < 4; ++j)                                        \
                dmma884(acc[i][j][0], acc[i][j][1], a8[BUF][i].x, V == 6 ? b8[BUF][j].y : b8[BUF][j].x); \
        _Pragma("unroll") for (int i = 0; i < 8; ++i)                                            \
            _Pragma("unroll") for (int j = 0; j < 4; ++j)                                        \
                dmma884(acc[i][j][0], acc[i][j][1], a8[BUF][i].y, V == 6 ? b8[BUF][j].x : b8[BUF][j].y); \
    }
    LOAD_FRAGS(0, 0, off0);
    LOAD_FRAGS(1, 0, off1);
    if (V == 7) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acc[i][j][0], acc[i][j][1], h ? a8[0][i + 4 * rep].y : a8[0][i + 4 * rep].x, h ? b8[0][j].y : b8[0][j].x);
                }
#pragma unroll
            for (int rep = 0; rep < 2; ++rep)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acc[i][j][0], acc[i][j][1], h ? a8[1][i + 4 * rep].y : a8[1][i + 4 * rep].x, h ? b8[1][j].y : b8[1][j].x);
                }
        }
    } else if (V == 1 || V == 6) {
        for (int it = 0; it < iters; ++it) {
            MMA_HALF(0);
            MMA_HALF(1);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            const int s = it % NSTAGE, sn = (it + 1) % NSTAGE;
            LOAD_FRAGS(1, s, off1);
            MMA_HALF(0);
            LOAD_FRAGS(0, sn, off0);
            MMA_HALF(1);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + tid] = s;
}

template <int V, int NCW, int PV = 0>
static void run(const char *name, double *out, int sms) {
    auto kern = probe_kernel<V, NCW, PV>;
    const int smem = NSTAGE * STAGE_BYTES;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int threads = NCW * 32 + (V == 3 ? 128 : 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    kern<<<sms, threads, smem>>>(out, 100, 1234);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<sms, threads, smem>>>(out, iters, 1234);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * 256 * 128.0 * iters * NCW * sms;
    printf("%-58s %2d warps  %8.3f ms  %6.2f TFLOP/s  (%s)\n", name, NCW, best, flop / best / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, (size_t)sms * 1024 * 8);
    run<7, 8>("7: 16 accumulators, distinct A/B", out, sms);
    run<1, 8>("1: 64x32 warp tile, fragments resident", out, sms);
    run<6, 8>("6: same, A.x with B.y (banks)", out, sms);
    run<2, 8>("2: + 12 LDS.128 per 64 DMMA", out, sms);
    run<3, 8>("3: + 4 producer warps generating Theta", out, sms);
    run<3, 8, 1>("3/1: producers: Philox only", out, sms);
    run<3, 8, 2>("3/2: producers: Box-Muller + F2F only", out, sms);
    run<3, 8, 3>("3/3: producers: full, integer widening", out, sms);
    run<1, 4>("1: 4 warps (1 per SMSP)", out, sms);
    run<2, 4>("2: 4 warps (1 per SMSP) + LDS", out, sms);
    return 0;
}
