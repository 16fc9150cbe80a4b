// Are the FP64 vector pipe (DFMA) and the FP64 tensor pipe (DMMA) separate execution units
// on sm_100?  Even warps issue DMMA, odd warps DFMA; compare the combined rate with each alone.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: even warps DMMA, odd warps DFMA
__global__ void probe(double *out, int iters, int mode, unsigned long long *flops) {
    const int warp = threadIdx.x >> 5;
    const bool do_mma = mode == 0 || (mode == 2 && ((warp >> 2) & 1) == 0);   // both roles on all 4 sub-partitions
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001;
    double c[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) { c[j][0] = j; c[j][1] = -j; }
    if (do_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma884(c[j][0], c[j][1], a, b);
        }
    } else {
        for (int it = 0; it < 4 * iters; ++it) {            // same flops per warp as a DMMA warp
#pragma unroll
            for (int j = 0; j < 16; ++j) { c[j][0] = fma(c[j][0], b, a); c[j][1] = fma(c[j][1], b, a); }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, (size_t)sms * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, warps = 16;
    for (int mode = 0; mode < 3; ++mode) {
        probe<<<sms, warps * 32>>>(out, 100, mode, nullptr);
        cudaEventRecord(e0);
        probe<<<sms, warps * 32>>>(out, iters, mode, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // flops per warp per iteration: DMMA 16 * 2*8*8*4 = 8192; DFMA 32 lanes * 32 fma * 2 = 2048
        double mma_w = mode == 0 ? warps : (mode == 2 ? warps / 2 : 0), fma_w = mode == 1 ? warps : (mode == 2 ? warps / 2 : 0);
        double fl = (mma_w * 8192.0 + fma_w * 8192.0) * iters * sms;
        printf("mode %d (%s): %.3f ms  %.2f TFLOP/s  [DMMA part %.2f, DFMA part %.2f]\n", mode,
               mode == 0 ? "DMMA only" : mode == 1 ? "DFMA only" : "DMMA + DFMA", ms, fl / ms / 1e9,
               mma_w * 8192.0 * iters * sms / ms / 1e9, fma_w * 8192.0 * iters * sms / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
