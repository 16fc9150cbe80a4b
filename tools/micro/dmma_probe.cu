// dmma_probe.cu -- developer probe: (1) checks the assumed fragment layout of
// mma.sync.m16n8k16.f64 and m16n8k4 / m8n8k4, (2) measures DMMA and DFMA peak issue rates.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
        "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
          "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// layout check: C(16x8) = A(16x16) * B(16x8), A row-major [16][16], B as [k][n]
__global__ void layout_kernel(const double *A, const double *B, double *C) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a[8], b[4], c[4] = {0, 0, 0, 0};
    for (int j = 0; j < 4; ++j) {
        a[2 * j] = A[g * 16 + t + 4 * j];
        a[2 * j + 1] = A[(g + 8) * 16 + t + 4 * j];
        b[j] = B[(t + 4 * j) * 8 + g];
    }
    dmma16816(c, a, b);
    C[g * 8 + 2 * t] = c[0];
    C[g * 8 + 2 * t + 1] = c[1];
    C[(g + 8) * 8 + 2 * t] = c[2];
    C[(g + 8) * 8 + 2 * t + 1] = c[3];
}

template <int NACC>
__global__ void dmma_rate_kernel(double *out, int iters) {
    double a[8], b[4], c[NACC][4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + i;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma16816(c[j], a, b);
    }
    double s = 0;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_rate_kernel(double *out, int iters) {
    double c[NACC], a = threadIdx.x * 1e-3, b = 1.0000001;
    for (int j = 0; j < NACC; ++j) c[j] = j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) c[j] = fma(c[j], b, a);
    }
    double s = 0;
    for (int j = 0; j < NACC; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    // ---- layout
    std::vector<double> A(256), B(128), C(128), R(128, 0.0);
    for (int i = 0; i < 256; ++i) A[i] = (rand() % 17) - 8;
    for (int i = 0; i < 128; ++i) B[i] = (rand() % 13) - 6;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) for (int k = 0; k < 16; ++k) R[i * 8 + j] += A[i * 16 + k] * B[k * 8 + j];
    double *dA, *dB, *dC;
    cudaMalloc(&dA, 256 * 8); cudaMalloc(&dB, 128 * 8); cudaMalloc(&dC, 128 * 8);
    cudaMemcpy(dA, A.data(), 256 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 128 * 8, cudaMemcpyHostToDevice);
    layout_kernel<<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(C.data(), dC, 128 * 8, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < 128; ++i) err = fmax(err, fabs(C[i] - R[i]));
    printf("layout m16n8k16 max abs err = %g (%s)\n", err, err == 0 ? "OK" : "MISMATCH");
    // ---- rates
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        dmma_rate_kernel<8><<<sms, warps * 32>>>(out, 100);
        cudaEventRecord(e0);
        dmma_rate_kernel<8><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * sms;
        printf("DMMA m16n8k16: %2d warps/SM  %.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", warps, flop / ms / 1e9,
               flop / 2 / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        dfma_rate_kernel<16><<<sms, warps * 32>>>(out, 100);
        cudaEventRecord(e0);
        dfma_rate_kernel<16><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 16 * 32.0 * iters * warps * sms;
        printf("DFMA        : %2d warps/SM  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, flop / ms / 1e9,
               flop / 2 / (ms * 1e-3) / sms / (clk * 1e3));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
