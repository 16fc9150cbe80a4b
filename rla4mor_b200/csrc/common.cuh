// common.cuh -- shared helpers for librla_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/rla_b200.h"

namespace rla {

void set_error(const char *fmt, ...);

inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    set_error("%s", buf);
    return code;
}

#define RLA_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::rla::fail(RLA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                               cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define RLA_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) return ::rla::fail(RLA_ERR_INVALID, __VA_ARGS__);                    \
    } while (0)

inline int ceil_log2_i64(int64_t n) {   // d = int(ceil(log2(n))), n >= 1
    int d = 0;
    while ((int64_t(1) << d) < n) ++d;
    return d;
}

// number of SMs of the current device (cached)
int sm_count();
// bookkeeping for rla_launch_count(): every kernel launch of this library is counted
void count_launch(int n = 1);

// streaming 16-byte / 8-byte global loads that do not allocate in L1
__device__ __forceinline__ double2 ldg_stream_f64x2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// flip the sign of x when bit 31 of `m` is set (integer pipe, no FP64 op)
__device__ __forceinline__ double xor_sign(double x, uint32_t m) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    return __hiloint2double(hi ^ (int)(m & 0x80000000u), lo);
}
__device__ __forceinline__ float xor_sign(float x, uint32_t m) {
    return __int_as_float(__float_as_int(x) ^ (int)(m & 0x80000000u));
}

}  // namespace rla
