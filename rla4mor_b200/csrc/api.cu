// api.cu -- version, error string, device query.
#include "common.cuh"
#include <string.h>

namespace rla {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}
}  // namespace rla

namespace rla {
static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
}  // namespace rla

extern "C" int rla_version(void) { return 100; }
extern "C" unsigned long long rla_launch_count(void) { return __atomic_load_n(&rla::g_launches, __ATOMIC_RELAXED); }
// kernels of this library replayed from a captured CUDA graph (the host-side launch calls do not run again)
extern "C" void rla_launch_count_add(long long n) { rla::count_launch((int)n); }
extern "C" const char *rla_last_error(void) { return rla::g_err; }

// Pitched copy between host and device (cudaMemcpy2DAsync): the column slabs of a
// row-major host block that streaming.py pipelines to the GPU.  direction 0: H2D, 1: D2H.
extern "C" int rla_copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                                size_t height, int direction, void *stream) {
    RLA_REQUIRE(dst && src, "rla_copy2d_async: null pointer");
    RLA_REQUIRE(direction == 0 || direction == 1, "rla_copy2d_async: direction must be 0 (H2D) or 1 (D2H)");
    if (width_bytes == 0 || height == 0) return RLA_OK;
    RLA_CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height,
                                     direction == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                     (cudaStream_t)stream));
    return RLA_OK;
}
