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
extern "C" const char *rla_last_error(void) { return rla::g_err; }
