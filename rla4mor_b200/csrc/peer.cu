// peer.cu -- the exchange step of the row-sharded sketch (kernel K5)
// written over NVLink peer memory instead of a library collective.
//
// When the vector dimension n is split into G slabs (one process per GPU), every rank holds an
// (m, k) PARTIAL sketch and the sketch is their sum.  The partial is tiny (2 MiB at
// BASELINE configs[4]: m = 256, k = 1024), so the exchange is latency-bound: each rank's
// sketch kernel writes its partial straight into a buffer the peers have mapped (CUDA IPC,
// one cudaMalloc per rank), and ONE kernel per rank
//   1. publishes "my partial of epoch e is complete" with a system-scope release store into
//      every peer's flag row,
//   2. waits (acquire loads on its own flag row, bounded by a wall-clock timeout) until all
//      peers have published epoch e,
//   3. reads the G partials over NVLink (16-byte loads), applies the per-sample sign of the
//      SRHT slab factorisation H_{2^d} = H_G (x) H_{2^d/G} -- (-1)^popcount(high_i & g) --
//      and sums them IN RANK ORDER, so every rank obtains the bit-identical sketch.
// Partials are double buffered by epoch parity: a rank can only start epoch e+2 after all
// peers have published e+1, i.e. after they finished reading epoch e, so no second barrier
// is needed.  No NCCL call, no extra sign-multiply kernel, no staging copy.
#include "common.cuh"
#include <string.h>

namespace rla {

constexpr int PEER_MAX_WORLD = 16;

struct PeerArgs {
    const double *part[PEER_MAX_WORLD];       // partial sketch of every rank (this epoch's parity)
    unsigned long long *flags[PEER_MAX_WORLD];// flag row of every rank: flags[p][r] = last epoch rank r published to p
    const int32_t *high;                      // (k) high index bits s_i >> log2(slab), or null (no signs)
    double *out;
    int *status;                              // set to 1 when the wait timed out
    unsigned long long epoch;
    unsigned long long timeout_ns;
    int64_t m, k, ldp, ldo;
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// peer data: never served from a stale L1 line
__device__ __forceinline__ double2 ld_peer_f64x2(const double *p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerArgs a) {
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    // 1. publish (the partial was written by the preceding kernel on this stream).  Every CTA
    //    publishes the same value, so progress never depends on one particular CTA being scheduled.
    if (threadIdx.x < a.world) {
        __threadfence_system();
        st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
    }
    // 2. wait for every peer's partial of this epoch
    if (threadIdx.x < a.world) {
        const unsigned long long *mine = a.flags[a.rank] + threadIdx.x;
        // timeout on the SM-local cycle counter (2 cycles per ns bounds the clock from above)
        const long long t0 = clock64(), limit = (long long)(2 * a.timeout_ns);
        while (ld_acquire_sys(mine) < a.epoch) {
            if (clock64() - t0 > limit) { s_fail = 1; break; }
        }
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) *a.status = 1;
        return;
    }
    // 3. rank-ordered signed sum over NVLink
    const bool vec = ((a.k | a.ldp | a.ldo) & 1) == 0;
    const int64_t kw = vec ? a.k / 2 : a.k;
    const int64_t total = a.m * kw;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / kw, cw = e - row * kw;
        if (vec) {
            const int64_t col = 2 * cw;
            int2 hi = make_int2(0, 0);
            if (a.high) hi = *reinterpret_cast<const int2 *>(a.high + col);
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int p = 0; p < a.world; ++p) {
                double2 v = ld_peer_f64x2(a.part[p] + row * a.ldp + col);
                v.x = xor_sign(v.x, (uint32_t)__popc(hi.x & p) << 31);
                v.y = xor_sign(v.y, (uint32_t)__popc(hi.y & p) << 31);
                acc.x += v.x;
                acc.y += v.y;
            }
            *reinterpret_cast<double2 *>(a.out + row * a.ldo + col) = acc;
        } else {
            const int hi = a.high ? a.high[cw] : 0;
            double acc = 0.0;
            for (int p = 0; p < a.world; ++p)
                acc += xor_sign(ld_peer_f64(a.part[p] + row * a.ldp + cw), (uint32_t)__popc(hi & p) << 31);
            a.out[row * a.ldo + cw] = acc;
        }
    }
}

}  // namespace rla

using namespace rla;

extern "C" int rla_peer_buffer_create(size_t bytes, void **dev_ptr, unsigned char *handle64) {
    RLA_REQUIRE(dev_ptr && handle64 && bytes > 0, "rla_peer_buffer_create: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    RLA_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(RLA_ERR_CUDA, "rla_peer_buffer_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return RLA_OK;
}

extern "C" int rla_peer_buffer_open(const unsigned char *handle64, void **dev_ptr) {
    RLA_REQUIRE(dev_ptr && handle64, "rla_peer_buffer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RLA_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RLA_OK;
}

extern "C" int rla_peer_buffer_close(void *dev_ptr) {
    if (dev_ptr) RLA_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return RLA_OK;
}

extern "C" int rla_peer_buffer_destroy(void *dev_ptr) {
    if (dev_ptr) RLA_CUDA_CHECK(cudaFree(dev_ptr));
    return RLA_OK;
}

extern "C" int rla_peer_allreduce_f64(const void *const *part_ptrs, void *const *flag_ptrs, int world, int rank,
                                      uint64_t epoch, int64_t m, int64_t k, int64_t ldp,
                                      const int32_t *high_dev, double *out_dev, int64_t ldo,
                                      int *status_dev, double timeout_s, void *stream) {
    RLA_REQUIRE(part_ptrs && flag_ptrs && out_dev && status_dev, "rla_peer_allreduce_f64: null pointer");
    RLA_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD, "rla_peer_allreduce_f64: world size must be in [1, %d]", PEER_MAX_WORLD);
    RLA_REQUIRE(rank >= 0 && rank < world, "rla_peer_allreduce_f64: rank out of range");
    RLA_REQUIRE(m >= 0 && k >= 0 && ldp >= k && ldo >= k, "rla_peer_allreduce_f64: bad shape");
    RLA_REQUIRE(epoch > 0, "rla_peer_allreduce_f64: epochs start at 1 (the flag rows are zero-initialised)");
    PeerArgs a;
    memset(&a, 0, sizeof a);
    for (int p = 0; p < world; ++p) {
        RLA_REQUIRE(part_ptrs[p] && flag_ptrs[p], "rla_peer_allreduce_f64: null peer pointer");
        RLA_REQUIRE(((uintptr_t)part_ptrs[p] & 15) == 0 && ((uintptr_t)flag_ptrs[p] & 7) == 0,
                    "rla_peer_allreduce_f64: partials must be 16-byte aligned");
        a.part[p] = (const double *)part_ptrs[p];
        a.flags[p] = (unsigned long long *)flag_ptrs[p];
    }
    RLA_REQUIRE(((uintptr_t)out_dev & 15) == 0 && (!high_dev || ((uintptr_t)high_dev & 7) == 0),
                "rla_peer_allreduce_f64: out / high must be 16 / 8-byte aligned");
    a.high = high_dev; a.out = out_dev; a.status = status_dev;
    a.epoch = epoch;
    a.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
    a.m = m; a.k = k; a.ldp = ldp; a.ldo = ldo; a.world = world; a.rank = rank;
    const int64_t work = m * ((k + 1) / 2);
    int64_t blocks = (work + 255) / 256;
    // grid-stride loop; no more CTAs than the GPU holds at once
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    peer_allreduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
