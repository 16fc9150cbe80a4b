// jacobi_cluster.cu -- one-sided Jacobi SVD of the small (m, k) sketch factor inside ONE
// thread-block cluster: the whole matrix lives in the distributed shared memory of C CTAs for
// the whole iteration and never touches L2 / HBM between the first load and the final store.
//
// The "thin QR / SVD of the k x m sketch" of BASELINE configs[4] is latency-bound:
// what costs is the number and the price of the sequential synchronisation points.  The
// grid-synchronised kernel of factor.cu hands 8-row blocks from CTA to CTA through global
// memory (flag wait + cp.async load + store + fence: ~10 us per block round); here a block
// round ends with a push over DSMEM and one cluster barrier (~2 us), and a rotation step reads
// and writes only ONE of the two rows in shared memory (the other stays in registers).
// Measured (256 x 256 triangular factor of configs[4], 11 sweeps): 4.3 ms -> 2.1 ms; the phase
// counters the kernel returns put 64 % of it in the rotation steps (843 cycles each: a chain of
// ~28 dependent FP64 operations and five shuffle levels -- carrying the norms along instead of
// recomputing them, or cosine / sine from two reciprocal square roots instead of sqrt + division,
// changed nothing: the step is bound by latency, not by the instructions a warp issues), 19 % in
// the end-of-round barrier, 5 % in the push.
// Inside a rotation step (cycle counters in one warp, same run): 21 % the shared-memory loads and
// the three dot products, 29 % the three 5-level warp reductions, 37 % cosine and sine (sqrt, division,
// rsqrt: ~400 cycles per rotation), 13 % the rotation itself; the CTA barrier of a step costs nothing
// (the warps run in lockstep).
//
// Layout: rows are [sketch part (k) | accumulated rotations (m, optional) | zero pad] of pitch
// P = 64 NL doubles, grouped in 2C blocks of B rows.  CTA i holds the pair (top_i, bot_i) of the
// circle-method tournament (2C - 1 block rounds per sweep):
//     top block: one row per warp IN REGISTERS (NL double2 per lane) during the round, parked in
//                the `stage` buffer between rounds;
//     bot block: shared memory, double buffered (the next holder's copy is written remotely
//                while the current one may still be read).
// Step t of a round: warp w rotates (top row w, bot row (w + t) mod B); one CTA barrier per step.
// End of a round: every CTA pushes its top rows to the right neighbour's stage and its bot rows
// to the left neighbour's spare bot buffer (st.shared::cluster through generic pointers; the
// ends of the chain turn round: top_{C-1} -> bot_{C-1}, bot_0 -> top_1, top_0 stays), then one
// cluster barrier.  A second, split barrier (arrive after the stage was read, wait before the
// push) keeps a fast CTA from overwriting a stage its neighbour has not read yet; it costs
// nothing, its two halves bracket the rotation steps.  Pairs inside a block are rotated once
// per sweep in shared memory before round 0.  Convergence: every CTA publishes its rotation
// count into every CTA's shared memory before the last barrier of the sweep; all CTAs take the
// same decision without a global-memory round trip.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace rla {

__device__ __forceinline__ uint32_t jc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t jc_cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void jc_cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void jc_cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `p` (a shared-memory address of this CTA) in the CTA of rank `rank`
template <typename T>
__device__ __forceinline__ T *jc_map(T *p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<uint64_t>(p)), "r"(rank));
    return reinterpret_cast<T *>(out);
}

__device__ __forceinline__ double jc_warp_sum(double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// cosine / sine of the rotation that makes rows x, y orthogonal: a = <x,x>, b = <y,y>, g = <x,y>
// (the formula of factor.cu's jacobi_rotate: one square root, one division, one rsqrt)
__device__ __forceinline__ void jc_angle(double a, double b, double g, double &c, double &s) {
    const double d = b - a, h = 2.0 * g;
    const double r = sqrt(fma(d, d, h * h));
    double t = fabs(h) / (fabs(d) + r);
    if ((d < 0.0) != (h < 0.0)) t = -t;
    c = rsqrt(fma(t, t, 1.0));
    s = c * t;
}

// x in registers (NL double2 per lane: columns 64 j + 2 lane, +1), y in shared memory.
template <int NL>
__device__ __forceinline__ int jc_rotate_reg(double2 (&x)[NL], double *yrow, int kcols, int lane, double tol) {
    double2 y[NL];
    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        y[j] = *reinterpret_cast<const double2 *>(yrow + 64 * j + 2 * lane);
        if (64 * j + 2 * lane < kcols) {
            a = fma(x[j].x, x[j].x, a); b = fma(y[j].x, y[j].x, b); g = fma(x[j].x, y[j].x, g);
            a = fma(x[j].y, x[j].y, a); b = fma(y[j].y, y[j].y, b); g = fma(x[j].y, y[j].y, g);
        }
    }
    a = jc_warp_sum(a); b = jc_warp_sum(b); g = jc_warp_sum(g);
    if (g * g <= (tol * tol) * (a * b) || g == 0.0) return 0;
    double c, s;
    jc_angle(a, b, g, c, s);
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const double2 xo = x[j];
        x[j] = make_double2(c * xo.x - s * y[j].x, c * xo.y - s * y[j].y);
        *reinterpret_cast<double2 *>(yrow + 64 * j + 2 * lane) = make_double2(s * xo.x + c * y[j].x, s * xo.y + c * y[j].y);
    }
    return 1;
}

// both rows in shared memory (pairs inside a block, once per sweep)
template <int NL>
__device__ __forceinline__ int jc_rotate_smem(double *xrow, double *yrow, int kcols, int lane, double tol) {
    double2 x[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) x[j] = *reinterpret_cast<const double2 *>(xrow + 64 * j + 2 * lane);
    const int rot = jc_rotate_reg<NL>(x, yrow, kcols, lane, tol);
    if (rot) {
#pragma unroll
        for (int j = 0; j < NL; ++j) *reinterpret_cast<double2 *>(xrow + 64 * j + 2 * lane) = x[j];
    }
    return rot;
}

// block held at tournament position `pos` (0 .. 2C-1) after `R` block rounds: position 0 never
// moves, positions 1 .. 2C-1 advance by one per round (the circle method)
__device__ __forceinline__ int jc_block_at(int pos, int R, int nb) {
    if (pos == 0) return 0;
    int v = (pos - 1 - R) % (nb - 1);
    if (v < 0) v += nb - 1;
    return 1 + v;
}

constexpr int JC_MAX_CLUSTER = 16;

// info: [0] sweeps done, [1] converged, [2] 0 (no timeouts: the hardware co-schedules a cluster)
template <int NL, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
jacobi_cluster_kernel(double *A, int k, int64_t lda, double *V, int m, int B, double tol, int max_sweeps,
                      double *sval, int32_t *info) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_cnt[2][JC_MAX_CLUSTER];
    __shared__ int s_rot;
    constexpr int P = 64 * NL;
    const int C = (int)jc_cluster_size(), rank = (int)jc_cluster_rank(), nb = 2 * C;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int vcols = V ? m : 0;
    double *stage = sm;
    double *bots0 = sm + (size_t)B * P;                   // bot buffer `cur` = bots0 + cur * B * P
    const size_t bstride = (size_t)B * P;

    // ---- first load: block `blk` -> buffer (one row per warp; the V part starts as identity)
    auto load_row = [&](int blk, double *dst) {
        const int g = blk * B + w;
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int i = 64 * j + 2 * lane;
            double2 v = make_double2(0.0, 0.0);
            if (g < m) {
                if (i < k) v = *reinterpret_cast<const double2 *>(A + (int64_t)g * lda + i);
                else if (i < k + vcols) v = make_double2(i - k == g ? 1.0 : 0.0, i + 1 - k == g ? 1.0 : 0.0);
            }
            *reinterpret_cast<double2 *>(dst + (size_t)w * P + i) = v;
        }
    };
    load_row(rank, stage);
    load_row(nb - 1 - rank, bots0);
    if (tid == 0) s_rot = 0;
    __syncthreads();

    // destinations of the end-of-round pushes (fixed for the whole run, only the bot parity flips)
    const int top_to_bot = rank == C - 1;                 // my top block becomes my own next bot block
    const int bot_to_top = rank == 0;                     // my bot block becomes the top block of CTA 1
    double *top_dst_stage = (rank == 0 || top_to_bot) ? stage : jc_map(stage, (uint32_t)(rank + 1));
    double *bot_dst_stage = jc_map(stage, 1u);
    double *bot_dst_bots0 = jc_map(bots0, (uint32_t)(rank > 0 ? rank - 1 : 0));   // left neighbour's bot buffer 0

    int cur = 0, R = 0, sweep = 0, converged = 0;
    long long ph[5] = {0, 0, 0, 0, 0}, tck = clock64();   // phase cycles: load+norms, steps, wait A, push, barrier B
    auto tick = [&](int i) { const long long n = clock64(); ph[i] += n - tck; tck = n; };
    for (; sweep < max_sweeps; ++sweep) {
        int rot = 0;
        for (int r = 0; r < nb - 1; ++r, ++R) {
            double *bot = bots0 + cur * bstride;
            if (r == 0 && B > 1) {
                // pairs inside the two blocks: warps [0, B/2) on the top block, [B/2, B) on the bot block
                const int half = B >> 1, which = w / half, pi = w % half;
                double *base = which ? bot : stage;
                for (int t = 0; t < B - 1; ++t) {
                    const int u = pi == 0 ? 0 : ((pi - 1 + t) % (B - 1)) + 1;
                    const int v = ((B - 2 - pi + t) % (B - 1)) + 1;
                    const int p = min(u, v), q = max(u, v);
                    rot += jc_rotate_smem<NL>(base + (size_t)p * P, base + (size_t)q * P, k, lane, tol);
                    __syncthreads();
                }
            }
            // top row of this warp -> registers; tell the cluster that my stage may be overwritten
            double2 x[NL];
#pragma unroll
            for (int j = 0; j < NL; ++j) x[j] = *reinterpret_cast<const double2 *>(stage + (size_t)w * P + 64 * j + 2 * lane);
            jc_cluster_arrive();
            tick(0);
            int q = w;
            for (int t = 0; t < B; ++t) {
                rot += jc_rotate_reg<NL>(x, bot + (size_t)q * P, k, lane, tol);
                if (++q == B) q = 0;
                __syncthreads();
            }
            tick(1);
            jc_cluster_wait();                            // every CTA has read its stage
            tick(2);
            // ---- push: top rows (registers) and bot rows (shared memory) to their next holders
            {
                double *dst = top_to_bot ? bots0 + (cur ^ 1) * bstride : top_dst_stage;
#pragma unroll
                for (int j = 0; j < NL; ++j) *reinterpret_cast<double2 *>(dst + (size_t)w * P + 64 * j + 2 * lane) = x[j];
                double *bdst = bot_to_top ? bot_dst_stage : bot_dst_bots0 + (cur ^ 1) * bstride;
#pragma unroll
                for (int j = 0; j < NL; ++j)
                    *reinterpret_cast<double2 *>(bdst + (size_t)w * P + 64 * j + 2 * lane) =
                        *reinterpret_cast<const double2 *>(bot + (size_t)w * P + 64 * j + 2 * lane);
            }
            if (r == nb - 2) {
                // last round of the sweep: publish my rotation count to every CTA
                if (lane == 0 && rot) atomicAdd(&s_rot, rot);
                __syncthreads();
                if (tid < C) *jc_map(&s_cnt[sweep & 1][rank], (uint32_t)tid) = s_rot;
                __syncthreads();
                if (tid == 0) s_rot = 0;
            }
            tick(3);
            jc_cluster_arrive();
            jc_cluster_wait();                            // pushes (and counts) are visible
            tick(4);
            cur ^= 1;
        }
        int total = 0;
        for (int i = 0; i < C; ++i) total += s_cnt[sweep & 1][i];
        if (total == 0) { converged = 1; ++sweep; break; }
    }

    // ---- store: the top block sits in `stage`, the bot block in bots[cur]
    auto store_row = [&](int blk, const double *src) {
        const int g = blk * B + w;
        if (g >= m) return;
        double nrm = 0.0;
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int i = 64 * j + 2 * lane;
            const double2 v = *reinterpret_cast<const double2 *>(src + (size_t)w * P + i);
            if (i < k) {
                *reinterpret_cast<double2 *>(A + (int64_t)g * lda + i) = v;
                nrm = fma(v.x, v.x, nrm); nrm = fma(v.y, v.y, nrm);
            } else if (i < k + vcols) {
                *reinterpret_cast<double2 *>(V + (int64_t)g * m + (i - k)) = v;
            }
        }
        nrm = jc_warp_sum(nrm);
        if (lane == 0) sval[g] = sqrt(nrm);
    };
    store_row(jc_block_at(rank, R, nb), stage);
    store_row(jc_block_at(nb - 1 - rank, R, nb), bots0 + cur * bstride);
    if (rank == 0 && tid == 0) {
        info[0] = sweep; info[1] = converged; info[2] = 0;
        for (int i = 0; i < 5; ++i) info[3 + i] = (int32_t)(ph[i] >> 10);   // kilo-cycles per phase (CTA 0, thread 0)
    }
    // nobody may exit while a peer could still write into its shared memory: all pushes were
    // completed by the last barrier, so no further synchronisation is needed here
}

struct JcCfg { int C, B, NL, maxw; size_t smem; };

// shape -> (cluster size, rows per block); C == 0 when the cluster kernel does not apply
static JcCfg jc_config(int64_t k, int64_t m, int want_v) {
    JcCfg c = {0, 0, 0, 0, 0};
    if (k < 2 || (k & 1) || m < 4 || (want_v && (m & 1)) || m > 1024 || k > 4096) return c;
    if (const char *env = getenv("RLA_JACOBI_CLUSTER")) {
        if (atoi(env) == 0) return c;
    }
    const int64_t L = k + (want_v ? m : 0);
    const int nl = (int)((L + 63) / 64);
    int NL, maxw;
    if (nl <= 4) { NL = 4; maxw = 16; }
    else if (nl <= 8) { NL = 8; maxw = 16; }
    else if (nl <= 16) { NL = 16; maxw = 8; }
    else if (nl <= 20) { NL = 20; maxw = 8; }
    else return c;
    int forced = 0;
    if (const char *env = getenv("RLA_JACOBI_CLUSTER")) forced = atoi(env);     // development: force a cluster size
    // smallest cluster whose blocks have at most 8 rows (a step is one warp per row pair and the
    // warps of a CTA share four FP64 pipes: 256 x 256 takes 2.8 ms with 16 rows per block on 8 CTAs,
    // 2.1 ms with 8 rows on 16); a larger cluster means more block rounds (barriers) per sweep
    for (int pass = 0; pass < 2; ++pass) {
        for (int C : {2, 4, 8, 16}) {
            if (forced > 1 && C != forced) continue;
            int B = (int)((m + 2 * C - 1) / (2 * C));
            B += B & 1;
            if (B < 2) B = 2;
            const int wlim = pass == 0 ? std::min(8, maxw) : maxw;
            if (B > wlim) continue;
            const size_t smem = (size_t)3 * B * NL * 64 * sizeof(double);
            if (smem > 200 * 1024) continue;
            c.C = C; c.B = B; c.NL = NL; c.maxw = maxw; c.smem = smem;
            return c;
        }
    }
    return c;
}

template <int NL, int MAXW>
static int jc_launch(const JcCfg &c, double *a, int k, int64_t lda, double *V, int m, double tol, int max_sweeps,
                     double *s, int32_t *info, cudaStream_t st, bool probe_only) {
    auto kern = jacobi_cluster_kernel<NL, MAXW>;
    RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    if (c.C > 8) RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)c.C);
    cfg.blockDim = dim3((unsigned)(32 * c.B));
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)c.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (probe_only) {
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
        return nclusters >= 1 ? 1 : 0;
    }
    int B = c.B;
    RLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a, k, lda, V, m, B, tol, max_sweeps, s, info));
    count_launch();
    return RLA_OK;
}

static int jc_dispatch(const JcCfg &c, double *a, int k, int64_t lda, double *V, int m, double tol, int max_sweeps,
                       double *s, int32_t *info, cudaStream_t st, bool probe_only) {
    switch (c.NL) {
        case 4: return jc_launch<4, 16>(c, a, k, lda, V, m, tol, max_sweeps, s, info, st, probe_only);
        case 8: return jc_launch<8, 16>(c, a, k, lda, V, m, tol, max_sweeps, s, info, st, probe_only);
        case 16: return jc_launch<16, 8>(c, a, k, lda, V, m, tol, max_sweeps, s, info, st, probe_only);
        case 20: return jc_launch<20, 8>(c, a, k, lda, V, m, tol, max_sweeps, s, info, st, probe_only);
    }
    return probe_only ? 0 : fail(RLA_ERR_INVALID, "jacobi cluster: bad NL %d", c.NL);
}

}  // namespace rla

using namespace rla;

// Cluster size the cluster-resident Jacobi kernel would use for this shape; 0: not applicable
// (shape too large for the distributed shared memory of 16 CTAs, odd k, no device, or the
// device cannot co-schedule the cluster) -- use rla_svd_jacobi_block_f64 / rla_svd_jacobi_f64.
extern "C" int rla_svd_jacobi_cluster_size(int64_t k, int64_t m, int want_v) {
    const JcCfg c = jc_config(k, m, want_v);
    if (!c.C) return 0;
    int dev = 0, cl = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&cl, cudaDevAttrClusterLaunch, dev) != cudaSuccess || !cl) { cudaGetLastError(); return 0; }
    // cached per configuration: the occupancy query is a driver call
    static int cache[5][33][21] = {};
    int ci = c.C == 2 ? 1 : c.C == 4 ? 2 : c.C == 8 ? 3 : 4;
    int &slot = cache[ci][c.B][c.NL];
    if (slot == 0) slot = jc_dispatch(c, nullptr, 0, 0, nullptr, 0, 0.0, 0, nullptr, nullptr, nullptr, true) == 1 ? 1 : -1;
    return slot > 0 ? c.C : 0;
}

// Same data convention as rla_svd_jacobi_f64 (rows of a_dev (m, k) are rotated in place, s_dev
// gets the row norms, V_dev (m, m) or NULL the accumulated rotations); info_dev: 8 int32
// {sweeps done, converged, 0, kilo-cycles of CTA 0 in: stage load + norms, rotation steps, wait for the
// stage barrier, push, end-of-round barrier}.  One launch of one cluster, no host synchronisation.
extern "C" int rla_svd_jacobi_cluster_f64(double *a, int64_t k, int64_t m, int64_t lda, double *s, double *V,
                                          int32_t *info_dev, int max_sweeps, double tol, void *stream) {
    RLA_REQUIRE(a && s && info_dev, "rla_svd_jacobi_cluster_f64: null pointer");
    RLA_REQUIRE(lda >= k && max_sweeps >= 1 && ((uintptr_t)a & 15) == 0 && (lda & 1) == 0 &&
                (!V || ((uintptr_t)V & 15) == 0),
                "rla_svd_jacobi_cluster_f64: bad sizes / alignment");
    const JcCfg c = jc_config(k, m, V != nullptr);
    RLA_REQUIRE(c.C && rla_svd_jacobi_cluster_size(k, m, V != nullptr) == c.C,
                "rla_svd_jacobi_cluster_f64: shape not supported (rla_svd_jacobi_cluster_size returned 0)");
    return jc_dispatch(c, a, (int)k, lda, V, (int)m, tol, max_sweeps, s, info_dev, (cudaStream_t)stream, false);
}
