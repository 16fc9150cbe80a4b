// factor.cu -- the small factorisations of the (m, k) sketch as ONE persistent, grid-synchronised
// kernel each (cooperative launch): Gram-Schmidt with pyMOR's semantics
// (mor/sketched_reductor.py:94) and a block one-sided Jacobi SVD (the "thin QR / SVD of the
// k x m sketch" of BASELINE configs[4]).  Both are latency-bound: what matters is the number and
// the cost of the sequential synchronisation points, not flops or bytes.
//
//  * gs_grid_kernel: RIGHT-looking modified Gram-Schmidt.  Rows are dealt round-robin to G CTAs and
//    live in registers.  Step i: the owner of row i normalises it and hands q_i over; every CTA
//    then projects q_i out of the rows it owns (all its dot products in one block reduction).  Row
//    i has then received exactly the projections j = 0..i-1, in that order, that pyMOR's
//    left-looking loop applies, so R, the removal test (norm <= rtol * initial) and the
//    re-iteration test (norm < threshold * old norm) are pyMOR's.  A re-iteration pass is done by
//    the whole grid at once: every CTA computes the coefficients against the final rows it owns
//    and a partial correction, the partials are summed in two levels in CTA order.  (pyMOR
//    re-iterates with sequential MGS; the second-pass coefficients are O(eps), so the two differ by
//    O(eps^2).)  Round 2: rows, decisions, partial corrections and coefficients all travel as
//    tagged 8-byte words ("LL" protocol: data and its own ready flag in one atomic word) -- no
//    fence, no flag, no atomic on any path; a reader puts all its loads in flight before it checks
//    any.  256 x 1024: 18 ms (one CTA) -> 1.6 ms (round 1) -> 1.0 ms.
//
//  * jacobi_block_kernel: rows are grouped in blocks of B; a CTA holds a PAIR of blocks (2B rows of
//    the sketch and of the accumulated rotations) in shared memory and rotates all B*B cross pairs
//    (and, once per sweep, the pairs inside each block) with one warp per pair and CTA-local
//    barriers only.  Block pairs follow a round-robin schedule with one grid barrier per block
//    round: (m/B - 1) barriers per sweep instead of (m - 1) kernel launches, convergence is tested
//    on the device (no host synchronisation), singular values come out of the same launch.
#include "common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace rla {

__device__ __forceinline__ double ldcg_f64(const double *p) { return __ldcg(p); }

__device__ __forceinline__ int ld_acquire_gpu_i32(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_i32(int32_t *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(int32_t *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Block-wide wait until *flag >= want (thread 0 polls with acquire loads; a timeout on the SM cycle
// counter turns a protocol error into an error code instead of a hung GPU).  Returns the value seen,
// or -1 on timeout.  All threads get the same result.
__device__ __forceinline__ int block_wait_ge(const int32_t *flag, int want, unsigned long long timeout_ns, int *s_slot) {
    if (threadIdx.x == 0) {
        int v = ld_acquire_gpu_i32(flag);
        if (v < want) {
            // timeout on the SM-local cycle counter (cheap); 2 cycles per ns is an upper bound of the clock
            const long long t0 = clock64(), limit = (long long)(2 * timeout_ns);
            while ((v = ld_acquire_gpu_i32(flag)) < want) {
                if (clock64() - t0 > limit) { v = -1; break; }
            }
        }
        *s_slot = v;
    }
    __syncthreads();
    const int v = *s_slot;
    __syncthreads();
    return v;
}

// arrive-and-wait barrier over the G CTAs of a cooperative launch on a monotone counter
__device__ __forceinline__ int grid_barrier(int32_t *ctr, int &target, int G, unsigned long long timeout_ns, int *s_slot) {
    __threadfence();
    __syncthreads();
    target += G;
    if (threadIdx.x == 0) red_release_gpu_add(ctr, 1);
    return block_wait_ge(ctr, target, timeout_ns, s_slot);
}

// deterministic block sums of N values at once: shuffle tree in the warp, warp partials in warp order
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double (*red)[32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
        for (int off = 16; off; off >>= 1) v[n] += __shfl_xor_sync(0xffffffffu, v[n], off);
    }
    __syncthreads();                                   // previous readers of `red` are done
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) red[n][warp] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[n][w];
        v[n] = s;
    }
}

template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// Warp sums of N values per lane at once (N a power of two <= 16): at every level half of the
// values go to the partner lane, so the count halves while the sums grow -- the same additions as N
// butterflies (same pairs at every level, bit-identical sums) with N - 1 + (5 - log2 N) shuffles
// instead of 5 N.  On return v[0] of lane l is the warp sum of the ORIGINAL v[l >> (5 - log2 N)].
template <int N>
__device__ __forceinline__ void warp_reduce_multi(double (&v)[N], int lane) {
    constexpr int LG = Log2<N>::v;
#pragma unroll
    for (int s = 0; s < LG; ++s) {
        const int M = N >> s, o = 16 >> s;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < M / 2; ++j) {
            const double send = upper ? v[j] : v[j + M / 2];
            const double keep = upper ? v[j + M / 2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
#pragma unroll
    for (int o = 16 >> LG; o; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

// block_sum_n for the 256-thread Gram-Schmidt CTAs: multi-value warp reduction, the eight warp
// partials read back with 16-byte loads and added in warp order (no loop, no dependent loads).
// red: [N][8] doubles, 16-byte aligned.  Same sums, bit for bit, as block_sum_n.
template <int N>
__device__ __forceinline__ void block_sum8(double (&v)[N], double (*red)[8]) {
    constexpr int SH = 5 - Log2<N>::v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_multi<N>(v, lane);
    __syncthreads();                                   // previous readers of `red` are done
    if ((lane & ((1 << SH) - 1)) == 0) red[lane >> SH][warp] = v[0];
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double2 a = *reinterpret_cast<const double2 *>(&red[n][0]), b = *reinterpret_cast<const double2 *>(&red[n][2]);
        const double2 c = *reinterpret_cast<const double2 *>(&red[n][4]), d = *reinterpret_cast<const double2 *>(&red[n][6]);
        v[n] = ((((((0.0 + a.x) + a.y) + b.x) + b.y) + c.x) + c.y) + d.x + d.y;
    }
}

constexpr int GS_MAX_REITER = 3;
enum { GS_REMOVED = 1, GS_FINAL = 2, GS_REITER = 3 };

// Publication of a finished row WITHOUT a fence or a flag (round 2; the "LL" protocol collective
// libraries use for small messages): every double travels as two 8-byte words {32 data bits, 32-bit
// sequence tag}, written with one 16-byte store per element into the ring slot of the event; the
// tag = (event + 1) * 4 + decision.  A reader polls the very words it needs (its own EPT elements)
// until both halves of each carry the tag of the event it waits for: data and "ready" arrive in
// the same 8-byte atomic word, so there is no flag to order against and no memory fence on the
// writer's critical path (the release store after the 8 KB row write was ~60 % of the 3.8 us the
// owner of a row spent between receiving q_{i-1} and handing over q_i).  Slots are zeroed by the
// host before the launch (tag 0 never matches) and a slot is re-used after GS_LL_SLOTS(G) events:
// a CTA publishes only after it has consumed every earlier event and owns one row in any G
// consecutive rows, so nobody lags more than (GS_MAX_REITER + 1) (G + 1) events.
__device__ __forceinline__ void ll_store_f64(unsigned long long *p, double v, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    const unsigned long long w0 = (b & 0xffffffffull) | t, w1 = (b >> 32) | t;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// Reading side, in three branch-free pieces so that a thread can have ALL its loads in flight
// before it looks at any of them (a validity branch between two loads serialises them: one L2
// round trip per word instead of one per batch).
__device__ __forceinline__ void ll_ld(const unsigned long long *p, unsigned long long &w0, unsigned long long &w1) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
// both words carry a tag of event `want` (= event + 1)
__device__ __forceinline__ bool ll_valid(unsigned long long w0, unsigned long long w1, uint32_t want) {
    const uint32_t t0 = (uint32_t)(w0 >> 32), t1 = (uint32_t)(w1 >> 32);
    return ((t0 >> 2) == want) & (t1 == t0);
}
__device__ __forceinline__ double ll_value(unsigned long long w0, unsigned long long w1) {
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ int ll_decision(unsigned long long w0) { return (int)((w0 >> 32) & 3ull); }

// Workspace: status int32 | ll: GS_LL_SLOTS(G) ring slots of 2 * EPT * 256 words (row hand-over), then G
// partial slots (+ 2 RPC coefficient words each) and G leader slots (re-iteration, see below).
template <int RPC, int EPT>
__global__ void __launch_bounds__(256, 1)
gs_grid_kernel(double *A, int64_t r, int64_t k, int64_t lda, int64_t offset, double *R,
               int32_t *flags, double atol, double rtol, double thr,
               unsigned long long *ll, int nslot, int32_t *status, unsigned long long timeout_ns) {
    // (no __restrict__: other CTAs write these buffers while the kernel runs)
    __shared__ __align__(16) double red[RPC][8];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    constexpr int64_t SLOT_WORDS = 2 * (int64_t)EPT * 256;
    constexpr int64_t PART_WORDS = SLOT_WORDS + 2 * RPC;   // a partial correction + the RPC coefficients behind it
    double x[RPC][EPT];
    double init[RPC], nrm[RPC];
    int state[RPC];                                    // 0 pending, 1 final (orthonormal), 2 removed / absent

    {
        double ss[RPC];
#pragma unroll
        for (int s = 0; s < RPC; ++s) {
            const int64_t l = (int64_t)s * G + cta;
            ss[s] = 0.0;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int64_t c = tid + (int64_t)e * nthr;
                x[s][e] = (l < r && c < k) ? A[l * lda + c] : 0.0;
                ss[s] = fma(x[s][e], x[s][e], ss[s]);
            }
            // column l of R belongs to this CTA alone: identity, then the projections as they arrive
            if (l < r)
                for (int64_t j = tid; j < r; j += nthr) R[j * r + l] = (j == l) ? 1.0 : 0.0;
        }
        block_sum8<RPC>(ss, red);
#pragma unroll
        for (int s = 0; s < RPC; ++s) {
            const int64_t l = (int64_t)s * G + cta;
            init[s] = nrm[s] = sqrt(ss[s]);
            state[s] = l >= r ? 2 : (l < offset ? 1 : (init[s] <= atol ? 2 : 0));
            if (l < r && tid == 0) flags[l] = (state[s] == 2) ? 1 : 0;
        }
        __syncthreads();
    }

    // Owner side of one decision event for row i (slot sl of this CTA): optionally projects q_prev
    // out of the row first (the look-ahead), judges it, publishes the row (and the decision inside
    // its tags) into ring slot e; returns the decision.  The row is COPIED out of its slot into
    // xs[] (RPC * EPT predicated moves) so that the body exists once, not once per slot: this kernel
    // is latency-bound and its code used to be 150 KB -- the owner's path, run once per G rows by
    // every CTA, was an instruction-cache miss from end to end (per-row time doubled with RPC).
    auto decide = [&](int i, int sl, int iter, int e, bool project, const double (&qp)[EPT]) -> int {
        int d = GS_REMOVED;
        unsigned long long *slot = ll + (int64_t)(e % nslot) * SLOT_WORDS;
        double xs[EPT], ini = 0.0, old = 0.0;
        int st = 2;
#pragma unroll
        for (int s = 0; s < RPC; ++s) {
            if (s == sl) {
#pragma unroll
                for (int q = 0; q < EPT; ++q) xs[q] = x[s][q];
                ini = init[s]; old = nrm[s]; st = state[s];
            }
        }
        if (project && st == 0) {
            double p1[1] = {0.0};
#pragma unroll
            for (int q = 0; q < EPT; ++q) p1[0] = fma(qp[q], xs[q], p1[0]);
            block_sum8<1>(p1, red);
#pragma unroll
            for (int q = 0; q < EPT; ++q) xs[q] = fma(-p1[0], qp[q], xs[q]);
            if (tid == 0) R[(int64_t)(i - 1) * r + i] = p1[0];      // first and only main-path term of R[i-1, i]
        }
        if (st != 2) {
            double ss[1] = {0.0};
#pragma unroll
            for (int q = 0; q < EPT; ++q) ss[0] = fma(xs[q], xs[q], ss[0]);
            block_sum8<1>(ss, red);
            const double norm = sqrt(ss[0]);
            if (norm <= rtol * ini) d = GS_REMOVED;
            else if (!(norm < thr * old) || iter >= GS_MAX_REITER) d = GS_FINAL;
            else d = GS_REITER;
            old = norm;
            if (d == GS_FINAL) {
                const double inv = 1.0 / norm;
#pragma unroll
                for (int q = 0; q < EPT; ++q) xs[q] *= inv;
                if (tid == 0) R[(int64_t)i * r + i] = norm;
                st = 1;
            }
            if (d == GS_REMOVED) {
                if (tid == 0) flags[i] = 1;
                st = 2;
            }
        }
        const uint32_t tag = (uint32_t)((e + 1) << 2) | (uint32_t)d;
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int c = tid + q * nthr;
            ll_store_f64(slot + 2 * c, xs[q], tag);                         // the hand-over: no fence, no flag
            if (d == GS_FINAL && c < k) A[(int64_t)i * lda + c] = xs[q];    // the result (nobody reads it back)
        }
#pragma unroll
        for (int s = 0; s < RPC; ++s) {
            if (s == sl) {
#pragma unroll
                for (int q = 0; q < EPT; ++q) x[s][q] = xs[q];
                nrm[s] = old; state[s] = st;
            }
        }
        return d;
    };

    const int r32 = (int)r, off32 = (int)(offset < r ? offset : r);
    int ev = 0;
    bool ahead = false;                                // decision of the current row already published (look-ahead)
    int ahead_dec = 0;
    int owner = 0, slot = 0;                           // i % G, i / G without a division per row
    for (int i = 0; i < r32; ++i) {
        const bool mine = cta == owner;
        int iter = 0;
        while (true) {
            int dec = GS_FINAL;
            double q[EPT];
            if (i < off32) {
                // given orthonormal row, untouched in global memory
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int c = tid + e * nthr;
                    q[e] = c < k ? ldcg_f64(A + (int64_t)i * lda + c) : 0.0;
                }
            } else if (mine) {
                dec = ahead ? ahead_dec : decide(i, slot, iter, ev, false, q);
                ahead = false;
#pragma unroll
                for (int s = 0; s < RPC; ++s) {
                    if (s == slot) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) q[e] = x[s][e];       // the owner has the row in registers
                    }
                }
                ++ev;
            } else {
                const unsigned long long *src = ll + (int64_t)(ev % nslot) * SLOT_WORDS + 2 * tid;
                const uint32_t want = (uint32_t)(ev + 1);
                const long long t0 = clock64(), limit = (long long)(2 * timeout_ns);
                // ONE thread per CTA spins; the others wait at the CTA barrier and then read -- and
                // validate -- their own words (G x 256 spinning threads on the same lines of L2 measured
                // 3-5 % slower, with or without a __nanosleep back-off)
                unsigned long long w0[EPT], w1[EPT];
                if (tid == 0) {
                    do {
                        ll_ld(src, w0[0], w1[0]);
                        if (clock64() - t0 > limit) { *status = 1; break; }
                    } while (!ll_valid(w0[0], w1[0], want));
                }
                __syncthreads();
                while (true) {
#pragma unroll
                    for (int e = 0; e < EPT; ++e) ll_ld(src + 2 * e * nthr, w0[e], w1[e]);
                    bool ok = true;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) ok &= ll_valid(w0[e], w1[e], want);
                    if (ok) break;
                    if (clock64() - t0 > limit) {
                        *status = 1;
                        return;
                    }
                }
#pragma unroll
                for (int e = 0; e < EPT; ++e) q[e] = ll_value(w0[e], w1[e]);
                dec = ll_decision(w0[0]);
                ++ev;
            }
            if (dec == GS_REMOVED) break;
            if (dec == GS_FINAL) {
                // Look-ahead: the owner of row i+1 completes that row first and publishes it before it
                // updates its other rows, so the chain q_i -> q_{i+1} never waits for trailing updates.
                const int i1 = i + 1;
                int o1 = owner + 1, sl1 = slot;
                if (o1 == G) { o1 = 0; ++sl1; }
                int skip = -1;
                if (i1 < r32 && i1 >= off32 && o1 == cta) {
                    ahead_dec = decide(i1, sl1, 0, ev, true, q);
                    ahead = true;
                    skip = sl1;
                }
                // project q_i out of the (other) pending rows this CTA owns
                double p[RPC];
                bool any = false;
#pragma unroll
                for (int s = 0; s < RPC; ++s) {
                    const int l = s * G + cta;
                    p[s] = 0.0;
                    if (l > i && state[s] == 0 && s != skip) {
                        any = true;
#pragma unroll
                        for (int e = 0; e < EPT; ++e) p[s] = fma(q[e], x[s][e], p[s]);
                    }
                }
                if (any) {                             // uniform over the CTA
                    block_sum8<RPC>(p, red);
#pragma unroll
                    for (int s = 0; s < RPC; ++s) {
                        const int l = s * G + cta;
                        if (l > i && state[s] == 0 && s != skip) {
#pragma unroll
                            for (int e = 0; e < EPT; ++e) x[s][e] = fma(-p[s], q[e], x[s][e]);
                            if (tid == 0) R[(int64_t)i * r + l] = p[s];    // plain store: R[i, l] was 0 (no load on the path)
                        }
                    }
                }
                break;
            }
            // dec == GS_REITER: re-iteration of row i (q holds its current, unnormalised content).  NOT
            // rare: pyMOR re-iterates whenever a pass shrinks the row below 0.9 of its norm, i.e. for
            // every row past i ~ 0.19 k of a random block (a third of the rows of BASELINE configs[4]).
            // Every CTA computes the coefficients against the final rows it owns and its partial
            // correction; partials travel with the fence-free tagged words of the row hand-over and are
            // summed in TWO levels, in CTA order: group leaders (every NG-th CTA) add the partials of
            // their NG members, the owner adds the leaders' sums -- two hand-over latencies and
            // 2 * NG loads per thread instead of a fence + an atomic + G serial loads (22 -> 8 kcycles).
            {
                const int e_this = ev - 1;
                const uint32_t ptag = ((uint32_t)(e_this + 1) << 2) | (uint32_t)GS_REITER, pwant = (uint32_t)(e_this + 1);
                const long long t0 = clock64(), limit = (long long)(2 * timeout_ns);
                unsigned long long *part = ll + (int64_t)nslot * SLOT_WORDS;            // G partial slots
                unsigned long long *lead = part + (int64_t)G * PART_WORDS;              // leader slots
                double c_[RPC], v[EPT];
#pragma unroll
                for (int e = 0; e < EPT; ++e) v[e] = 0.0;
#pragma unroll
                for (int s = 0; s < RPC; ++s) {
                    const int j = s * G + cta;
                    c_[s] = 0.0;
                    if (j < i && state[s] == 1) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) c_[s] = fma(q[e], x[s][e], c_[s]);
                    }
                }
                block_sum8<RPC>(c_, red);
#pragma unroll
                for (int s = 0; s < RPC; ++s) {
                    const int j = s * G + cta;
                    const bool use = j < i && state[s] == 1;
                    if (use) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) v[e] = fma(c_[s], x[s][e], v[e]);
                    }
                    // coefficient of final row j = s G + cta, behind the partial vector of this CTA
                    if (tid == 0) ll_store_f64(part + (int64_t)cta * PART_WORDS + SLOT_WORDS + 2 * s, use ? c_[s] : 0.0, ptag);
                }
                const int NG = G <= 4 ? G : (G <= 16 ? 4 : (G <= 64 ? 8 : 12));        // group size
                const int my_lead = cta - cta % NG;
                if (cta != my_lead) {
#pragma unroll
                    for (int e = 0; e < EPT; ++e) ll_store_f64(part + (int64_t)cta * PART_WORDS + 2 * (tid + e * nthr), v[e], ptag);
                } else {
                    // leader: own partial first, then the members' in CTA order
                    const int nmem = min(NG, G - cta) - 1;
                    constexpr int U = EPT <= 4 ? 4 : (EPT <= 8 ? 2 : 1);
                    for (int m0 = 0; m0 < nmem; m0 += U) {
                        unsigned long long w0[U][EPT], w1[U][EPT];
                        while (true) {
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int m = min(m0 + u, nmem - 1);          // a short last batch re-reads its last member
#pragma unroll
                                for (int e = 0; e < EPT; ++e)
                                    ll_ld(part + (int64_t)(cta + 1 + m) * PART_WORDS + 2 * (tid + e * nthr), w0[u][e], w1[u][e]);
                            }
                            bool ok = true;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) ok &= ll_valid(w0[u][e], w1[u][e], pwant);
                            }
                            if (ok) break;
                            if (clock64() - t0 > limit) { *status = 1; return; }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (m0 + u < nmem) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) v[e] += ll_value(w0[u][e], w1[u][e]);
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < EPT; ++e) ll_store_f64(lead + (int64_t)(cta / NG) * SLOT_WORDS + 2 * (tid + e * nthr), v[e], ptag);
                }
                if (mine) {
                    const int nlead = (G + NG - 1) / NG;
                    double tot[EPT];
#pragma unroll
                    for (int e = 0; e < EPT; ++e) tot[e] = 0.0;
                    constexpr int U = EPT <= 4 ? 4 : (EPT <= 8 ? 2 : 1);
                    for (int g0 = 0; g0 < nlead; g0 += U) {
                        unsigned long long w0[U][EPT], w1[U][EPT];
                        while (true) {
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int g = min(g0 + u, nlead - 1);
#pragma unroll
                                for (int e = 0; e < EPT; ++e)
                                    ll_ld(lead + (int64_t)g * SLOT_WORDS + 2 * (tid + e * nthr), w0[u][e], w1[u][e]);
                            }
                            bool ok = true;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) ok &= ll_valid(w0[u][e], w1[u][e], pwant);
                            }
                            if (ok) break;
                            if (clock64() - t0 > limit) { *status = 1; return; }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (g0 + u < nlead) {
#pragma unroll
                                for (int e = 0; e < EPT; ++e) tot[e] += ll_value(w0[u][e], w1[u][e]);
                            }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < RPC; ++s) {
                        if (s == slot) {
#pragma unroll
                            for (int e = 0; e < EPT; ++e) x[s][e] -= tot[e];
                        }
                    }
                    // coefficients of all final rows j < i: R[j, i] += c_j (c_j sits in slot (j % G), word (j / G))
                    for (int j = tid; j < i; j += nthr) {
                        const unsigned long long *src = part + (int64_t)(j % G) * PART_WORDS + SLOT_WORDS + 2 * (j / G);
                        unsigned long long c0, c1;
                        do {
                            ll_ld(src, c0, c1);
                            if (clock64() - t0 > limit) { *status = 1; return; }
                        } while (!ll_valid(c0, c1, pwant));
                        R[(int64_t)j * r + i] += ll_value(c0, c1);
                    }
                    __syncthreads();                   // the row is complete in every thread before it is judged again
                }
                ++iter;
            }
        }
        if (++owner == G) { owner = 0; ++slot; }
    }
}

// ------------------------------------------------------------------ block one-sided Jacobi
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// rotate rows p, q (shared memory) of the sketch part (length k) and of the V part (length m);
// returns 1 when a rotation was applied.  One warp.
__device__ __forceinline__ int jacobi_rotate(double *ap, double *aq, int64_t k, double *vp, double *vq, int64_t m,
                                             double tol) {
    const int lane = threadIdx.x & 31;
    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll 4
    for (int64_t i = 2 * lane; i < k; i += 64) {       // k is even, rows are 16-byte aligned
        const double2 x = *reinterpret_cast<const double2 *>(ap + i), y = *reinterpret_cast<const double2 *>(aq + i);
        a = fma(x.x, x.x, a); b = fma(y.x, y.x, b); g = fma(x.x, y.x, g);
        a = fma(x.y, x.y, a); b = fma(y.y, y.y, b); g = fma(x.y, y.y, g);
    }
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    if (g * g <= (tol * tol) * (a * b) || g == 0.0) return 0;
    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (b - a) / (2 g), written with one
    // square root and one division: t = sign(d h) |h| / (|d| + sqrt(d^2 + h^2)), d = b - a, h = 2 g
    const double d = b - a, h = 2.0 * g;
    const double r = sqrt(fma(d, d, h * h));
    double t = fabs(h) / (fabs(d) + r);
    if ((d < 0.0) != (h < 0.0)) t = -t;
    const double c = rsqrt(fma(t, t, 1.0)), s = c * t;
#pragma unroll 4
    for (int64_t i = 2 * lane; i < k; i += 64) {
        const double2 x = *reinterpret_cast<const double2 *>(ap + i), y = *reinterpret_cast<const double2 *>(aq + i);
        *reinterpret_cast<double2 *>(ap + i) = make_double2(c * x.x - s * y.x, c * x.y - s * y.y);
        *reinterpret_cast<double2 *>(aq + i) = make_double2(s * x.x + c * y.x, s * x.y + c * y.y);
    }
    if (vp) {                                          // m is even and the V part is 16-byte aligned
#pragma unroll 4
        for (int64_t i = 2 * lane; i < m; i += 64) {
            const double2 x = *reinterpret_cast<const double2 *>(vp + i), y = *reinterpret_cast<const double2 *>(vq + i);
            *reinterpret_cast<double2 *>(vp + i) = make_double2(c * x.x - s * y.x, c * x.y - s * y.y);
            *reinterpret_cast<double2 *>(vq + i) = make_double2(s * x.x + c * y.x, s * x.y + c * y.y);
        }
    }
    return 1;
}

// sched: rounds x npairs x 2 block indices over an even number of blocks; (b, -1) marks the bye of
// block b.  counters: max_sweeps int32, blkflag: one int32 per block, bar: one int32 -- all zero on
// entry.  info: [0] sweeps done, [1] converged, [2] timeout.
// Inside a sweep there is NO grid barrier: block b is read and written by exactly one CTA per block
// round, so its next holder only waits for blkflag[b] (rounds completed on b); one barrier per
// sweep collects the rotation count.
template <int B>
__global__ void __launch_bounds__(32 * B, 1)
jacobi_block_kernel(double *A, int64_t k, int64_t lda, double *V, int64_t m,
                    const int32_t *__restrict__ sched, int rounds, double tol, int max_sweeps,
                    int32_t *counters, double *sval, int32_t *info, int32_t *blkflag, int32_t *bar,
                    unsigned long long timeout_ns) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_rot, s_slot;
    const int npairs = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, nthr = blockDim.x;
    const int64_t ka = k, va = V ? m : 0;
    const int64_t pitch = (ka + va + 3) & ~int64_t(1); // even: every row stays 16-byte aligned
    double *rows = sm;                                 // 2B rows: [sketch part (k) | V part (m)]
    int bar_target = 0;

    if (V) {
        for (int64_t i = (int64_t)cta * nthr + tid; i < m * m; i += (int64_t)npairs * nthr) V[i] = ((i / m) == (i % m)) ? 1.0 : 0.0;
        if (grid_barrier(bar, bar_target, npairs, timeout_ns, &s_slot) < 0) {
            if (tid == 0) info[2] = 1;
            return;
        }
    }
    // global -> shared with cp.async (16-byte, L2-coherent .cg): every chunk of every row is in flight
    // at once instead of one L2 round trip per row
    auto load_block = [&](int blk, int rbase) {
        const int64_t ca = ka / 2, cv = va / 2;        // 16-byte chunks per row part (k even; m even when V)
        for (int rr = 0; rr < B; ++rr) {
            const int64_t g = (int64_t)blk * B + rr;
            double *dst = rows + (rbase + rr) * pitch;
            if (g < m) {
                for (int64_t c = tid; c < ca + cv; c += nthr) {
                    const double *src = c < ca ? A + g * lda + 2 * c : V + g * m + 2 * (c - ca);
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + 2 * c);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
                }
            }
        }
    };
    auto load_wait = [&]() {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    };
    auto store_block = [&](int blk, int rbase) {
        for (int rr = 0; rr < B; ++rr) {
            const int64_t g = (int64_t)blk * B + rr;
            const double *src = rows + (rbase + rr) * pitch;
            if (g < m) {
                for (int64_t i = 2 * tid; i < ka; i += 2 * nthr)
                    *reinterpret_cast<double2 *>(A + g * lda + i) = *reinterpret_cast<const double2 *>(src + i);
                for (int64_t i = 2 * tid; i < va; i += 2 * nthr)
                    *reinterpret_cast<double2 *>(V + g * m + i) = *reinterpret_cast<const double2 *>(src + ka + i);
            }
        }
    };
    // pairs inside block `blk` held at rows [rbase, rbase + B): round-robin over its B rows;
    // `w0` = first warp working on it, B/2 warps busy per inner round
    auto intra = [&](int blk, int rbase, int w0, int &rot) {
        for (int t = 0; t < B - 1; ++t) {
            const int pi = warp - w0;
            if (pi >= 0 && pi < B / 2) {
                const int u = pi == 0 ? 0 : ((pi - 1 + t) % (B - 1)) + 1;
                const int w = ((B - 2 - pi + t) % (B - 1)) + 1;
                const int p = min(u, w), q = max(u, w);
                if ((int64_t)blk * B + q < m) {
                    double *rp = rows + (rbase + p) * pitch, *rq = rows + (rbase + q) * pitch;
                    rot += jacobi_rotate(rp, rq, ka, V ? rp + ka : nullptr, V ? rq + ka : nullptr, va, tol);
                }
            }
            __syncthreads();
        }
    };

    int sweep = 0, converged = 0, ground = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int rd = 0; rd < rounds; ++rd, ++ground) {
            const int bx = sched[((int64_t)rd * npairs + cta) * 2], by = sched[((int64_t)rd * npairs + cta) * 2 + 1];
            if (tid == 0) s_rot = 0;
            if (bx < 0) continue;                      // (cannot happen with one dummy block at most)
            // wait until the previous holders of my blocks have written them back
            if (ground > 0) {
                if (tid == 0) s_slot = 0;
                __syncthreads();
                if (tid == 0 || (tid == 32 && by >= 0)) {            // the two flags are polled concurrently
                    const int32_t *f = blkflag + (tid == 0 ? bx : by);
                    const long long t0 = clock64(), limit = (long long)(2 * timeout_ns);
                    while (ld_acquire_gpu_i32(f) < ground) {
                        if (clock64() - t0 > limit) { s_slot = -1; break; }
                    }
                }
                __syncthreads();
                if (s_slot < 0) {
                    if (tid == 0) info[2] = 1;
                    return;
                }
            }
            int rot = 0;
            if (by >= 0) {
                const int b0 = min(bx, by), b1 = max(bx, by);
                load_block(b0, 0);
                load_block(b1, B);
                load_wait();
                if (rd == 0 && B > 1) {
                    // both blocks at once: warps [0, B/2) on b0, [B/2, B) on b1
                    for (int t = 0; t < B - 1; ++t) {
                        const int blk = warp / (B / 2), pi = warp % (B / 2);
                        const int u = pi == 0 ? 0 : ((pi - 1 + t) % (B - 1)) + 1;
                        const int w = ((B - 2 - pi + t) % (B - 1)) + 1;
                        const int p = min(u, w), q = max(u, w);
                        if ((int64_t)(blk ? b1 : b0) * B + q < m) {
                            double *rp = rows + (blk * B + p) * pitch, *rq = rows + (blk * B + q) * pitch;
                            rot += jacobi_rotate(rp, rq, ka, V ? rp + ka : nullptr, V ? rq + ka : nullptr, va, tol);
                        }
                        __syncthreads();
                    }
                }
                for (int t = 0; t < B; ++t) {
                    const int p = warp, q = (warp + t) % B;
                    if ((int64_t)b0 * B + p < m && (int64_t)b1 * B + q < m) {
                        double *rp = rows + p * pitch, *rq = rows + (B + q) * pitch;
                        rot += jacobi_rotate(rp, rq, ka, V ? rp + ka : nullptr, V ? rq + ka : nullptr, va, tol);
                    }
                    __syncthreads();
                }
                store_block(b0, 0);
                store_block(b1, B);
            } else if (rd == 0 && B > 1) {
                // the block that sits out round 0 still gets its inner pairs rotated once per sweep
                load_block(bx, 0);
                load_wait();
                intra(bx, 0, 0, rot);
                store_block(bx, 0);
            }
            if ((tid & 31) == 0 && rot) atomicAdd(&s_rot, rot);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                if (s_rot) atomicAdd(counters + sweep, s_rot);
                st_release_gpu_i32(blkflag + bx, ground + 1);
                if (by >= 0) st_release_gpu_i32(blkflag + by, ground + 1);
            }
            __syncthreads();
        }
        if (grid_barrier(bar, bar_target, npairs, timeout_ns, &s_slot) < 0) {
            if (tid == 0) info[2] = 1;
            return;
        }
        if (__ldcg(counters + sweep) == 0) { converged = 1; ++sweep; break; }
    }
    // singular values: row norms (one warp per row)
    const int nw = nthr >> 5;
    for (int64_t g = (int64_t)cta * nw + warp; g < m; g += (int64_t)npairs * nw) {
        double v = 0.0;
        for (int64_t i = tid & 31; i < k; i += 32) {
            const double x = __ldcg(A + g * lda + i);
            v = fma(x, x, v);
        }
        v = warp_sum(v);
        if ((tid & 31) == 0) sval[g] = sqrt(v);
    }
    if (cta == 0 && tid == 0) { info[0] = sweep; info[1] = converged; }
}

// ------------------------------------------------------------------ T = R^-1, R upper triangular
// (the reference's T = pinv(R) for the square, full-rank R of Gram-Schmidt, mor/sketched_reductor.py:95)
// One warp per column j of T: back substitution R t = e_j, the inner sum split over the lanes.
__global__ void __launch_bounds__(128)
trinv_upper_kernel(const double *__restrict__ R, int64_t r, int64_t ldr, double *__restrict__ T, int64_t ldt) {
    extern __shared__ double tsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * 4 + warp;
    if (j >= r) return;
    double *t = tsm + (int64_t)warp * r;
    for (int64_t i = j + 1 + lane; i < r; i += 32) T[i * ldt + j] = 0.0;
    if (lane == 0) t[j] = 1.0 / R[j * ldr + j];
    __syncwarp();
    for (int64_t i = j - 1; i >= 0; --i) {
        double acc = 0.0;
        for (int64_t l = i + 1 + lane; l <= j; l += 32) acc = fma(R[i * ldr + l], t[l], acc);
        acc = warp_sum(acc);
        if (lane == 0) t[i] = -acc / R[i * ldr + i];
        __syncwarp();
    }
    for (int64_t i = lane; i <= j; i += 32) T[i * ldt + j] = t[i];
}

// Row-oriented variant for r <= 1024: one warp per ROW i of T (T R = I, right-looking): lane
// `lane` carries the running right-hand side acc_j of the columns j = lane + 32 c in registers; step l
// finalises t_l = acc_l / R_ll (one multiplication by the prefetched reciprocal, one shuffle
// broadcast) and subtracts t_l R[l, j] from the columns j > l -- a row of R, read coalesced and
// prefetched one step ahead, no reduction.  The chain per step is mul -> shuffle -> fma instead of
// a 5-level warp reduction and a division: 256 x 256 in 0.03 ms instead of 0.26 ms.
template <int NC>
__global__ void __launch_bounds__(128)
trinv_rows_kernel(const double *__restrict__ R, int r, int64_t ldr, double *__restrict__ T, int64_t ldt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + warp;
    if (i >= r) return;
    double acc[NC], tv[NC], rn[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        acc[c] = (lane + 32 * c == i) ? 1.0 : 0.0;
        tv[c] = 0.0;
        rn[c] = (lane + 32 * c < r) ? R[(int64_t)i * ldr + lane + 32 * c] : 0.0;   // row l = i
    }
    double inv = 1.0 / R[(int64_t)i * ldr + i];
    const int ci = i >> 5;
#pragma unroll
    for (int c0 = 0; c0 < NC; ++c0) {
        if (c0 < ci || 32 * c0 >= r) continue;
        const int l0 = c0 == ci ? (i & 31) : 0;
        const int l1 = min(32, r - 32 * c0);
        for (int ll = l0; ll < l1; ++ll) {
            const int l = 32 * c0 + ll;
            // prefetch row l + 1 of R and the reciprocal of its diagonal entry
            double rnext[NC];
            const int ln = l + 1;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                rnext[c] = (c >= c0 && ln < r && lane + 32 * c < r) ? R[(int64_t)ln * ldr + lane + 32 * c] : 0.0;
            const double dnext = ln < r ? R[(int64_t)ln * ldr + ln] : 1.0;
            const double t = __shfl_sync(0xffffffffu, acc[c0] * inv, ll);
            if (lane == ll) tv[c0] = t;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c >= c0 && lane + 32 * c > l) acc[c] = fma(-t, rn[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) rn[c] = rnext[c];
            inv = 1.0 / dnext;
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int j = lane + 32 * c;
        if (j < r) T[(int64_t)i * ldt + j] = tv[c];       // zero left of the diagonal
    }
}

static int coop_ok() {
    int dev = 0, ok = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&ok, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) return 0;
    return ok;
}

struct GsCfg { int rpc, ept, grid; };
// shape -> kernel configuration, rpc == 0 when the grid kernel does not apply
static GsCfg gs_config(int64_t r, int64_t k) {
    GsCfg c = {0, 0, 0};
    if (r < 16 || k < 1) return c;                      // a handful of rows: the one-CTA kernel is as fast
    const int sms = sm_count();
    const int64_t ept = (k + 255) / 256;
    int rpc = 0, e = 0;
    if (ept <= 4) { rpc = 8; e = 4; }
    else if (ept <= 8) { rpc = 8; e = 8; }
    else if (ept <= 16) { rpc = (r + 1) / 2 <= sms ? 2 : 4; e = 16; }   // 16 elements per thread: two rows per CTA keep the rows in registers
    else return c;
    // fewer rows per CTA = fewer dot products per step and CTA (the step is latency-bound), as long
    // as the grid stays co-resident; RLA_GS_RPC overrides (development)
    if (e <= 8) {
        int want = r <= 128 ? 2 : 4;                   // measured: 256 x 1024 is fastest with 4 (re-iterations sum G partials)
        if (const char *env = getenv("RLA_GS_RPC")) want = atoi(env);
        for (int cand : {2, 4, 8})
            if (cand >= want && (r + cand - 1) / cand <= sms) { rpc = cand; break; }
    }
    const int64_t g = (r + rpc - 1) / rpc;
    if (g > sms) return c;
    c.rpc = rpc; c.ept = e; c.grid = (int)g;
    return c;
}

}  // namespace rla

using namespace rla;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ring slots of the row hand-over: nobody lags more than (GS_MAX_REITER + 1) (G + 1) events (see the kernel)
static int gs_ll_slots(int G) { return (GS_MAX_REITER + 1) * (G + 2); }

struct GsLayout { size_t status, ll, total; };
static GsLayout gs_layout(const GsCfg &c, int64_t r, int64_t k) {
    (void)r; (void)k;
    GsLayout L;
    const size_t slot_words = (size_t)2 * c.ept * 256, part_words = slot_words + (size_t)2 * c.rpc;
    L.status = 0;
    L.ll = 16;
    L.total = L.ll + ((size_t)gs_ll_slots(c.grid) * slot_words + (size_t)c.grid * part_words + (size_t)c.grid * slot_words) *
                         sizeof(unsigned long long);
    return L;
}

extern "C" size_t rla_gram_schmidt_workspace_bytes(int64_t r, int64_t k) {
    const GsCfg c = gs_config(r, k);
    if (!c.rpc) return 0;
    return gs_layout(c, r, k).total;
}

// byte offset of the int32 status word inside the workspace (0 = ok, 1 = a flag wait timed out);
// -1 when the grid kernel does not apply to this shape (the one-CTA kernel has no waits)
extern "C" int64_t rla_gram_schmidt_status_offset(int64_t r, int64_t k) {
    const GsCfg c = gs_config(r, k);
    if (!c.rpc || !coop_ok()) return -1;
    return (int64_t)gs_layout(c, r, k).status;
}

extern "C" int rla_gram_schmidt_ws_f64(double *a, int64_t r, int64_t k, int64_t lda, int64_t offset, double *R,
                                       int32_t *flags, double atol, double rtol, double thr, void *ws, size_t ws_bytes,
                                       void *stream) {
    const GsCfg c = gs_config(r, k);
    const size_t need = rla_gram_schmidt_workspace_bytes(r, k);
    if (!c.rpc || !ws || ws_bytes < need || !coop_ok() || ((uintptr_t)ws & 15))
        return rla_gram_schmidt_f64(a, r, k, lda, offset, R, flags, atol, rtol, thr, stream);
    RLA_REQUIRE(lda >= k && offset >= 0, "rla_gram_schmidt_ws_f64: bad sizes");
    RLA_REQUIRE(a && R && flags, "rla_gram_schmidt_ws_f64: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const GsLayout L = gs_layout(c, r, k);
    char *base = static_cast<char *>(ws);
    int32_t *status = reinterpret_cast<int32_t *>(base + L.status);
    unsigned long long *ll = reinterpret_cast<unsigned long long *>(base + L.ll);
    int nslot = gs_ll_slots(c.grid);
    // status and every tagged word start from zero (tag 0 matches no event)
    RLA_CUDA_CHECK(cudaMemsetAsync(base, 0, L.total, st));
    unsigned long long timeout_ns = 5000000000ull;
    void *args[] = {&a, &r, &k, &lda, &offset, &R, &flags, &atol, &rtol, &thr, &ll, &nslot, &status, &timeout_ns};
    const void *fn = nullptr;
    if (c.ept == 4) fn = c.rpc == 8 ? (const void *)gs_grid_kernel<8, 4> : c.rpc == 4 ? (const void *)gs_grid_kernel<4, 4>
                                                                                    : (const void *)gs_grid_kernel<2, 4>;
    else if (c.ept == 8) fn = c.rpc == 8 ? (const void *)gs_grid_kernel<8, 8> : c.rpc == 4 ? (const void *)gs_grid_kernel<4, 8>
                                                                                         : (const void *)gs_grid_kernel<2, 8>;
    else fn = c.rpc == 2 ? (const void *)gs_grid_kernel<2, 16> : (const void *)gs_grid_kernel<4, 16>;
    // cooperative launch: guarantees that all CTAs are co-resident (they wait on each other's flags)
    RLA_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(c.grid), dim3(256), args, 0, st));
    count_launch();
    return RLA_OK;
}

// rows per block the block-Jacobi kernel would use for this shape (0: not applicable, use
// rla_svd_jacobi_f64)
extern "C" int rla_svd_jacobi_block_rows(int64_t k, int64_t m, int want_v) {
    if (k < 2 || (k & 1) || m < 4 || (want_v && (m & 1)) || !coop_ok()) return 0;
    const int sms = sm_count();
    // 8 rows per block measured best with and without the accumulated rotations (256 x 256 factor:
    // 5.3 / 4.3 ms at B = 8, 7.0 / 5.5 at 16, 5.7 / 4.7 at 4)
    int bmax = 8;
    if (const char *env = getenv("RLA_JACOBI_B")) bmax = std::max(2, std::min(16, atoi(env)));   // development
    for (int B = bmax; B >= 2; B >>= 1) {
        const size_t smem = (size_t)2 * B * (size_t)((k + (want_v ? m : 0) + 3) & ~int64_t(1)) * sizeof(double);
        const int64_t nblk = (m + B - 1) / B;
        const int64_t npairs = (nblk + 1) / 2;
        if (smem <= 200 * 1024 && npairs <= sms && nblk >= 2) return B;
    }
    return 0;
}

// scratch_dev: (max_sweeps + ceil(m / B) + 8) int32 (sweep counters, block flags, barrier, info[3])
extern "C" size_t rla_svd_jacobi_block_scratch_ints(int64_t m, int B, int max_sweeps) {
    if (B < 1) return 0;
    return (size_t)max_sweeps + (size_t)((m + B - 1) / B) + 8;
}

extern "C" int rla_svd_jacobi_block_f64(double *a, int64_t k, int64_t m, int64_t lda, double *s, double *V,
                                        const int32_t *sched_dev, int B, int32_t *scratch_dev, int max_sweeps,
                                        double tol, void *stream) {
    RLA_REQUIRE(a && s && sched_dev && scratch_dev, "rla_svd_jacobi_block_f64: null pointer");
    RLA_REQUIRE(B == rla_svd_jacobi_block_rows(k, m, V != nullptr) && B >= 2,
                "rla_svd_jacobi_block_f64: B must be rla_svd_jacobi_block_rows(k, m, want_v)");
    RLA_REQUIRE(lda >= k && max_sweeps >= 1 && ((uintptr_t)a & 15) == 0 && (lda & 1) == 0,
                "rla_svd_jacobi_block_f64: bad sizes / alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblk = (m + B - 1) / B;
    const int nbe = (int)(nblk + (nblk & 1));
    int rounds = nbe - 1, npairs = nbe / 2;
    const size_t smem = (size_t)2 * B * (size_t)((k + (V ? m : 0) + 3) & ~int64_t(1)) * sizeof(double);
    const size_t ints = rla_svd_jacobi_block_scratch_ints(m, B, max_sweeps);
    RLA_CUDA_CHECK(cudaMemsetAsync(scratch_dev, 0, sizeof(int32_t) * ints, st));
    int32_t *info = scratch_dev, *bar = scratch_dev + 4, *counters = scratch_dev + 8, *blkflag = counters + max_sweeps;
    unsigned long long timeout_ns = 5000000000ull;
    void *args[] = {&a, &k, &lda, &V, &m, &sched_dev, &rounds, &tol, &max_sweeps, &counters, &s, &info, &blkflag, &bar,
                    &timeout_ns};
    const void *fn = B == 16 ? (const void *)jacobi_block_kernel<16>
                   : B == 8 ? (const void *)jacobi_block_kernel<8>
                   : B == 4 ? (const void *)jacobi_block_kernel<4> : (const void *)jacobi_block_kernel<2>;
    RLA_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RLA_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(npairs), dim3(32 * B), args, smem, st));
    count_launch();
    return RLA_OK;
}

extern "C" int rla_trinv_upper_f64(const double *R, int64_t r, int64_t ldr, double *T, int64_t ldt, void *stream) {
    RLA_REQUIRE(r >= 0 && ldr >= r && ldt >= r, "rla_trinv_upper_f64: bad sizes");
    if (r == 0) return RLA_OK;
    RLA_REQUIRE(R && T, "rla_trinv_upper_f64: null pointer");
    if (r <= 1024) {
        const unsigned grid = (unsigned)((r + 3) / 4);
        cudaStream_t st = (cudaStream_t)stream;
        if (r <= 128) trinv_rows_kernel<4><<<grid, 128, 0, st>>>(R, (int)r, ldr, T, ldt);
        else if (r <= 256) trinv_rows_kernel<8><<<grid, 128, 0, st>>>(R, (int)r, ldr, T, ldt);
        else if (r <= 512) trinv_rows_kernel<16><<<grid, 128, 0, st>>>(R, (int)r, ldr, T, ldt);
        else trinv_rows_kernel<32><<<grid, 128, 0, st>>>(R, (int)r, ldr, T, ldt);
        count_launch();
        RLA_CUDA_CHECK(cudaGetLastError());
        return RLA_OK;
    }
    RLA_REQUIRE(r <= 6144, "rla_trinv_upper_f64: r=%lld > 6144", (long long)r);
    const size_t smem = (size_t)4 * (size_t)r * sizeof(double);
    RLA_CUDA_CHECK(cudaFuncSetAttribute((const void *)trinv_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trinv_upper_kernel<<<(unsigned)((r + 3) / 4), 128, smem, (cudaStream_t)stream>>>(R, r, ldr, T, ldt);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
