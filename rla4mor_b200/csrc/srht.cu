// srht.cu -- fused SRHT: sign flip -> Walsh-Hadamard -> row subsample, one pass over HBM.
//
// Replaces srht(x, k, seed) (rla/srht.py:136-177) as called by
// SrhtEmbedding.apply (rla/embeddings.py:167-172).
//
// Math:  y[c,i] = scale * sum_j (-1)^popcount(s_i & j) r_j x[c,j].
// Split j = jh*4096 + jl and s_i = sh_i*4096 + sl_i (H_{2^d} = H_{2^a} (x) H_{4096}):
//     y[c,i] = scale * sum_jh (-1)^popcount(sh_i & jh) * T_jh[sl_i],
//     T_jh   = H_4096 (r .* x[c, jh*4096 : (jh+1)*4096])
// so only a 12-stage transform is done per 4096-element tile (in registers, one
// swizzled shared-memory transpose in the middle); the remaining d-12 stages
// collapse, for the k sampled outputs only, into one signed gather-accumulate per
// (sample, tile).  x is read from HBM exactly once and never written.
//
// Work decomposition: a CTA owns a power-of-two chunk of consecutive tiles of ONE vector (row
// of x) and keeps the sample accumulators in registers.  Tiles are visited in Gray-code
// order so the per-sample sign (-1)^popcount(sh & jh) is maintained by one conditional sign
// flip of the accumulator per step (integer XOR, no popcount, no FP64 op).  Per-chunk partial
// sketches go to a workspace and are summed in chunk order by a small finalize kernel
// (deterministic, no atomics).  Two kernels share this scheme:
//   srht_ws_kernel   (default) warp-specialised: four 64-thread transform groups that hold
//                    only tile data + four gather warps that hold the accumulators;
//   srht_main_kernel single-role 128-thread CTAs for problems of a few tiles per SM.
#include "tile.cuh"
#include <algorithm>
#include <new>
#include <vector>
#include <stdlib.h>
#include <string.h>

namespace rla {

template <typename T>
struct SrhtArgs {
    const T *x;
    int64_t ldx, n;
    int64_t m;
    const uint64_t *signw;   // [ntiles_valid][64]
    const uint32_t *desc;    // [NSLOT][128]   bits 0..11 tile_pos(sl), bits 12..31 sh
    T *ws;                   // [nchunks][m][NSLOT*128]
    int log2L;               // tiles per CTA = 2^log2L (>= 2)
    int aligned;             // rows are 16-byte aligned: full tiles take the fast load
    int64_t nchunks;         // chunks per row actually launched
    int64_t ntiles_valid;    // tiles that contain at least one element < n
};

// ---------------------------------------------------------------------------
// Single-role variant for small problems (a few tiles per SM): 128-thread CTAs, two per SM;
// every thread transforms AND accumulates, tile pairs are processed in lockstep.  Short
// prologue, no idle gather warps; register-bound (64 data + NSLOT accumulator doubles), so
// loads do not overlap the transform -- the warp-specialised kernel below is the fast path.
template <typename T, int NSLOT>
__global__ void __launch_bounds__(CTA, 2) srht_main_kernel(const SrhtArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int tid = threadIdx.x, grp = tid >> 6, tg = tid & 63;
    const int64_t row = blockIdx.x / a.nchunks;
    const int64_t chunk = blockIdx.x % a.nchunks;
    const int64_t c0 = chunk << a.log2L;
    const int npairs = 1 << (a.log2L - 1);
    const T *rowp = a.x + row * a.ldx;
    T *buf = sm + grp * TILE;
    // sample descriptors of this thread live in shared memory behind the two tile buffers
    uint32_t *sdesc = reinterpret_cast<uint32_t *>(sm + 2 * TILE) + tid;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) sdesc[s * CTA] = __ldg(a.desc + s * CTA + tid);

    T acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) acc[s] = T(0);
    auto is_fast = [&](int64_t jh) -> bool { return a.aligned && (jh + 1) * (int64_t)TILE <= a.n; };

    for (int u = 0; u < npairs; ++u) {
        // tile pair u in Gray-code order; valid tiles form a prefix
        const int64_t jhA = c0 + 2 * (int64_t)(u ^ (u >> 1));
        const int64_t jh = jhA + grp;
        // sign flip when moving from pair u-1 to pair u: bit (ctz(u)+1) of sh = desc bit 13+ctz(u)
        const int fl_shift = 31 - (13 + (u ? __ffs(u) - 1 : 0));
        if (jhA < a.ntiles_valid) {
            T v[64];
            if (jh < a.ntiles_valid) {
                const uint64_t sw = __ldg(a.signw + jh * GROUP + tg);
                if (is_fast(jh)) load_tile_fast(v, rowp + jh * TILE, tg);
                else load_tile_slow(v, rowp, a.n, jh, tg, buf, grp);
                flip_signs(v, sw);
            } else {
#pragma unroll
                for (int r = 0; r < 64; ++r) v[r] = T(0);
            }
            tile_fwht(v, buf, tg, grp);
            __syncthreads();
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                const uint32_t dsc = sdesc[s * CTA];
                const int off = dsc & (TILE - 1);
                const T vA = sm[off], vB = sm[TILE + off];
                const T w = vA + xor_sign(vB, dsc << 19);        // bit 12 = sh bit 0
                acc[s] = xor_sign(acc[s], dsc << fl_shift) + w;
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) acc[s] = xor_sign(acc[s], sdesc[s * CTA] << fl_shift);
        }
    }
    // accumulators are relative to the sign of the last A tile: undo it
    const int glast = (npairs - 1) ^ ((npairs - 1) >> 1);
    const uint32_t jlast = (uint32_t)(c0 + 2 * (int64_t)glast);
    T *wsp = a.ws + ((chunk * a.m + row) * (int64_t)(NSLOT * CTA));
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const uint32_t dsc = sdesc[s * CTA];
        const uint32_t par = __popc((dsc >> TILE_LOG2) & jlast) & 1u;
        wsp[s * CTA + tid] = xor_sign(acc[s], par << 31);
    }
}

// ---------------------------------------------------------------------------
// Warp-specialised variant (default): one CTA of 384 threads per SM.
//   threads   0..127  gather warps: own the NSLOT sample accumulators (and their descriptors)
//                     in registers, serve the four tile buffers in tile order
//   threads 128..383  four independent 64-thread transform groups; group q owns tile buffer q
//                     and the tiles t = q, q+4, q+8, ... of the CTA's Gray-code tile sequence
// A transform thread holds only its 64 tile elements (no accumulators), so the loads of its
// next tile are issued right behind the write-back stores of the current one and fly while
// the gather warps read it.  The four groups drift freely (each meets only the gather warps,
// through two named barriers per group: A_q "tile ready", B_q "buffer free"), so one group's
// FP64 butterflies overlap another's shared-memory passes and a third's load latency.
// setmaxnreg rebalances the launch allocation (384 x 168 registers): transform warps 184,
// gather warps 136 (32 accumulators + 32 sample descriptors in registers);
// 256*184 + 128*136 = 384*168.
constexpr int WS_THREADS = 3 * CTA;
constexpr int WS_GROUPS = 4;

__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <typename T, int NSLOT>
__global__ void __launch_bounds__(WS_THREADS, 1) srht_ws_kernel(const SrhtArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);                 // 4 tile buffers
    // warps 0..3 gather, warps 4..11 transform: the issue arbiter favours high warp ids, and
    // the transform warps are the critical path
    const int tid = threadIdx.x - CTA;
    const int64_t row = blockIdx.x / a.nchunks;
    const int64_t chunk = blockIdx.x % a.nchunks;
    const int64_t c0 = chunk << a.log2L;
    const int ntl = 1 << a.log2L;                            // tiles of this CTA
    // barrier ids: 1..4 group-internal (64), 5..8 A_q, 9..12 B_q (64 + 128 = 192 threads)

    if (tid >= 0) {
        // ------------------------------------------------------------ transform warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        const int q = tid >> 6, tg = tid & 63;
        const T *rowp = a.x + row * a.ldx;
        T *buf = sm + q * TILE;
        auto tile_of = [&](int t) -> int64_t { return c0 + (int64_t)(t ^ (t >> 1)); };
        auto is_fast = [&](int64_t jh) -> bool { return a.aligned && (jh + 1) * (int64_t)TILE <= a.n; };
        T v[64];
        uint64_t sw = 0;
        // Rotated loop with ONE load site per path: the loads of tile t + 4 are issued behind
        // the write-back of tile t; the first trip (t < 0) only loads.  (A second load site in
        // a prologue makes the compiler merge the two register sets through local memory,
        // which serialises the prefetch.)
        for (int t = q - WS_GROUPS; t < ntl; t += WS_GROUPS) {
            bool loaded = false;
            if (t >= 0) {
                const int64_t jh = tile_of(t);
                // B_q: the gather of this group's previous tile has finished reading the buffer.
                // A fast tile is already in registers and round 1 does not touch the buffer, so
                // the wait is deferred to just before the first store into it (measured: 24.3 ->
                // 24.0 ms on the C3 block; the gather then has a whole round of slack)
                bool waited = t < WS_GROUPS;
                if (!waited && (jh >= a.ntiles_valid || !is_fast(jh))) { bar_sync(9 + q, 192); waited = true; }
                if (jh < a.ntiles_valid) {
                    if (!is_fast(jh)) {
                        const int64_t j0 = jh * TILE;
#pragma unroll 4
                        for (int e = tg; e < TILE; e += GROUP) buf[e] = (j0 + e < a.n) ? Elem<T>::load1(rowp + j0 + e) : T(0);
                        bar_sync(1 + q, 64);
#pragma unroll
                        for (int h = 0; h < 32; ++h) { v[2 * h] = buf[128 * h + 2 * tg]; v[2 * h + 1] = buf[128 * h + 2 * tg + 1]; }
                        bar_sync(1 + q, 64);
                    }
                    constexpr int M = Elem<T>::MASK;
                    int tgo = tg;
                    asm volatile("" : "+r"(tgo));              // opaque: bounds what LICM can hoist
                    flip_and_butterflies64(v, sw);              // sign flip + tile bits 0, 7..11
                    // the M + 1 swizzled column offsets of this thread, shared by the three passes
                    const T *so[M + 1];
#pragma unroll
                    for (int c = 0; c <= M; ++c) so[c] = buf + (tgo ^ c);
                    if (!waited) bar_sync(9 + q, 192);
#pragma unroll
                    for (int r = 0; r < 64; ++r) const_cast<T *>(so[r & M])[r * 64] = v[r];
                    bar_sync(1 + q, 64);
                    // round 2 reads this thread's row with 16-byte loads in position order: chunk
                    // j ^ (tg & 7) goes to register chunk j (conflict free); see tile_pos_ws()
                    // for why the butterflies still produce the right outputs (up to signs that
                    // the sample descriptors carry)
                    constexpr int E = 16 / (int)sizeof(T);           // elements per 16-byte chunk
                    using Chunk = typename Elem<T>::Chunk;
                    Chunk *rowc = reinterpret_cast<Chunk *>(buf + tgo * 64);
                    int cx[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) cx[i] = i ^ (tgo & 7);
#pragma unroll
                    for (int j = 0; j < 64 / E; ++j) Elem<T>::unpack(rowc[(j & ~7) + cx[j & 7]], &v[E * j]);
                    butterflies64(v);                           // tile bits 1..6
                    // write-back to the same chunks; when the next tile of this group is a fast
                    // one, its loads are issued right behind the stores that free the registers
                    const int64_t jn2 = tile_of(t + WS_GROUPS);
                    if (t + WS_GROUPS < ntl && jn2 < a.ntiles_valid && is_fast(jn2)) {
                        const T *np_ = rowp + jn2 * TILE + 2 * tg;
#pragma unroll
                        for (int j = 0; j < 64 / E; ++j) {
                            rowc[(j & ~7) + cx[j & 7]] = Elem<T>::pack(&v[E * j]);
#pragma unroll
                            for (int i = 0; i < E / 2; ++i) {
                                const int h = (E * j) / 2 + i;
                                Elem<T>::load2(np_ + 128 * h, v[2 * h], v[2 * h + 1]);
                            }
                        }
                        loaded = true;
                    } else {
#pragma unroll
                        for (int j = 0; j < 64 / E; ++j) rowc[(j & ~7) + cx[j & 7]] = Elem<T>::pack(&v[E * j]);
                    }
                }
                // A_q: the transformed tile is in shared memory
                bar_arrive(5 + q, 192);
            }
            if (t + WS_GROUPS < ntl) {
                const int64_t jn = tile_of(t + WS_GROUPS);
                sw = (jn < a.ntiles_valid) ? __ldg(a.signw + jn * GROUP + tg) : 0ull;
                // every path redefines all of v (nothing of the old tile stays live): fast tiles
                // are loaded here (or were, interleaved with the write-back above); ragged /
                // unaligned tiles are staged later by the slow path
                if (!loaded) {
                    if (jn < a.ntiles_valid && is_fast(jn)) {
                        load_tile_fast(v, rowp + jn * TILE, tg);
                    } else {
#pragma unroll
                        for (int r = 0; r < 64; ++r) v[r] = T(0);
                    }
                }
            }
        }
        return;
    }

    // ------------------------------------------------------------------ gather warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 136;");
    const int gt = threadIdx.x;
    // sample descriptors and accumulators of this thread both live in registers
    uint32_t dsc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) dsc[s] = __ldg(a.desc + s * CTA + gt);
    T acc[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) acc[s] = T(0);
    for (int t = 0; t < ntl; ++t) {
        const int64_t jh = c0 + (int64_t)(t ^ (t >> 1));
        // moving from tile t-1 to t flips bit ctz(t) of the tile index = desc bit 12 + ctz(t)
        const int fl_shift = 31 - (12 + (t ? __ffs(t) - 1 : 0));
        const int q = t & (WS_GROUPS - 1);
        bar_sync(5 + q, 192);
        if (jh < a.ntiles_valid) {
            const T *tb = sm + q * TILE;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                acc[s] = xor_sign(acc[s], dsc[s] << fl_shift) + tb[dsc[s] & (TILE - 1)];
        } else {
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) acc[s] = xor_sign(acc[s], dsc[s] << fl_shift);
        }
        if (t + WS_GROUPS < ntl) bar_arrive(9 + q, 192);               // somebody will wait for it
    }
    // accumulators are relative to the sign of the last tile: undo it
    const uint32_t jlast = (uint32_t)(c0 + (int64_t)((ntl - 1) ^ ((ntl - 1) >> 1)));
    T *wsp = a.ws + ((chunk * a.m + row) * (int64_t)(NSLOT * CTA));
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        // bit 31 of the descriptor: sign twist of the round-2 output (tile_pos_ws)
        const uint32_t par = (__popc((dsc[s] >> TILE_LOG2) & jlast) ^ (dsc[s] >> 31)) & 1u;
        wsp[s * CTA + gt] = xor_sign(acc[s], par << 31);
    }
}

// y[row, i] = scale * sum_chunks ws[pass(i)][chunk][row][slot(i)]
template <typename T>
__global__ void srht_finalize_kernel(const T *__restrict__ ws, const int32_t *__restrict__ slotmap,
                                     T *__restrict__ y, int64_t ldy, int64_t k, int64_t m,
                                     int64_t nchunks, int nst, int64_t pass_stride, T scale) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.y;
    if (i >= k) return;
    const int32_t sm = slotmap[i];
    const int pass = sm / nst, slot = sm % nst;
    const T *p = ws + pass * pass_stride + row * nst + slot;
    T s = T(0);
    for (int64_t q = 0; q < nchunks; ++q) s += p[q * m * nst];
    y[row * ldy + i] = scale * s;
}

}  // namespace rla

using namespace rla;

// One assignment of the distinct samples to (pass, slot, thread) cells plus the map from
// sample index to cell.  A plan carries two: the smallest power-of-two slot count that fits
// (single-role kernel, registers are scarce there) and 32 slots (warp-specialised kernel:
// measured fastest for every k, half-empty slots included -- more slack per bank bucket
// means no overflow samples and no bank conflicts in the gather).
struct SrhtDescSet {
    int nslot = 0;
    size_t off_desc = 0, off_slot = 0;
};

struct rla_srht_plan {
    int64_t n = 0, k = 0;
    int d = 0;
    int elem_bytes = 8;
    int64_t ntiles = 0;        // padded tiles per row (>= 2)
    int64_t ntiles_valid = 0;
    int npass = 1;
    int64_t nuniq = 0;
    size_t off_sign = 0;
    SrhtDescSet sets[2];       // [0] minimal slot count, [1] 32 slots
    std::vector<unsigned char> image;
    const unsigned char *dev = nullptr;
};

static size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

// slot assignment: lane (tid % lanes) should equal the bank group of the gathered word
static void build_desc_set(rla_srht_plan *p, const SrhtDescSet &set, bool ws_layout,
                           const std::vector<int64_t> &uniq, const int64_t *idx, int64_t per_pass) {
    const int mask = p->elem_bytes == 8 ? 15 : 31;
    const int lanes = mask + 1;                // lanes per shared-memory wavefront
    const int nst = set.nslot * CTA;
    uint32_t *desc = reinterpret_cast<uint32_t *>(p->image.data() + set.off_desc);
    std::vector<int32_t> slot_of_uniq(uniq.size());
    const int threads_per_bucket = CTA / lanes;
    // descriptor of one distinct sample: bits 0..11 position in the tile buffer, bits 12..30 the
    // tile-index bits sh, bit 31 (warp-specialised layout only) a sign folded in at the end
    auto make_desc = [&](int64_t s) -> uint32_t {
        const int e = (int)(s & (TILE - 1));
        int neg = 0;
        const int pos = ws_layout ? tile_pos_ws(e, p->elem_bytes, &neg) : tile_pos(e, mask);
        return (uint32_t)pos | ((uint32_t)(s >> TILE_LOG2) << TILE_LOG2) | ((uint32_t)neg << 31);
    };
    for (int pass = 0; pass < p->npass; ++pass) {
        const int64_t u0 = pass * per_pass, u1 = std::min<int64_t>(p->nuniq, u0 + per_pass);
        std::vector<int> cnt(lanes, 0);
        std::vector<char> used(nst, 0);
        std::vector<int64_t> overflow;
        const int cap = set.nslot * threads_per_bucket;
        for (int64_t u = u0; u < u1; ++u) {
            const uint32_t dsc = make_desc(uniq[u]);
            const int b = (int)(dsc & (uint32_t)mask);
            if (cnt[b] < cap) {
                const int c = cnt[b]++;
                const int tid = b + lanes * (c % threads_per_bucket), s = c / threads_per_bucket;
                const int cell = s * CTA + tid;
                used[cell] = 1;
                desc[(size_t)pass * nst + cell] = dsc;
                slot_of_uniq[u] = pass * nst + cell;
            } else {
                overflow.push_back(u);
            }
        }
        int cell = 0;
        for (int64_t u : overflow) {           // bucket full: any free cell (costs a bank conflict)
            while (used[cell]) ++cell;
            used[cell] = 1;
            desc[(size_t)pass * nst + cell] = make_desc(uniq[u]);
            slot_of_uniq[u] = pass * nst + cell;
        }
    }
    int32_t *slotmap = reinterpret_cast<int32_t *>(p->image.data() + set.off_slot);
    for (int64_t i = 0; i < p->k; ++i) {
        const size_t u = std::lower_bound(uniq.begin(), uniq.end(), idx[i]) - uniq.begin();
        slotmap[i] = slot_of_uniq[u];
    }
}

extern "C" int rla_srht_plan_create(rla_srht_plan **out, const int8_t *signs, int64_t n,
                                    const int64_t *idx, int64_t k, int elem_bytes) {
    RLA_REQUIRE(out && signs && (idx || k == 0), "rla_srht_plan_create: null argument");
    RLA_REQUIRE(n >= 1 && k >= 0, "rla_srht_plan_create: need n >= 1, k >= 0 (n=%lld k=%lld)", (long long)n, (long long)k);
    RLA_REQUIRE(elem_bytes == 8 || elem_bytes == 4, "rla_srht_plan_create: elem_bytes must be 8 or 4");
    const int d = ceil_log2_i64(n);
    RLA_REQUIRE(d <= 31, "rla_srht_plan_create: n too large (d=%d > 31)", d);
    const int64_t np2 = int64_t(1) << d;
    for (int64_t i = 0; i < k; ++i)
        RLA_REQUIRE(idx[i] >= 0 && idx[i] < np2, "rla_srht_plan_create: idx[%lld]=%lld outside [0, 2^%d)",
                    (long long)i, (long long)idx[i], d);
    rla_srht_plan *p = new (std::nothrow) rla_srht_plan;
    RLA_REQUIRE(p, "out of host memory");
    p->n = n; p->k = k; p->d = d; p->elem_bytes = elem_bytes;
    const int dp = std::max(d, TILE_LOG2 + 1);
    p->ntiles = int64_t(1) << (dp - TILE_LOG2);
    p->ntiles_valid = (n + TILE - 1) / TILE;
    // distinct sample values
    std::vector<int64_t> uniq(idx, idx + k);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    p->nuniq = (int64_t)uniq.size();
    p->npass = (int)std::max<int64_t>(1, (p->nuniq + MAX_SLOTS - 1) / MAX_SLOTS);
    const int64_t per_pass = (p->nuniq + p->npass - 1) / p->npass;
    int nslot = 4;
    while ((int64_t)nslot * CTA < per_pass) nslot *= 2;
    p->sets[0].nslot = nslot;
    p->sets[1].nslot = MAX_NSLOT;
    const size_t slot_bytes = align16((size_t)std::max<int64_t>(k, 1) * 4);
    p->off_sign = 0;
    size_t off = align16(p->off_sign + (size_t)p->ntiles_valid * GROUP * 8);
    for (SrhtDescSet &set : p->sets) {
        set.off_desc = off;
        set.off_slot = align16(off + (size_t)p->npass * set.nslot * CTA * 4);
        off = set.off_slot + slot_bytes;
    }
    p->image.assign(off, 0);
    // packed sign bits: word [jh*64 + tg], bit rho=2h+l <-> element jh*4096 + 128h + 2tg + l
    uint64_t *sw = reinterpret_cast<uint64_t *>(p->image.data() + p->off_sign);
    for (int64_t j = 0; j < n; ++j) {
        if (signs[j] < 0) {
            const int64_t jh = j >> TILE_LOG2;
            const int e = (int)(j & (TILE - 1));
            const int h = e >> 7, tg = (e >> 1) & 63, l = e & 1;
            sw[jh * GROUP + tg] |= uint64_t(1) << (2 * h + l);
        }
    }
    build_desc_set(p, p->sets[0], false, uniq, idx, per_pass);
    build_desc_set(p, p->sets[1], true, uniq, idx, per_pass);
    *out = p;
    return RLA_OK;
}

extern "C" void rla_srht_plan_destroy(rla_srht_plan *p) { delete p; }
extern "C" size_t rla_srht_plan_device_bytes(const rla_srht_plan *p) { return p ? p->image.size() : 0; }
extern "C" int rla_srht_plan_passes(const rla_srht_plan *p) { return p ? p->npass : 0; }

extern "C" int rla_srht_plan_upload(rla_srht_plan *p, void *plan_dev, void *stream) {
    RLA_REQUIRE(p && plan_dev, "rla_srht_plan_upload: null argument");
    RLA_REQUIRE((reinterpret_cast<uintptr_t>(plan_dev) & 15) == 0, "rla_srht_plan_upload: buffer must be 16-byte aligned");
    RLA_CUDA_CHECK(cudaMemcpyAsync(plan_dev, p->image.data(), p->image.size(), cudaMemcpyHostToDevice,
                                   (cudaStream_t)stream));
    RLA_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));   // image is pageable host memory
    p->dev = static_cast<const unsigned char *>(plan_dev);
    return RLA_OK;
}

static int srht_variant_for(int64_t total_tiles);

// tiles per CTA (a power of two): as many as possible while the grid still has ~8 waves of
// CTAs, so the per-CTA prologue / drain and the partial-sketch write stay small next to the
// tile reads, and never fewer than 4 tiles when the row has them.
static int choose_log2L(const rla_srht_plan *p, int64_t m) {
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("RLA_SRHT_LOG2L");
        forced = e ? atoi(e) : -1;
    }
    int maxl = 0;
    while ((int64_t(1) << (maxl + 1)) <= p->ntiles) ++maxl;
    if (forced >= 1) return std::min(forced, maxl);
    const int64_t total = m * p->ntiles_valid;
    const int ctas_per_sm = srht_variant_for(total) == 0 ? 1 : 2;
    const int64_t target = (int64_t)sm_count() * ctas_per_sm * 8;
    int l = 1;
    while (l < maxl && (total >> (l + 1)) >= target) ++l;
    if (l < 2 && maxl >= 2) l = 2;
    return l;
}

static int64_t valid_chunks(const rla_srht_plan *p, int log2L) {
    const int64_t L = int64_t(1) << log2L;
    return (p->ntiles_valid + L - 1) / L;
}

extern "C" size_t rla_srht_workspace_bytes(const rla_srht_plan *p, int64_t m) {
    if (!p || m <= 0) return 0;
    // the chunk count depends on m only through choose_log2L; take the max over the two
    // extreme choices so that a forced/tuned L never overflows the buffer
    const int64_t nch = valid_chunks(p, choose_log2L(p, m));
    return (size_t)p->npass * nch * m * MAX_NSLOT * CTA * p->elem_bytes;      // the larger of the two descriptor sets
}

// 0: warp-specialised 384-thread CTAs, one per SM (default);
// 1: single-role 128-thread CTAs, two per SM: shorter prologue, better for tiny problems
//    (a few tiles per SM), where launch and drain dominate.  RLA_SRHT_VARIANT overrides.
static int srht_variant_for(int64_t total_tiles) {
    const char *e = getenv("RLA_SRHT_VARIANT");      // read per call: the parity tests switch it
    const int v = e ? atoi(e) : -1;
    if (v >= 0) return v;
    return total_tiles >= 16384 ? 0 : 1;
}

template <typename T, int NSLOT>
static int launch_main(const SrhtArgs<T> &a, int64_t grid, cudaStream_t st) {
    if (srht_variant_for(a.m * a.ntiles_valid) == 0) {
        auto kern = srht_ws_kernel<T, NSLOT>;
        const int smem = WS_GROUPS * TILE * sizeof(T);
        RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)grid, WS_THREADS, smem, st>>>(a);
    } else {
        auto kern = srht_main_kernel<T, NSLOT>;
        const int smem = 2 * TILE * sizeof(T) + NSLOT * CTA * 4;
        RLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)grid, CTA, smem, st>>>(a);
    }
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

template <typename T>
static int dispatch_nslot(int nslot, const SrhtArgs<T> &a, int64_t grid, cudaStream_t st) {
    switch (nslot) {
        case 4: return launch_main<T, 4>(a, grid, st);
        case 8: return launch_main<T, 8>(a, grid, st);
        case 16: return launch_main<T, 16>(a, grid, st);
        case 32: return launch_main<T, 32>(a, grid, st);
    }
    return fail(RLA_ERR_INVALID, "srht: bad nslot %d", nslot);
}

template <typename T>
static int srht_apply(const rla_srht_plan *p, const T *x, int64_t m, int64_t ldx, T scale, T *y,
                      int64_t ldy, void *ws, size_t ws_bytes, void *stream) {
    RLA_REQUIRE(p && p->dev, "rla_srht_apply: plan not uploaded");
    RLA_REQUIRE(p->elem_bytes == (int)sizeof(T), "rla_srht_apply: plan was built for %d-byte elements", p->elem_bytes);
    RLA_REQUIRE(m >= 0 && ldx >= p->n && ldy >= p->k, "rla_srht_apply: bad m/ldx/ldy");
    if (m == 0 || p->k == 0) return RLA_OK;
    RLA_REQUIRE(x && y && ws, "rla_srht_apply: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int log2L = choose_log2L(p, m);
    const int64_t nch = valid_chunks(p, log2L);
    const int variant = srht_variant_for(m * p->ntiles_valid);
    const SrhtDescSet &set = p->sets[variant == 0 ? 1 : 0];
    const int nst = set.nslot * CTA;
    const size_t need = (size_t)p->npass * nch * m * nst * sizeof(T);
    if (ws_bytes < need) return fail(RLA_ERR_WORKSPACE, "rla_srht_apply: workspace %zu < %zu bytes", ws_bytes, need);
    const int64_t grid = m * nch;
    RLA_REQUIRE(grid < (int64_t(1) << 31), "rla_srht_apply: grid too large");
    const bool vec = (reinterpret_cast<uintptr_t>(x) % (2 * sizeof(T)) == 0) && (ldx % 2 == 0);
    for (int pass = 0; pass < p->npass; ++pass) {
        SrhtArgs<T> a;
        a.x = x; a.ldx = ldx; a.n = p->n; a.m = m;
        a.signw = reinterpret_cast<const uint64_t *>(p->dev + p->off_sign);
        a.desc = reinterpret_cast<const uint32_t *>(p->dev + set.off_desc) + (size_t)pass * nst;
        a.ws = static_cast<T *>(ws) + (size_t)pass * nch * m * nst;
        a.log2L = log2L; a.nchunks = nch; a.ntiles_valid = p->ntiles_valid; a.aligned = vec ? 1 : 0;
        int rc = dispatch_nslot<T>(set.nslot, a, grid, st);
        if (rc != RLA_OK) return rc;
    }
    const int fin_threads = 256;
    dim3 fgrid((unsigned)((p->k + fin_threads - 1) / fin_threads), 1, 1);
    // gridDim.y is limited to 65535: loop over row blocks
    for (int64_t r0 = 0; r0 < m; r0 += 65535) {
        const int64_t rows = std::min<int64_t>(65535, m - r0);
        fgrid.y = (unsigned)rows;
        srht_finalize_kernel<T><<<fgrid, fin_threads, 0, st>>>(
            static_cast<const T *>(ws) + r0 * nst, reinterpret_cast<const int32_t *>(p->dev + set.off_slot),
            y + r0 * ldy, ldy, p->k, m, nch, nst, (int64_t)nch * m * nst, scale);
        count_launch();
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

extern "C" int rla_srht_apply_f64(const rla_srht_plan *p, const double *x, int64_t m, int64_t ldx,
                                  double scale, double *y, int64_t ldy, void *ws, size_t ws_bytes, void *stream) {
    return srht_apply<double>(p, x, m, ldx, scale, y, ldy, ws, ws_bytes, stream);
}
extern "C" int rla_srht_apply_f32(const rla_srht_plan *p, const float *x, int64_t m, int64_t ldx,
                                  float scale, float *y, int64_t ldy, void *ws, size_t ws_bytes, void *stream) {
    return srht_apply<float>(p, x, m, ldx, scale, y, ldy, ws, ws_bytes, stream);
}

// ------------------------------------------------------------- explicit rows
namespace rla {
__global__ void srht_rows_kernel(const int8_t *__restrict__ signs, int64_t n, const int64_t *__restrict__ idx,
                                 const int64_t *__restrict__ rows, double value, double *__restrict__ out,
                                 int64_t ldo) {
    const int64_t i = blockIdx.y;
    const uint32_t s = (uint32_t)idx[rows[i]];
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t par = (__popc(s & (uint32_t)j) & 1u) ^ (signs[j] < 0 ? 1u : 0u);
        out[i * ldo + j] = par ? -value : value;
    }
}
}  // namespace rla

extern "C" int rla_srht_rows_f64(const int8_t *signs, int64_t n, const int64_t *idx, const int64_t *rows,
                                 int64_t nrows, double value, double *out, int64_t ldo, void *stream) {
    RLA_REQUIRE(n >= 1 && nrows >= 0 && ldo >= n, "rla_srht_rows_f64: bad sizes");
    if (nrows == 0) return RLA_OK;
    RLA_REQUIRE(signs && idx && rows && out, "rla_srht_rows_f64: null pointer");
    for (int64_t r0 = 0; r0 < nrows; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, nrows - r0);
        dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, 4096), (unsigned)nr);
        srht_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(signs, n, idx, rows + r0, value, out + r0 * ldo, ldo);
        count_launch();
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
