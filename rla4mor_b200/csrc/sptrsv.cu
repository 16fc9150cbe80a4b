// sptrsv.cu -- sparse triangular solves with a block of right-hand sides: the device side of
// InverseLuOperator.apply (utilities/factorization.py:118-124: `slu.solve(V.T).T`), i.e. the
// `inverse_product` R^-1 applied in front of every sketch (mor/sketched_reductor.py:69,73;
// SURVEY.md section 8f rank 1).  The sparse LU factorisation itself stays on the host (SciPy
// SuperLU, as in the reference); its factors  Pr A Pc = L U  are uploaded once as CSR.
//
// Layout: the block (m, n) of the reference (one vector per row) is transposed ON THE WAY IN to
// X (n, ldx) with the m right-hand sides contiguous, with the row permutation applied by the same
// kernel; the solves run in place on X; the way out transposes back and applies the column
// permutation.  A row of a factor then reads, for every off-diagonal entry (i, j), the 8*m
// contiguous bytes X[j, :] (one 512-byte request per warp at m = 64): HBM/L2-bound.
//
// Scheduling is by LEVEL SETS computed on the host (rla_sptrsv_plan_host): rows of one level only
// depend on earlier levels.
//   * wide levels: one launch per level, one warp per row and per chunk of 64 right-hand sides;
//   * runs of narrow levels (<= 32 rows each -- the tail of the elimination tree: at n = 1e6 for a
//     2-D FEM factor, 4800 of 5161 levels, 16e3 rows holding HALF of the entries) are cut into
//     GROUPS of 32 consecutive rows.  One launch per group: CTA r sums the entries of row r whose
//     columns were solved by earlier launches (8 warps, 4 loads in flight each); the last CTA to
//     arrive (ticket counter, no spin-wait) resolves the dependencies INSIDE the group
//     sequentially out of shared memory.  500 launches instead of 4800 sequential levels.
// Deterministic: every sum has a fixed order.
#include "common.cuh"
#include <algorithm>
#include <utility>
#include <vector>

namespace rla {

constexpr int TRSV_RHS = 64;                          // right-hand sides per chunk (2 per lane)

__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

// X[dest(i), c] = B[c, i]   (dest = perm[i] or i)
__global__ void trsv_in_kernel(const double *__restrict__ B, int64_t m, int64_t n, int64_t ldb,
                               const int32_t *__restrict__ perm, double *__restrict__ X, int64_t ldx) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t c = c0 + y, i = i0 + threadIdx.x;
        tile[y][threadIdx.x] = (c < m && i < n) ? B[c * ldb + i] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t i = i0 + y, c = c0 + threadIdx.x;
        if (i < n && c < ldx) X[(int64_t)(perm ? perm[i] : i) * ldx + c] = tile[threadIdx.x][y];
    }
}

// out[c, i] = X[src(i), c]   (src = perm[i] or i)
__global__ void trsv_out_kernel(const double *__restrict__ X, int64_t m, int64_t n, int64_t ldx,
                                const int32_t *__restrict__ perm, double *__restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t i = i0 + y, c = c0 + threadIdx.x;
        tile[y][threadIdx.x] = (i < n && c < m) ? X[(int64_t)(perm ? perm[i] : i) * ldx + c] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t c = c0 + y, i = i0 + threadIdx.x;
        if (c < m && i < n) out[c * ldo + i] = tile[threadIdx.x][y];
    }
}

// Everything the kernels touch is stored IN SCHEDULE ORDER: row p of X, rowptr, split, diag is the
// p-th row of the processing order, and external column indices are positions too -- no
// order / inverse-order indirection on the critical path of a launch.
// out[p, :] = in[map[p], :]  (re-ordering of the block between the L and the U schedule)
__global__ void trsv_permute_rows_kernel(const double *__restrict__ in, const int32_t *__restrict__ map,
                                         double *__restrict__ out, int64_t n, int64_t ldx) {
    const int64_t w2 = ldx / 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * w2; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = t / w2, c = 2 * (t - p * w2);
        *reinterpret_cast<double2 *>(out + p * ldx + c) =
            *reinterpret_cast<const double2 *>(in + (int64_t)map[p] * ldx + c);
    }
}

struct TrsvArgs {
    const int64_t *rowptr;      // CSR of the strictly triangular part, rows and columns renumbered by
    const int32_t *col;         //   schedule position, entries of a row sorted by column
    const double *val;
    const double *diag;         // diagonal by position, or null for a unit diagonal
    const int64_t *split;       // per row: first entry whose column lies in the row's own group
    double *X;
    double *E;                  // scratch: external sums of the multi-row groups of a launch
    unsigned int *counter;      // scratch: arrival counters (zero between launches)
    int64_t ldx, m;
};

constexpr int TRSV_GROUP = 32;                        // rows per group of the narrow tail

constexpr int TRSV_WARPS = 16;                        // warps of a CTA that owns one (long) row
constexpr int TRSV_WARPS_SMALL = 4;                   // ... in steps of many short rows

// -sum_e val[e] * X[col[e], c..c+1] over the entries [e0, e1) of one row, for warp `w` of `nw`
// warps sharing the row.  A warp takes 32 consecutive entries at a time: (col, val) come in with
// ONE coalesced load per lane and are broadcast by shuffles; the X rows are fetched 8 at a time
// (8 independent 16-byte loads in flight per lane).  Every lane of the warp must call this.
template <int INFLIGHT>
__device__ __forceinline__ double2 row_sum_warp(const TrsvArgs &a, int64_t e0, int64_t e1, int64_t c, bool live,
                                                int w, int nw) {
    const int lane = threadIdx.x & 31;
    double2 acc = make_double2(0.0, 0.0);
    int64_t base = e0 + 32 * (int64_t)w;
    int32_t jn = 0;
    double vn = 0.0;
    if (base + lane < e1) { jn = a.col[base + lane]; vn = a.val[base + lane]; }
    for (; base < e1; base += 32 * (int64_t)nw) {
        const int32_t jc = jn;
        const double vc = vn;
        // prefetch the (col, val) of this warp's next 32 entries while the X rows of these are fetched
        const int64_t nb = base + 32 * (int64_t)nw + lane;
        jn = 0; vn = 0.0;
        if (nb < e1) { jn = a.col[nb]; vn = a.val[nb]; }
        const int cnt = (int)min((int64_t)32, e1 - base);
        for (int t0 = 0; t0 < cnt; t0 += INFLIGHT) {    // INFLIGHT independent 16-byte loads in flight per lane
            int32_t j[INFLIGHT];
            double v[INFLIGHT];
            double2 x[INFLIGHT];
#pragma unroll
            for (int u = 0; u < INFLIGHT; ++u) {
                j[u] = __shfl_sync(0xffffffffu, jc, (t0 + u) & 31);
                v[u] = __shfl_sync(0xffffffffu, vc, (t0 + u) & 31);
            }
#pragma unroll
            for (int u = 0; u < INFLIGHT; ++u)
                x[u] = (live && t0 + u < cnt) ? ldcg2(a.X + (int64_t)j[u] * a.ldx + c) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < INFLIGHT; ++u) {
                if (t0 + u < cnt) { acc.x = fma(-v[u], x[u].x, acc.x); acc.y = fma(-v[u], x[u].y, acc.y); }
            }
        }
    }
    return acc;
}

// one wide level: warp = one row x one chunk of 64 right-hand sides
__global__ void __launch_bounds__(256)
trsv_wide_kernel(const TrsvArgs a, int64_t lo, int64_t hi) {
    const int lane = threadIdx.x & 31;
    const int64_t idx = lo + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= hi) return;                             // whole warp
    const int64_t c = (int64_t)blockIdx.y * TRSV_RHS + 2 * lane;
    const bool live = c < a.ldx;
    const int64_t i = idx;
    double2 acc = row_sum_warp<8>(a, a.rowptr[i], a.rowptr[i + 1], c, live, 0, 1);
    if (live) {
        double *xi = a.X + i * a.ldx + c;
        const double2 b = *reinterpret_cast<const double2 *>(xi);
        acc.x += b.x; acc.y += b.y;
        if (a.diag) { const double d = a.diag[i]; acc.x /= d; acc.y /= d; }
        *reinterpret_cast<double2 *>(xi) = acc;
    }
}

// the same sum with the row split over the TRSV_WARPS warps of the CTA, combined in warp order;
// result valid in warp 0
template <int NW>
__device__ __forceinline__ double2 row_sum_cta(const TrsvArgs &a, int64_t e0, int64_t e1, int64_t c, bool live,
                                               double2 (*part)[32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 acc = row_sum_warp<(NW >= 16 ? 16 : 8)>(a, e0, e1, c, live, warp, NW);
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = part[0][lane];
#pragma unroll
        for (int w = 1; w < NW; ++w) { acc.x += part[w][lane].x; acc.y += part[w][lane].y; }
    }
    return acc;
}

// One STEP of the schedule: a set of independent GROUPS.  A group is up to 32 rows that depend on
// each other (a chain of consecutive levels inside one supernode of the factor) and, outside the
// group, only on rows solved by earlier launches.  blockIdx.x = row (of some group of the launch),
// blockIdx.y = chunk of right-hand sides.  CTA (r, chunk, g) sums the EXTERNAL entries of
// its row with all its warps; a single-row group is finished right there; otherwise the last CTA
// of the group to arrive (ticket counter, no spin-wait) stages the group's internal entries and
// right-hand sides in shared memory and resolves the internal dependencies sequentially (one
// warp, lanes = right-hand sides).  One launch per ~32 levels instead of one per level.
template <int NW>
__global__ void __launch_bounds__(32 * NW, NW >= 16 ? 1 : 8)
trsv_groups_kernel(const TrsvArgs a, const int64_t *__restrict__ grp_start, const int32_t *__restrict__ grp_rows,
                   const int32_t *__restrict__ grp_of_pos, int64_t g_first, int64_t p_first) {
    __shared__ double2 part[NW][32];
    __shared__ double2 xs[2][32];                      // x of the row just resolved, double buffered
    __shared__ double D[TRSV_GROUP][TRSV_GROUP + 1];   // dense image of the group's internal entries
    __shared__ double s_diag[TRSV_GROUP];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // CTA <-> one row: order position p_first + blockIdx.x (the rows of the groups of a launch are
    // contiguous in the order), no idle CTAs
    const int64_t p = p_first + blockIdx.x;
    const int64_t gid = grp_of_pos[p];
    const int nrows = grp_rows[gid];
    const int64_t g0 = grp_start[gid];
    const int r = (int)(p - g0);
    const int64_t eslot = gid - g_first;               // multi-row groups of a launch: scratch slot
    const int64_t c = (int64_t)blockIdx.y * TRSV_RHS + 2 * lane;
    const bool live = c < a.ldx;
    const int64_t i = p;
    double2 acc = row_sum_cta<NW>(a, a.rowptr[i], a.split[i], c, live, part);
    if (nrows == 1) {                                  // nothing internal: finish the row here
        if (warp == 0 && live) {
            double *xi = a.X + i * a.ldx + c;
            const double2 b = *reinterpret_cast<const double2 *>(xi);
            acc.x += b.x; acc.y += b.y;
            if (a.diag) { const double d = a.diag[i]; acc.x /= d; acc.y /= d; }
            *reinterpret_cast<double2 *>(xi) = acc;
        }
        return;
    }
    double *E = a.E + eslot * TRSV_GROUP * a.ldx;
    unsigned int *counter = a.counter + eslot * gridDim.y + blockIdx.y;
    if (warp == 0) {                                   // the writers fence, then one thread takes the ticket
        if (live) *reinterpret_cast<double2 *>(E + (int64_t)r * a.ldx + c) = acc;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            const unsigned int ticket = atomicAdd(counter, 1u);
            s_last = ticket == (unsigned int)nrows - 1;
            if (s_last) {
                *counter = 0;                          // ready for the next launch
                __threadfence();
            }
        }
    }
    __syncthreads();
    if (!s_last) return;
    // Stage the group: its internal entries as a dense 32 x 32 lower-triangular image D, and
    // b + external sum of the two rows (warp, warp + 16) this warp owns, in registers.
    constexpr int RPW = TRSV_GROUP / NW;               // rows of the group per warp: q = warp + h * NW
    for (int t = threadIdx.x; t < TRSV_GROUP * (TRSV_GROUP + 1); t += blockDim.x) (&D[0][0])[t] = 0.0;
    if (threadIdx.x < nrows) s_diag[threadIdx.x] = a.diag ? a.diag[g0 + threadIdx.x] : 1.0;
    __syncthreads();
    double2 acc2[RPW];
#pragma unroll
    for (int h = 0; h < RPW; ++h) {
        const int q = warp + h * NW;
        acc2[h] = make_double2(0.0, 0.0);
        if (q < nrows) {
            const int64_t iq = g0 + q;
            const int64_t e0 = a.split[iq];
            const int cnt = (int)(a.rowptr[iq + 1] - e0);
            for (int t = lane; t < cnt; t += 32) D[q][a.col[e0 + t]] = a.val[e0 + t];
            if (live) {
                const double2 b = ldcg2(a.X + iq * a.ldx + c), ext = ldcg2(E + (int64_t)q * a.ldx + c);
                acc2[h] = make_double2(b.x + ext.x, b.y + ext.y);
            }
        }
    }
    __syncthreads();
    // Right-looking resolve: row q becomes final, every later row subtracts its entry (r, q) times
    // x_q.  One CTA barrier per row, rows spread over the warps, lanes = right-hand sides.
    for (int q = 0; q < nrows; ++q) {
        if ((q % NW) == warp) {
            double2 x = acc2[0];                       // acc2[q / NW] with static register indices
#pragma unroll
            for (int h = 1; h < RPW; ++h) if (q >= h * NW) x = acc2[h];
            if (a.diag) { const double d = s_diag[q]; x.x /= d; x.y /= d; }
            xs[q & 1][lane] = x;
            if (live) *reinterpret_cast<double2 *>(a.X + (g0 + q) * a.ldx + c) = x;
        }
        __syncthreads();
        const double2 x = xs[q & 1][lane];
#pragma unroll
        for (int h = 0; h < RPW; ++h) {
            const int r2 = warp + h * NW;
            if (r2 > q && r2 < nrows) {
                const double v = D[r2][q];
                acc2[h].x = fma(-v, x.x, acc2[h].x); acc2[h].y = fma(-v, x.y, acc2[h].y);
            }
        }
    }
}

}  // namespace rla

using namespace rla;

// Host-side analysis of one triangular CSR factor (HOST arrays in, HOST arrays out; plain C++):
//   level_out[i]   longest dependency chain ending in row i (lower != 0: deps j < i, else j > i)
//   order_out      the processing order of the rows, pos_out its inverse
//   rowptr2/col2/val2  the strictly triangular part, entries of each row sorted by pos[col]
//   diag_out       the diagonal (1.0 where absent)
//   groups         group g = order positions [grp_start[g], grp_start[g] + grp_rows[g]), <= group_rows rows
//   steps          kind 0: ONE level of more than `wide_min` rows, order positions [step_lo, step_hi),
//                  rows independent; kind 1: groups [step_lo, step_hi), independent of each other, the
//                  multi-row groups first, then (from step_mid on) the single-row groups.
//   split_out[i]   first entry of row i that refers to a row of its own group (row end when there is
//                  none); for those entries col2 holds the slot of the column inside the group.
// A kind-1 step is a BAND of up to `group_rows` consecutive levels whose connected components
// (dependencies inside the band) all have at most `group_rows` rows: each component is a group.  The
// band is shortened until that holds (a single level always qualifies: singleton groups).  At most
// max_multi multi-row groups per step (they need scratch).  Arrays of n entries each.
extern "C" int rla_sptrsv_plan_host(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val,
                                    int lower, int wide_min, int group_rows, int max_multi,
                                    int32_t *level_out, int32_t *order_out, int32_t *pos_out,
                                    int64_t *rowptr2, int32_t *col2, double *val2, double *diag_out,
                                    int64_t *split_out, int64_t *grp_start, int32_t *grp_rows, int64_t *ngroups_out,
                                    int64_t *step_lo, int64_t *step_mid, int64_t *step_hi, int32_t *step_kind,
                                    int64_t *nsteps_out, int32_t *nlevels_out) {
    RLA_REQUIRE(n >= 0 && rowptr && level_out && order_out && pos_out && rowptr2 && diag_out && split_out &&
                grp_start && grp_rows && ngroups_out && step_lo && step_mid && step_hi && step_kind && nsteps_out &&
                nlevels_out, "rla_sptrsv_plan_host: null pointer");
    RLA_REQUIRE(wide_min >= 1 && group_rows >= 1 && group_rows <= TRSV_GROUP && max_multi >= 1,
                "rla_sptrsv_plan_host: bad parameters");
    int32_t nl = 0;
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = lower ? t : n - 1 - t;
        int32_t l = 0;
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int32_t j = col[e];
            RLA_REQUIRE(j >= 0 && j < n && (lower ? j <= i : j >= i),
                        "rla_sptrsv_plan_host: entry (%lld, %d) on the wrong side of the diagonal", (long long)i, j);
            if (j != i) l = std::max(l, level_out[j] + 1);
        }
        level_out[i] = l;
        nl = std::max(nl, l + 1);
    }
    *nlevels_out = nl;
    // counting sort by level
    std::vector<int64_t> lvlptr((size_t)nl + 1, 0);
    for (int64_t i = 0; i < n; ++i) ++lvlptr[(size_t)level_out[i] + 1];
    for (int32_t l = 0; l < nl; ++l) lvlptr[(size_t)l + 1] += lvlptr[l];
    {
        std::vector<int64_t> fill(lvlptr.begin(), lvlptr.end() - 1);
        for (int64_t i = 0; i < n; ++i) {
            const int64_t p = fill[level_out[i]]++;
            order_out[p] = (int32_t)i;
            pos_out[i] = (int32_t)p;
        }
    }
    std::vector<int64_t> gstart_of_pos((size_t)n, -1);  // start position of the group of the row at a position
    std::vector<int32_t> parent, csize, band_rows, comp_first;
    int64_t ns = 0, ng = 0;
    int32_t l = 0;
    while (l < nl) {
        const int64_t rows = lvlptr[(size_t)l + 1] - lvlptr[l];
        if (rows > wide_min) {
            step_lo[ns] = lvlptr[l]; step_mid[ns] = lvlptr[(size_t)l + 1]; step_hi[ns] = lvlptr[(size_t)l + 1];
            step_kind[ns] = 0; ++ns;
            ++l;
            continue;
        }
        // longest band [l, l + W) of non-wide levels whose components have at most group_rows rows
        int32_t W = 1;
        while (W < group_rows && l + W < nl && lvlptr[(size_t)l + W + 1] - lvlptr[(size_t)l + W] <= wide_min) ++W;
        const int64_t p0 = lvlptr[l];
        for (;; W = std::max(1, W / 2)) {
            const int64_t p1 = lvlptr[(size_t)l + W];
            const int64_t nb = p1 - p0;
            parent.resize((size_t)nb); csize.assign((size_t)nb, 1);
            for (int64_t t = 0; t < nb; ++t) parent[t] = (int32_t)t;
            bool ok = true;
            if (W > 1) {
                auto find = [&](int32_t x) {
                    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
                    return x;
                };
                for (int64_t p = p0; p < p1 && ok; ++p) {
                    const int64_t i = order_out[p];
                    for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                        const int32_t j = col[e];
                        if (j == i || pos_out[j] < p0) continue;      // outside the band (earlier levels)
                        int32_t ra = find((int32_t)(p - p0)), rb = find((int32_t)(pos_out[j] - p0));
                        if (ra == rb) continue;
                        if (csize[ra] < csize[rb]) std::swap(ra, rb);
                        parent[rb] = ra;
                        csize[ra] += csize[rb];
                        if (csize[ra] > group_rows) { ok = false; break; }
                    }
                }
                if (ok) for (int64_t t = 0; t < nb; ++t) parent[t] = find((int32_t)t);
            }
            if (!ok) continue;                                         // shorten the band
            // accepted: components -> groups; rows of a component stay in level order (stable)
            band_rows.resize((size_t)nb);
            for (int64_t t = 0; t < nb; ++t) band_rows[t] = order_out[p0 + t];
            // multi-row components first (in order of first appearance), then singletons
            comp_first.assign((size_t)nb, -1);
            std::vector<std::vector<int32_t>> multi;
            std::vector<int32_t> single;
            for (int64_t t = 0; t < nb; ++t) {
                const int32_t root = parent[t];
                if (csize[root] == 1) { single.push_back((int32_t)t); continue; }
                if (comp_first[root] < 0) { comp_first[root] = (int32_t)multi.size(); multi.emplace_back(); }
                multi[comp_first[root]].push_back((int32_t)t);
            }
            // longest rows first: a step ends with its slowest CTA, so the long rows (ancestor
            // separators) must start at once instead of queueing behind thousands of short ones
            auto rowlen = [&](int32_t t) { const int64_t i = band_rows[t]; return rowptr[i + 1] - rowptr[i]; };
            std::vector<int64_t> mlen(multi.size(), 0);
            for (size_t g = 0; g < multi.size(); ++g)
                for (int32_t t : multi[g]) mlen[g] = std::max(mlen[g], rowlen(t));
            std::vector<size_t> mord(multi.size());
            for (size_t g = 0; g < multi.size(); ++g) mord[g] = g;
            std::stable_sort(mord.begin(), mord.end(), [&](size_t x, size_t y) { return mlen[x] > mlen[y]; });
            {
                std::vector<std::vector<int32_t>> sorted;
                sorted.reserve(multi.size());
                for (size_t g : mord) sorted.push_back(std::move(multi[g]));
                multi.swap(sorted);
            }
            std::stable_sort(single.begin(), single.end(), [&](int32_t x, int32_t y) { return rowlen(x) > rowlen(y); });
            int64_t p = p0;
            size_t mi = 0, si = 0;
            // emit steps: at most max_multi multi-row groups (and at most 65535 groups) per step
            while (mi < multi.size() || si < single.size()) {
                step_kind[ns] = 1;
                step_lo[ns] = ng;
                size_t took = 0;
                for (; mi < multi.size() && took < (size_t)max_multi; ++mi, ++took) {
                    grp_start[ng] = p; grp_rows[ng] = (int32_t)multi[mi].size();
                    for (int32_t t : multi[mi]) {
                        order_out[p] = band_rows[t]; pos_out[band_rows[t]] = (int32_t)p; gstart_of_pos[p] = grp_start[ng];
                        ++p;
                    }
                    ++ng;
                }
                step_mid[ns] = ng;
                if (mi == multi.size()) {
                    for (took = 0; si < single.size() && took < 65535; ++si, ++took) {
                        const int32_t t = single[si];
                        grp_start[ng] = p; grp_rows[ng] = 1;
                        order_out[p] = band_rows[t]; pos_out[band_rows[t]] = (int32_t)p; gstart_of_pos[p] = p;
                        ++p; ++ng;
                    }
                }
                step_hi[ns] = ng;
                ++ns;
            }
            l += W;
            break;
        }
    }
    *nsteps_out = ns;
    *ngroups_out = ng;
    // strictly triangular CSR IN SCHEDULE ORDER: row p = row order[p], external columns = positions
    std::vector<std::pair<int32_t, int64_t>> tmp;
    int64_t w = 0;
    for (int64_t p = 0; p < n; ++p) {
        const int64_t i = order_out[p];
        rowptr2[p] = w;
        diag_out[p] = 1.0;
        tmp.clear();
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            if (col[e] == i) { diag_out[p] = val[e]; continue; }
            tmp.emplace_back(pos_out[col[e]], e);
        }
        std::sort(tmp.begin(), tmp.end());
        const int64_t gs = gstart_of_pos[p];
        int64_t sp = -1;
        for (const auto &pe : tmp) {
            if (sp < 0 && gs >= 0 && pe.first >= gs) sp = w;
            col2[w] = sp >= 0 ? (int32_t)(pe.first - gs) : pe.first;   // internal: slot inside the group
            val2[w] = val[pe.second];
            ++w;
        }
        split_out[p] = sp < 0 ? w : sp;
    }
    rowptr2[n] = w;
    return RLA_OK;
}

extern "C" int rla_sptrsv_transpose_in_f64(const double *b_dev, int64_t m, int64_t n, int64_t ldb,
                                           const int32_t *perm_dev, double *x_dev, int64_t ldx, void *stream) {
    RLA_REQUIRE(m >= 0 && n >= 0 && ldb >= n && ldx >= m && (ldx & 1) == 0, "rla_sptrsv_transpose_in_f64: bad sizes");
    if (m == 0 || n == 0) return RLA_OK;
    RLA_REQUIRE(b_dev && x_dev, "rla_sptrsv_transpose_in_f64: null pointer");
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((ldx + 31) / 32));
    RLA_REQUIRE(grid.y <= 65535, "rla_sptrsv_transpose_in_f64: too many right-hand sides");
    trsv_in_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(b_dev, m, n, ldb, perm_dev, x_dev, ldx);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

extern "C" int rla_sptrsv_transpose_out_f64(const double *x_dev, int64_t m, int64_t n, int64_t ldx,
                                            const int32_t *perm_dev, double *out_dev, int64_t ldo, void *stream) {
    RLA_REQUIRE(m >= 0 && n >= 0 && ldo >= n && ldx >= m, "rla_sptrsv_transpose_out_f64: bad sizes");
    if (m == 0 || n == 0) return RLA_OK;
    RLA_REQUIRE(x_dev && out_dev, "rla_sptrsv_transpose_out_f64: null pointer");
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
    RLA_REQUIRE(grid.y <= 65535, "rla_sptrsv_transpose_out_f64: too many right-hand sides");
    trsv_out_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x_dev, m, n, ldx, perm_dev, out_dev, ldo);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

// In-place triangular solve T X = X on the (n, ldx) block (right-hand sides contiguous), driven by
// the step list of rla_sptrsv_plan_host (HOST arrays step_*); all other arrays on the device.
// scratch_dev: rla_sptrsv_scratch_bytes(ldx, max_multi) bytes, zero-filled once by the caller.
extern "C" size_t rla_sptrsv_scratch_bytes(int64_t ldx, int max_multi) {
    const size_t chunks = (size_t)((ldx + TRSV_RHS - 1) / TRSV_RHS);
    return (size_t)max_multi * ((size_t)TRSV_GROUP * (size_t)ldx * sizeof(double) + chunks * sizeof(unsigned int)) + 64;
}

extern "C" int rla_sptrsv_solve_f64(const int64_t *rowptr_dev, const int32_t *col_dev, const double *val_dev,
                                    const double *diag_dev, const int64_t *split_dev, const int64_t *grp_start_dev, const int32_t *grp_rows_dev,
                                    const int32_t *grp_of_pos_dev, const int64_t *grp_start_host, const int64_t *grp_csum_host,
                                    const int64_t *step_lo, const int64_t *step_mid, const int64_t *step_hi,
                                    const int32_t *step_kind, int64_t nsteps, int max_multi,
                                    double *x_dev, int64_t m, int64_t ldx,
                                    void *scratch_dev, size_t scratch_bytes, void *stream) {
    RLA_REQUIRE(nsteps >= 0 && m >= 0 && ldx >= m && (ldx & 1) == 0 && max_multi >= 1, "rla_sptrsv_solve_f64: bad sizes");
    if (nsteps == 0 || m == 0) return RLA_OK;
    RLA_REQUIRE(rowptr_dev && col_dev && val_dev && split_dev && grp_start_dev && grp_rows_dev &&
                grp_of_pos_dev && grp_start_host && grp_csum_host && step_lo && step_mid && step_hi && step_kind && x_dev && scratch_dev, "rla_sptrsv_solve_f64: null pointer");
    RLA_REQUIRE(((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)scratch_dev & 15) == 0,
                "rla_sptrsv_solve_f64: X and scratch must be 16-byte aligned");
    if (scratch_bytes < rla_sptrsv_scratch_bytes(ldx, max_multi))
        return fail(RLA_ERR_WORKSPACE, "rla_sptrsv_solve_f64: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *E = static_cast<double *>(scratch_dev);
    unsigned int *counter = reinterpret_cast<unsigned int *>(E + (size_t)max_multi * TRSV_GROUP * ldx);
    TrsvArgs a = {rowptr_dev, col_dev, val_dev, diag_dev, split_dev, x_dev, E, counter, ldx, m};
    const unsigned chunks = (unsigned)((ldx + TRSV_RHS - 1) / TRSV_RHS);
    RLA_REQUIRE(chunks <= 65535, "rla_sptrsv_solve_f64: too many right-hand sides");
    for (int64_t s = 0; s < nsteps; ++s) {
        if (step_kind[s] == 0) {
            const int64_t rows = step_hi[s] - step_lo[s];
            RLA_REQUIRE(rows >= 1, "rla_sptrsv_solve_f64: empty step %lld", (long long)s);
            dim3 grid((unsigned)((rows + 7) / 8), chunks);
            trsv_wide_kernel<<<grid, 256, 0, st>>>(a, step_lo[s], step_hi[s]);
            count_launch();
            continue;
        }
        const int64_t nmulti = step_mid[s] - step_lo[s], nsingle = step_hi[s] - step_mid[s];
        RLA_REQUIRE(nmulti >= 0 && nsingle >= 0 && nmulti <= max_multi && nsingle <= 65535 && nmulti + nsingle >= 1,
                    "rla_sptrsv_solve_f64: bad step %lld", (long long)s);
        // the rows of the groups of a step are contiguous in the order: multi-row groups first
        // (grp_csum_host = running sum of grp_rows, ngroups + 1 entries), then the single rows
        const int64_t p0 = grp_start_host[step_lo[s]];
        const int64_t rows_multi = grp_csum_host[step_mid[s]] - grp_csum_host[step_lo[s]];
        {
            const int64_t rows = rows_multi + nsingle;
            dim3 grid((unsigned)rows, chunks);                        // one launch: multi-row groups, then single rows
            // few rows (near the root of the elimination tree, long rows): 16 warps share a row;
            // many rows (short): 4 warps per row so that every CTA of the step is resident at once
            if (rows < 1024)
                trsv_groups_kernel<TRSV_WARPS><<<grid, 32 * TRSV_WARPS, 0, st>>>(a, grp_start_dev, grp_rows_dev,
                                                                                grp_of_pos_dev, step_lo[s], p0);
            else
                trsv_groups_kernel<TRSV_WARPS_SMALL><<<grid, 32 * TRSV_WARPS_SMALL, 0, st>>>(
                    a, grp_start_dev, grp_rows_dev, grp_of_pos_dev, step_lo[s], p0);
            count_launch();
        }
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}

// out[p, :] = in[map[p], :] on (n, ldx) blocks: the re-ordering between the schedules of L and U
extern "C" int rla_sptrsv_permute_rows_f64(const double *in_dev, const int32_t *map_dev, double *out_dev, int64_t n,
                                           int64_t ldx, void *stream) {
    RLA_REQUIRE(n >= 0 && ldx >= 0 && (ldx & 1) == 0, "rla_sptrsv_permute_rows_f64: bad sizes");
    if (n == 0 || ldx == 0) return RLA_OK;
    RLA_REQUIRE(in_dev && map_dev && out_dev && in_dev != out_dev, "rla_sptrsv_permute_rows_f64: bad pointers");
    const int64_t work = n * (ldx / 2);
    const unsigned blocks = (unsigned)std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 16);
    trsv_permute_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in_dev, map_dev, out_dev, n, ldx);
    count_launch();
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
