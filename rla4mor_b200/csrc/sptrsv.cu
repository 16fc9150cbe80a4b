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

struct TrsvArgs {
    const int64_t *rowptr;      // CSR of the strictly triangular part, entries of a row sorted by the
    const int32_t *col;         //   position of their column in the level order
    const double *val;
    const double *diag;         // diagonal, or null for a unit diagonal
    const int32_t *order;       // rows sorted by level
    const int32_t *pos;         // inverse of order
    const int64_t *split;       // per row: first entry whose column lies in the row's own group
    double *X;
    double *E;                  // scratch: external sums of the current group, TRSV_GROUP x ldx
    unsigned int *counter;      // scratch: one arrival counter per chunk of right-hand sides (zero)
    int64_t ldx, m;
};

constexpr int TRSV_GROUP = 32;                        // rows per group of the narrow tail

// one level: warp = one row x one chunk of 64 right-hand sides
__global__ void __launch_bounds__(256)
trsv_wide_kernel(const TrsvArgs a, int64_t lo, int64_t hi) {
    const int lane = threadIdx.x & 31;
    const int64_t idx = lo + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= hi) return;
    const int64_t c = (int64_t)blockIdx.y * TRSV_RHS + 2 * lane;
    if (c >= a.ldx) return;
    const int64_t i = a.order[idx];
    double *xi = a.X + i * a.ldx + c;
    double2 acc = *reinterpret_cast<const double2 *>(xi);
    const int64_t e1 = a.rowptr[i + 1];
    int64_t e = a.rowptr[i];
    for (; e + 4 <= e1; e += 4) {                      // four independent 16-byte loads in flight
        int32_t j[4];
        double v[4];
        double2 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { j[u] = a.col[e + u]; v[u] = a.val[e + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ldcg2(a.X + (int64_t)j[u] * a.ldx + c);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x = fma(-v[u], x[u].x, acc.x); acc.y = fma(-v[u], x[u].y, acc.y); }
    }
    for (; e < e1; ++e) {
        const double v = a.val[e];
        const double2 x = ldcg2(a.X + (int64_t)a.col[e] * a.ldx + c);
        acc.x = fma(-v, x.x, acc.x); acc.y = fma(-v, x.y, acc.y);
    }
    if (a.diag) { const double d = a.diag[i]; acc.x /= d; acc.y /= d; }
    *reinterpret_cast<double2 *>(xi) = acc;
}

// A GROUP of up to 32 consecutive rows of the narrow tail (several levels).  CTA r sums the
// EXTERNAL entries of row r (columns solved by earlier launches) with its 8 warps; the last CTA
// to arrive then resolves the group's internal triangular dependencies sequentially from shared
// memory (one warp, lanes = right-hand sides).  One launch per group instead of one
// synchronisation per level, and no spin-wait anywhere.
__global__ void __launch_bounds__(256)
trsv_group_kernel(const TrsvArgs a, int64_t g0, int nrows) {
    __shared__ double2 part[8][32];
    __shared__ double2 xs[TRSV_GROUP][32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x;
    const int64_t c = (int64_t)blockIdx.y * TRSV_RHS + 2 * lane;
    const bool live = c < a.ldx;
    const int64_t i = a.order[g0 + r];
    {
        double2 acc = make_double2(0.0, 0.0);
        if (live) {
            const int64_t e1 = a.split[i];
            int64_t e = a.rowptr[i] + warp;
            for (; e + 24 < e1; e += 32) {             // four independent 16-byte loads in flight per lane
                int32_t j[4];
                double v[4];
                double2 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { j[u] = a.col[e + 8 * u]; v[u] = a.val[e + 8 * u]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = ldcg2(a.X + (int64_t)j[u] * a.ldx + c);
#pragma unroll
                for (int u = 0; u < 4; ++u) { acc.x = fma(-v[u], x[u].x, acc.x); acc.y = fma(-v[u], x[u].y, acc.y); }
            }
            for (; e < e1; e += 8) {
                const double v = a.val[e];
                const double2 x = ldcg2(a.X + (int64_t)a.col[e] * a.ldx + c);
                acc.x = fma(-v, x.x, acc.x); acc.y = fma(-v, x.y, acc.y);
            }
        }
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && live) {
            acc = part[0][lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) { acc.x += part[w][lane].x; acc.y += part[w][lane].y; }
            *reinterpret_cast<double2 *>(a.E + (int64_t)r * a.ldx + c) = acc;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(a.counter + blockIdx.y, 1u);
        s_last = ticket == (unsigned int)nrows - 1;
        if (s_last) a.counter[blockIdx.y] = 0;         // ready for the next group (next launch)
    }
    __syncthreads();
    if (!s_last || warp != 0 || !live) return;
    __threadfence();
    for (int q = 0; q < nrows; ++q) {
        const int64_t iq = a.order[g0 + q];
        double *xi = a.X + iq * a.ldx + c;
        double2 acc = ldcg2(xi);
        const double2 ext = ldcg2(a.E + (int64_t)q * a.ldx + c);
        acc.x += ext.x; acc.y += ext.y;
        const int64_t e1 = a.rowptr[iq + 1];
        for (int64_t e = a.split[iq]; e < e1; ++e) {
            const double v = a.val[e];
            const double2 x = xs[a.col[e]][lane];        // internal entries store the slot in the group
            acc.x = fma(-v, x.x, acc.x); acc.y = fma(-v, x.y, acc.y);
        }
        if (a.diag) { const double d = a.diag[iq]; acc.x /= d; acc.y /= d; }
        xs[q][lane] = acc;
        *reinterpret_cast<double2 *>(xi) = acc;
        __syncwarp();
    }
}

}  // namespace rla

using namespace rla;

// Host-side analysis of one triangular CSR factor (HOST arrays in, HOST arrays out):
//   level_out[i]   longest dependency chain ending in row i (lower != 0: deps j < i, else j > i)
//   order_out      rows sorted by level (stable), pos_out its inverse
//   rowptr2/col2/val2  the strictly triangular part, entries of each row sorted by pos[col]
//   diag_out       the diagonal (1.0 where absent)
//   steps: step s covers order positions [step_lo[s], step_hi[s]); step_kind 0 = one level of
//   more than `narrow` rows (rows independent), 1 = a group of at most `group_rows` rows cut from a
//   run of narrow levels (dependencies inside the group allowed); split_out[i] = first entry of
//   row i whose column is inside its own group (row end for rows of wide levels).
// Returns the number of steps through nsteps_out (arrays must hold n entries).
extern "C" int rla_sptrsv_plan_host(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val,
                                    int lower, int narrow, int group_rows,
                                    int32_t *level_out, int32_t *order_out, int32_t *pos_out,
                                    int64_t *rowptr2, int32_t *col2, double *val2, double *diag_out,
                                    int64_t *split_out, int64_t *step_lo, int64_t *step_hi, int32_t *step_kind,
                                    int64_t *nsteps_out, int32_t *nlevels_out) {
    RLA_REQUIRE(n >= 0 && rowptr && level_out && order_out && pos_out && rowptr2 && diag_out && split_out &&
                step_lo && step_hi && step_kind && nsteps_out && nlevels_out, "rla_sptrsv_plan_host: null pointer");
    RLA_REQUIRE(narrow >= 1 && group_rows >= 1 && group_rows <= TRSV_GROUP, "rla_sptrsv_plan_host: bad group size");
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
    // steps
    int64_t ns = 0;
    std::vector<int64_t> group_start((size_t)n, 0);    // per order position: start of its group (wide: row itself irrelevant)
    int32_t l = 0;
    while (l < nl) {
        const int64_t rows = lvlptr[(size_t)l + 1] - lvlptr[l];
        if (rows > narrow) {
            step_lo[ns] = lvlptr[l]; step_hi[ns] = lvlptr[(size_t)l + 1]; step_kind[ns] = 0; ++ns;
            for (int64_t p = lvlptr[l]; p < lvlptr[(size_t)l + 1]; ++p) group_start[p] = -1;
            ++l;
        } else {
            int32_t end = l + 1;
            while (end < nl && lvlptr[(size_t)end + 1] - lvlptr[end] <= narrow) ++end;
            for (int64_t p = lvlptr[l]; p < lvlptr[end]; p += group_rows) {
                const int64_t q = std::min<int64_t>(p + group_rows, lvlptr[end]);
                step_lo[ns] = p; step_hi[ns] = q; step_kind[ns] = 1; ++ns;
                for (int64_t t = p; t < q; ++t) group_start[t] = p;
            }
            l = end;
        }
    }
    *nsteps_out = ns;
    // strictly triangular CSR with entries sorted by the position of their column
    std::vector<std::pair<int32_t, int64_t>> tmp;
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) {
        rowptr2[i] = w;
        diag_out[i] = 1.0;
        tmp.clear();
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            if (col[e] == i) { diag_out[i] = val[e]; continue; }
            tmp.emplace_back(pos_out[col[e]], e);
        }
        std::sort(tmp.begin(), tmp.end());
        const int64_t gs = group_start[pos_out[i]];
        int64_t sp = -1;
        for (const auto &pe : tmp) {
            if (sp < 0 && gs >= 0 && pe.first >= gs) sp = w;
            col2[w] = sp >= 0 ? (int32_t)(pe.first - gs) : col[pe.second];   // internal: slot inside the group
            val2[w] = val[pe.second];
            ++w;
        }
        split_out[i] = sp < 0 ? w : sp;
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
// scratch_dev: rla_sptrsv_scratch_bytes(ldx) bytes, zero-filled once by the caller.
extern "C" size_t rla_sptrsv_scratch_bytes(int64_t ldx) {
    const size_t chunks = (size_t)((ldx + TRSV_RHS - 1) / TRSV_RHS);
    return (size_t)TRSV_GROUP * (size_t)ldx * sizeof(double) + chunks * sizeof(unsigned int) + 64;
}

extern "C" int rla_sptrsv_solve_f64(const int64_t *rowptr_dev, const int32_t *col_dev, const double *val_dev,
                                    const double *diag_dev, const int32_t *order_dev, const int32_t *pos_dev,
                                    const int64_t *split_dev,
                                    const int64_t *step_lo, const int64_t *step_hi, const int32_t *step_kind,
                                    int64_t nsteps, double *x_dev, int64_t m, int64_t ldx,
                                    void *scratch_dev, size_t scratch_bytes, void *stream) {
    RLA_REQUIRE(nsteps >= 0 && m >= 0 && ldx >= m && (ldx & 1) == 0, "rla_sptrsv_solve_f64: bad sizes");
    if (nsteps == 0 || m == 0) return RLA_OK;
    RLA_REQUIRE(rowptr_dev && col_dev && val_dev && order_dev && pos_dev && split_dev && step_lo && step_hi &&
                step_kind && x_dev && scratch_dev, "rla_sptrsv_solve_f64: null pointer");
    RLA_REQUIRE(((uintptr_t)x_dev & 15) == 0 && ((uintptr_t)scratch_dev & 15) == 0,
                "rla_sptrsv_solve_f64: X and scratch must be 16-byte aligned");
    if (scratch_bytes < rla_sptrsv_scratch_bytes(ldx))
        return fail(RLA_ERR_WORKSPACE, "rla_sptrsv_solve_f64: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *E = static_cast<double *>(scratch_dev);
    unsigned int *counter = reinterpret_cast<unsigned int *>(E + (size_t)TRSV_GROUP * ldx);
    TrsvArgs a = {rowptr_dev, col_dev, val_dev, diag_dev, order_dev, pos_dev, split_dev, x_dev, E, counter, ldx, m};
    const unsigned chunks = (unsigned)((ldx + TRSV_RHS - 1) / TRSV_RHS);
    RLA_REQUIRE(chunks <= 65535, "rla_sptrsv_solve_f64: too many right-hand sides");
    for (int64_t s = 0; s < nsteps; ++s) {
        const int64_t rows = step_hi[s] - step_lo[s];
        RLA_REQUIRE(rows >= 1, "rla_sptrsv_solve_f64: empty step %lld", (long long)s);
        if (step_kind[s] == 0) {
            dim3 grid((unsigned)((rows + 7) / 8), chunks);
            trsv_wide_kernel<<<grid, 256, 0, st>>>(a, step_lo[s], step_hi[s]);
        } else {
            RLA_REQUIRE(rows <= TRSV_GROUP, "rla_sptrsv_solve_f64: group of %lld rows", (long long)rows);
            dim3 grid((unsigned)rows, chunks);
            trsv_group_kernel<<<grid, 256, 0, st>>>(a, step_lo[s], (int)rows);
        }
        count_launch();
    }
    RLA_CUDA_CHECK(cudaGetLastError());
    return RLA_OK;
}
