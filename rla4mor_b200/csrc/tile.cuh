// tile.cuh -- the 4096-element register/shared-memory Walsh-Hadamard tile shared by
// srht.cu (fused SRHT) and fwht.cu (full transform).
#pragma once
#include "common.cuh"

namespace rla {

constexpr int TILE_LOG2 = 12;
constexpr int TILE = 1 << TILE_LOG2;     // elements per tile
constexpr int GROUP = 64;                // threads per tile
constexpr int CTA = 128;                 // two tiles per iteration
constexpr int MAX_NSLOT = 32;            // sample accumulators per thread
constexpr int MAX_SLOTS = MAX_NSLOT * CTA;

// Position of tile element e in the shared-memory tile buffer.  The tile is a
// 64 x 64 matrix (A = bits 1..6 of e, B = bit 0 and bits 7..11); round 1 holds one
// A per thread, round 2 one B per thread; XOR swizzle keeps both sides bank-conflict
// free (mask 15 for 8-byte words: 16 lanes per wavefront; 31 for 4-byte words).
__host__ __device__ __forceinline__ int tile_pos(int e, int mask) {
    int A = (e >> 1) & 63;
    int B = (e & 1) | ((e >> 7) << 1);
    return B * 64 + (A ^ (B & mask));
}

// Final position (and sign twist) of tile element e in the warp-specialised SRHT kernel.
// Round 2 of that kernel reads its row of the exchange buffer with 16-byte loads in POSITION
// order instead of logical order; the row holds logical index A at position A ^ (B & mask),
// and the thread reads 16-byte chunk j ^ (B & 7) into register chunk j (bank-conflict free),
// so register p holds logical A = p ^ c with c = (B & mask) ^ ((B & 7) << log2 E)
// (E = elements per 16 bytes).  Butterflies are invariant under XOR relabelling up to signs:
// H P_c = D_c H with D_c = diag((-1)^popcount(s & c)), so the registers end up holding
// z[s] = (-1)^popcount(s & c) y[s] at the natural output index s.  z is written back chunk by
// chunk to the positions it was read from; the sign is folded into the sample descriptor.
__host__ __device__ __forceinline__ int tile_pos_ws(int e, int elem_bytes, int *negate) {
    const int mask = elem_bytes == 8 ? 15 : 31, le = elem_bytes == 8 ? 1 : 2, E = 1 << le;
    const int A = (e >> 1) & 63;                       // logical output index s of round 2
    const int B = (e & 1) | ((e >> 7) << 1);           // round-2 thread (row of the buffer)
    const int c = (B & mask) ^ ((B & 7) << le);
    int v = A & c, par = 0;
    while (v) { par ^= v & 1; v >>= 1; }
    *negate = par;
    return B * 64 + ((((A >> le) ^ (B & 7)) << le) | (A & (E - 1)));
}

template <typename T> struct Elem;
template <> struct Elem<double> {
    static constexpr int MASK = 15;
    using Chunk = double2;                             // 16 bytes of shared memory
    __device__ static __forceinline__ void unpack(const double2 &c, double *v) { v[0] = c.x; v[1] = c.y; }
    __device__ static __forceinline__ double2 pack(const double *v) { return make_double2(v[0], v[1]); }
    __device__ static __forceinline__ void load2(const double *p, double &a, double &b) {
        double2 v = ldg_stream_f64x2(p); a = v.x; b = v.y;
    }
    __device__ static __forceinline__ double load1(const double *p) { return ldg_stream_f64(p); }
};
template <> struct Elem<float> {
    static constexpr int MASK = 31;
    using Chunk = float4;
    __device__ static __forceinline__ void unpack(const float4 &c, float *v) { v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w; }
    __device__ static __forceinline__ float4 pack(const float *v) { return make_float4(v[0], v[1], v[2], v[3]); }
    __device__ static __forceinline__ void load2(const float *p, float &a, float &b) {
        float2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
        a = v.x; b = v.y;
    }
    __device__ static __forceinline__ float load1(const float *p) { return ldg_stream_f32(p); }
};

// 6 radix-2 stages over the 64 registers of one thread
template <typename T>
__device__ __forceinline__ void butterflies64(T (&v)[64]) {
#pragma unroll
    for (int b = 0; b < 6; ++b) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            if ((i & (1 << b)) == 0) {
                T p = v[i], q = v[i | (1 << b)];
                v[i] = p + q;
                v[i | (1 << b)] = p - q;
            }
        }
    }
}

// Rademacher sign flip (srht.py:165) fused with the first butterfly stage (register bit 0):
// the two sign masks of a pair are formed right where they are used, which keeps the
// register pressure of the flip at a handful of temporaries.  Then stages 1..5.
template <typename T>
__device__ __forceinline__ void flip_and_butterflies64(T (&v)[64], uint64_t sw) {
    const uint32_t lo = (uint32_t)sw, hi = (uint32_t)(sw >> 32);
#pragma unroll
    for (int h = 0; h < 32; ++h) {
        const uint32_t w = h < 16 ? lo : hi;
        const int b = (2 * h) & 31;
        const T x = xor_sign(v[2 * h], w << (31 - b));
        const T y = xor_sign(v[2 * h + 1], w << (30 - b));
        v[2 * h] = x + y;
        v[2 * h + 1] = x - y;
    }
#pragma unroll
    for (int b = 1; b < 6; ++b) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            if ((i & (1 << b)) == 0) {
                T p = v[i], q = v[i | (1 << b)];
                v[i] = p + q;
                v[i | (1 << b)] = p - q;
            }
        }
    }
}

// 64-thread barrier of one tile group (ids 1 and 2; id 0 is __syncthreads)
__device__ __forceinline__ void group_barrier(int grp) {
    if (grp == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else asm volatile("bar.sync 2, 64;" ::: "memory");
}

// Fast load of a full, 16-byte aligned tile straight into the round-1 register layout:
// register rho = 2*h + l  <->  tile element e = 128*h + 2*tg + l.
template <typename T>
__device__ __forceinline__ void load_tile_fast(T (&v)[64], const T *__restrict__ tilep, int tg) {
#pragma unroll
    for (int h = 0; h < 32; ++h) Elem<T>::load2(tilep + 128 * h + 2 * tg, v[2 * h], v[2 * h + 1]);
}

// Slow load (ragged last tile, or rows that are not 16-byte aligned): stage the tile in
// natural order through the group's shared-memory buffer; elements at or beyond n read
// as zero (the virtual zero padding of srht.py:167).
template <typename T>
__device__ __forceinline__ void load_tile_slow(T (&v)[64], const T *__restrict__ rowp, int64_t n,
                                               int64_t jh, int tg, T *__restrict__ buf, int grp) {
    const int64_t j0 = jh * TILE;
#pragma unroll 4
    for (int e = tg; e < TILE; e += GROUP) buf[e] = (j0 + e < n) ? Elem<T>::load1(rowp + j0 + e) : T(0);
    group_barrier(grp);
#pragma unroll
    for (int h = 0; h < 32; ++h) {
        v[2 * h] = buf[128 * h + 2 * tg];
        v[2 * h + 1] = buf[128 * h + 2 * tg + 1];
    }
    group_barrier(grp);
}

// Rademacher sign flip (srht.py:165) from the packed sign word of this thread.
template <typename T>
__device__ __forceinline__ void flip_signs(T (&v)[64], uint64_t sw) {
    const uint32_t lo = (uint32_t)sw, hi = (uint32_t)(sw >> 32);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        v[r] = xor_sign(v[r], lo << (31 - r));
        v[r + 32] = xor_sign(v[r + 32], hi << (31 - r));
    }
}

// Full 12-stage transform of one tile by a 64-thread group; result left in `buf`
// at tile_pos().  Caller must synchronise before other threads read buf.
template <typename T>
__device__ __forceinline__ void tile_fwht(T (&v)[64], T *__restrict__ buf, int tg, int grp) {
    constexpr int M = Elem<T>::MASK;
    // opaque copy: keeps the compiler from hoisting the 64 swizzled addresses out of the
    // tile loop (they would not fit in registers and end up in local memory)
    asm volatile("" : "+r"(tg));
    butterflies64(v);  // bits 0, 7..11
#pragma unroll
    for (int r = 0; r < 64; ++r) buf[r * 64 + (tg ^ (r & M))] = v[r];
    group_barrier(grp);
    // round 2: thread tg holds B = tg, registers run over A
#pragma unroll
    for (int r = 0; r < 64; ++r) v[r] = buf[tg * 64 + (r ^ (tg & M))];
    butterflies64(v);  // bits 1..6
#pragma unroll
    for (int r = 0; r < 64; ++r) buf[tg * 64 + (r ^ (tg & M))] = v[r];
}

}  // namespace rla
