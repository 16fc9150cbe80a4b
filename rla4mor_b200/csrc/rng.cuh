// rng.cuh -- counter-based generator of the virtual embedding matrix Theta.
//
// Theta[row, col] (row = sketch index, col = index in the vector dimension) is a pure
// function of (seed, kind, row, col): Philox4x32-10 keyed by the 64-bit seed, counter
// (col / 4, row, col / 2^34, kind), whose four 32-bit outputs become the four entries
// Theta[row, 4q .. 4q+3].  The GEMM kernel (gemm.cu) generates exactly the entries its
// DMMA fragments need, in registers; rla_theta_materialize_f64 exports the same values
// (same device function, hence bit-identical) for parity checks against the oracle.
//
//   kind 0: standard normal by Box-Muller in FP32 (MUFU lg2 / sqrt / sin / cos), widened to FP64
//   kind 1: Rademacher +-1 (one bit per entry)
//   kind 2: the normals of kind 0 rounded to TF32 (10 explicit mantissa bits, round to nearest):
//           every entry is exactly representable as an operand of the generation-5 tensor cores
//           (tcgen05.mma kind::tf32), which the float32 path of gemm32.cu relies on; FP64 blocks
//           sketched with kind 2 use the same rounded values, so Theta does not depend on the
//           dtype of the block
// The 1/sqrt(k) scale of the reference (rla/embeddings.py:269) is applied by the caller.
#pragma once
#include <stdint.h>

namespace rla {

struct PhiloxOut { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ PhiloxOut philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                            uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = mulhi32(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = mulhi32(M1, c2), l1 = M1 * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += W0; k1 += W1;
    }
    return PhiloxOut{c0, c1, c2, c3};
}

// two standard normals from two 32-bit words (FP32 Box-Muller on the special-function unit:
// MUFU.LG2 / MUFU.SQRT / MUFU.COS / MUFU.SIN, 12 instructions per pair; the IEEE sqrtf and the
// denormal-safe __log2f cost 12 more, and every instruction the generator issues is taken
// from the DMMA warps of its sub-partition).  Deterministic on sm_100; the approximations are
// hardware-defined, so (kind, n, k, seed) determines Theta bit for bit on this architecture only.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (a + 0.5) 2^-32, in (0, 1]
    const float ang = (float)(int32_t)b * 1.4629180792671596e-9f;                         // pi 2^-31 b, [-pi, pi)
    float l, r, c, s;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
    const float t = -1.3862943611198906f * l;                                             // -2 ln u1 >= 0
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
    n0 = r * c;
    n1 = r * s;
}

// round to nearest TF32, ties away from zero -- what cvt.rna.tf32.f32 does for finite values -- as
// two integer instructions (half an ulp of the 10-bit mantissa added to the magnitude, 13 bits
// cleared): the conversion would run on the XU pipe, which the generator already loads with MUFU
__device__ __forceinline__ float round_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// the four entries Theta[row, 4*q .. 4*q+3]  (q = col / 4) as floats, unscaled
template <int KIND>
__device__ __forceinline__ void theta4f(uint64_t seed, uint32_t row, uint64_t q, float (&out)[4]) {
    if (KIND == 0 || KIND == 2) {
        const PhiloxOut p = philox4x32_10((uint32_t)q, row, (uint32_t)(q >> 32), 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
        box_muller(p.x, p.y, out[0], out[1]);
        box_muller(p.z, p.w, out[2], out[3]);
        if (KIND == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = round_tf32(out[j]);
        }
    } else {
        // 128 sign bits per Philox block: block index q / 32, bits 4*(q%32) .. +3
        const uint64_t blk = q >> 5;
        const PhiloxOut p = philox4x32_10((uint32_t)blk, row, (uint32_t)(blk >> 32), 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t sel = (uint32_t)(q & 31);
        const uint32_t wsel = sel >> 3;
        const uint32_t word = wsel == 0 ? p.x : (wsel == 1 ? p.y : (wsel == 2 ? p.z : p.w));
        const uint32_t bits = word >> (4 * (sel & 7));
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = ((bits >> j) & 1u) ? -1.0f : 1.0f;
    }
}

// ... and as doubles (the FP64 kernels)
template <int KIND>
__device__ __forceinline__ void theta4(uint64_t seed, uint32_t row, uint64_t q, double (&out)[4]) {
    float f[4];
    theta4f<KIND>(seed, row, q, f);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = (double)f[j];
}

}  // namespace rla
