// philox.cuh -- counter-based Philox4x32-10 (Salmon et al., SC'11), hand-written for the E-step kernels.
//
// Replaces the reference's per-thread std::mt19937 streams (src/rng_utils.hpp:87-120, src/mcem.hpp:62-68).
// Stream definition (stable, documented in DESIGN.md):
//   key     = (seed, tag)            tag = (EM iteration << 8) | (stream_id << 3 ... ) see make_key()
//   counter = (sample l, GLOBAL observation index i, 128-bit block index b, candidate/evaluation id)
// so a draw depends only on (seed, iteration, i, l, b): independent of grid shape, chunking and of how the
// observations are sharded over GPUs.
#pragma once
#include "device_types.cuh"

// Philox4x32-10 is the production generator (10 rounds, the Random123 / cuRAND default).  The round count is a macro only
// so that profiling experiments can measure how much of a kernel's time the generator accounts for.
#ifndef MCCBN_PHILOX_ROUNDS
#define MCCBN_PHILOX_ROUNDS 10
#endif

namespace mccbn {

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ inline PhiloxKey make_key(uint32_t seed, uint32_t iter, uint32_t purpose) {
  PhiloxKey k;
  k.k0 = seed;
  k.k1 = (iter << 4) | (purpose & 0xF);
  return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < MCCBN_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;  // IMAD.WIDE.U32
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;  // LOP3
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;  // uniform across the grid: hoisted / folded by the compiler
    k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// N independent Philox blocks that differ only in counter word 2, advanced round by round in lock-step: the 2N
// multiplies of a round are independent, which hides the IMAD.WIDE -> LOP3 dependency latency of a single chain
// (used by the poset-specialised kernels, jit.cuh).
template <int N>
__device__ __forceinline__ void philox4x32_10_multi(const uint32_t c0, const uint32_t c1, const uint32_t c2_first,
                                                    const uint32_t c3, uint32_t k0, uint32_t k1, uint4 (&out)[N]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t a0[N], a1[N], a2[N], a3[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a0[i] = c0;
    a1[i] = c1;
    a2[i] = c2_first + (uint32_t)i;
    a3[i] = c3;
  }
#pragma unroll
  for (int r = 0; r < MCCBN_PHILOX_ROUNDS; ++r) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const uint64_t p0 = (uint64_t)M0 * a0[i];
      const uint64_t p1 = (uint64_t)M1 * a2[i];
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ a1[i] ^ k0;
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ a3[i] ^ k1;
      a1[i] = (uint32_t)p1;
      a3[i] = (uint32_t)p0;
      a0[i] = n0;
      a2[i] = n2;
    }
    k0 += W0;
    k1 += W1;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = make_uint4(a0[i], a1[i], a2[i], a3[i]);
}

// 32 random bits -> u in (0, 1] with 23 random mantissa bits: u = 2 - [1,2)
__device__ __forceinline__ float u01_open0(uint32_t bits) { return 2.0f - __uint_as_float((bits >> 9) | 0x3f800000u); }

__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // MUFU.LG2
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // MUFU.EX2
  return y;
}

// 2 x 32 bits -> double in (0, 1] with 53 random bits (fp64 validation mode)
__device__ __forceinline__ double u01_open0_d(uint32_t hi, uint32_t lo) {
  const uint64_t m = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);  // 53 bits
  return ((double)m + 1.0) * (1.0 / 9007199254740992.0);
}

}  // namespace mccbn
