// poseidon2.cuh -- Poseidon2 width-16 KoalaBear permutation, register resident, one permutation per thread.
//
// Algorithm: x^3 S-box, 4 + 4 external rounds, 13 internal rounds; round constants
// crates/primitives/src/lib.rs:563-1121 (selection :1107-1121); linear layers as restated by the
// reference at crates/recursion/core/include/poseidon2.hpp:21-71; internal diagonal
// V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/8, 1/2^24, -1/2^8, -1/8, -1/16, -1/2^24]
// (poseidon2_constants.hpp:1083-1100).  Sponge / compression: PaddingFreeSponge<16,8,8> and
// TruncatedPermutation<2,8,16> (crates/stark/src/kb31_poseidon2.rs:173-177; semantics
// crates/recursion/circuit/src/hash.rs:40-49,76-81).
//
// The state lives in 16 registers.  The ROUND loops are rolled (one copy of the external-round body, one of the
// internal-round body, ~760 instructions = 12 KB): the fully unrolled permutation is ~4500 instructions = 72 KB
// of straight-line code that every warp streams once per permutation, and that version was bound by instruction
// fetch, not by the integer pipes (ncu on B200: instruction-cache hit rate 57 %, 44 % of stall cycles
// "no instruction", 55.6 clk/perm/SM; rolled: hit rate 99.99 %, fmaheavy pipe 85 % busy, 47.8 clk/perm/SM --
// profiles/r1_p2bench_variants.txt).  Round constants come from the constant bank (uniform LDC per round).
// Within a round everything is unrolled over the 16 lanes.  Bound by the integer pipes (fma: IMAD*, alu:
// IADD3/VIADDMNMX; IMAD.WIDE and IMAD.HI occupy the fma pipe for 4 cycles, the rest for 2:
// profiles/r1_pipebench.txt), not by memory -- see DESIGN.md.
#pragma once
#include "kb31.cuh"
#include "../../include/zk_poseidon2_rc.h"

namespace p2 {

static __constant__ uint32_t EXT_RC[8][16] = ZK_P2_EXT_RC_MONTY;
static __constant__ uint32_t INT_RC[13] = ZK_P2_INT_RC_MONTY;

// S-box with the round constant folded in: (x + rc)^3.  t = x + rc - p lies in [-p, p) and is squared as a signed
// number, so the modular add needs no correction; x2 = t*t/R stays uncorrected in (-p, p/2); the second product
// uses the signed Montgomery reduction (signed m, signed mulhi), whose result lies in (-p, p): one correction
// for the whole S-box -- 10 instructions instead of 11 (IADD3, 2x {IMAD.WIDE, IMAD, IMAD.HI, IADD3}, VIADDMNMX).
__device__ __forceinline__ uint32_t sbox(uint32_t x, uint32_t rc) {
  int32_t t = (int32_t)(x + (rc - kb::P));
  int64_t T = (int64_t)t * t;  // < p^2
  uint32_t m = (uint32_t)T * kb::MU;
  int32_t x2 = (int32_t)((uint64_t)T >> 32) - (int32_t)__umulhi(m, kb::P);
  int64_t T2 = (int64_t)x2 * t;  // |T2| < p^2
  int32_t m2 = (int32_t)((uint32_t)T2 * kb::MU);
  int32_t r = (int32_t)(T2 >> 32) - __mulhi(m2, (int32_t)kb::P);  // low words cancel exactly; r in (-p, p)
  return min((uint32_t)r, (uint32_t)r + kb::P);
}

// x * 2^-k mod p without a multiplication by a Montgomery constant: p = 127 * 2^24 + 1, so
// 2^-k = -(p-1)/2^k (k <= 24) and  x / 2^k = (x >> k) - (x mod 2^k) * ((p-1) >> k)  (mod p).
// (x mod 2^k) * ((p-1) >> k) <= p - 1 - ((p-1) >> k) < p, so one conditional correction suffices.
// Measured 4.5 % faster per permutation than the Montgomery-constant form (profiles/r1_p2bench_variants.txt).
template <int K>
__device__ __forceinline__ uint32_t div2k(uint32_t x) {
  constexpr uint32_t c = (kb::P - 1) >> K;
  uint32_t d = (x >> K) - (x & ((1u << K) - 1)) * c;
  return min(d, d + kb::P);
}

__device__ __forceinline__ void m4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
  uint32_t t01 = kb::add(x0, x1);
  uint32_t t23 = kb::add(x2, x3);
  uint32_t t0123 = kb::add(t01, t23);
  uint32_t t01123 = kb::add(t0123, x1);
  uint32_t t01233 = kb::add(t0123, x3);
  uint32_t n3 = kb::add(t01233, kb::dbl(x0));
  uint32_t n1 = kb::add(t01123, kb::dbl(x2));
  uint32_t n0 = kb::add(t01123, t01);
  uint32_t n2 = kb::add(t01233, t23);
  x0 = n0; x1 = n1; x2 = n2; x3 = n3;
}

__device__ __forceinline__ void external_layer(uint32_t (&s)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) m4(s[i], s[i + 1], s[i + 2], s[i + 3]);
  uint32_t sums[4];
#pragma unroll
  for (int k = 0; k < 4; k++) sums[k] = kb::add(kb::add(s[k], s[4 + k]), kb::add(s[8 + k], s[12 + k]));
#pragma unroll
  for (int j = 0; j < 16; j++) s[j] = kb::add(s[j], sums[j & 3]);
}

__device__ __forceinline__ void internal_layer(uint32_t (&s)[16]) {
  // tree sum of the 16 lanes
  uint32_t a0 = kb::add(s[0], s[1]), a1 = kb::add(s[2], s[3]), a2 = kb::add(s[4], s[5]), a3 = kb::add(s[6], s[7]);
  uint32_t a4 = kb::add(s[8], s[9]), a5 = kb::add(s[10], s[11]), a6 = kb::add(s[12], s[13]), a7 = kb::add(s[14], s[15]);
  uint32_t b0 = kb::add(a0, a1), b1 = kb::add(a2, a3), b2 = kb::add(a4, a5), b3 = kb::add(a6, a7);
  uint32_t sum = kb::add(kb::add(b0, b1), kb::add(b2, b3));
  uint32_t d;
  s[0] = kb::sub(sum, kb::dbl(s[0]));                    // -2
  s[1] = kb::add(sum, s[1]);                             //  1
  s[2] = kb::add(sum, kb::dbl(s[2]));                    //  2
  s[3] = kb::add(sum, kb::halve(s[3]));                  //  1/2
  d = kb::dbl(s[4]);  s[4] = kb::add(sum, kb::add(d, s[4]));   //  3
  s[5] = kb::add(sum, kb::dbl(kb::dbl(s[5])));           //  4
  s[6] = kb::sub(sum, kb::halve(s[6]));                  // -1/2
  d = kb::dbl(s[7]);  s[7] = kb::sub(sum, kb::add(d, s[7]));   // -3
  s[8] = kb::sub(sum, kb::dbl(kb::dbl(s[8])));           // -4
  s[9] = kb::add(sum, div2k<8>(s[9]));                   //  1/2^8
  s[10] = kb::add(sum, div2k<3>(s[10]));                 //  1/8
  s[11] = kb::add(sum, div2k<24>(s[11]));                //  1/2^24
  s[12] = kb::sub(sum, div2k<8>(s[12]));                 // -1/2^8
  s[13] = kb::sub(sum, div2k<3>(s[13]));                 // -1/8
  s[14] = kb::sub(sum, div2k<4>(s[14]));                 // -1/16
  s[15] = kb::sub(sum, div2k<24>(s[15]));                // -1/2^24
}

__device__ __forceinline__ void permute(uint32_t (&s)[16]) {
  external_layer(s);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = sbox(s[i], EXT_RC[half * 4 + r][i]);
      external_layer(s);
    }
    if (half == 0) {
#pragma unroll 1
      for (int r = 0; r < 13; r++) {
        s[0] = sbox(s[0], INT_RC[r]);
        internal_layer(s);
      }
    }
  }
}

// TruncatedPermutation<2,8,16>: digest = permute(l || r)[0..8]
__device__ __forceinline__ void compress(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { s[i] = l[i]; s[8 + i] = r[i]; }
  permute(s);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = s[i];
}

}  // namespace p2
