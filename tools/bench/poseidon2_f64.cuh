// poseidon2_f64.cuh -- the same Poseidon2 width-16 KoalaBear permutation as poseidon2.cuh, computed EXACTLY on the
// FP64 pipe: one permutation per thread, the 16-word state in 16 doubles holding (possibly unreduced, possibly
// negative) integers.
//
// Why: the integer permutation (poseidon2.cuh) is bound by the fma (IMAD*) and alu pipes of a sub-partition and leaves
// a third of the issue slots and the whole FP64 pipe idle.  On B200 the FP64 pipe issues a warp-wide DFMA / DADD / DMUL
// every 2 cycles -- the same rate as IMAD -- and runs side by side with the two integer pipes (profiles/
// r2_pipebench.txt: DFMA + IMAD + IADD3 = 3.13 cycles for the three).  EXPERIMENT, NOT USED BY THE PRODUCT: giving some
// warps of the sponge this version and the others the integer one does not raise throughput (profiles/
// r2_p2bench_fp64_mix.txt: 47.7 clk/perm/SM integer only, 54.6 FP64 only at 95 % of the FP64 pipe, 47.2-48.2 for every
// mix; in the mix no pipe is above 54 % and issue slots stay at 66 %: the sub-partition's dispatch / operand delivery
// is the shared limit, not a pipe).  Kept because it is bit-exact and documents the ceiling.
//
// Exactness (every operation below is exact, so results are bit-identical to the integer permutation):
//   * values are integers of magnitude < 2^52 in doubles; additions and multiplications by small constants are exact;
//   * a modular product a*b (|a*b| < 2^82): hi = RN(a*b), lo = fma(a, b, -hi) is the exact rounding error (an integer),
//     q = rint(hi / p) through the 1.5 * 2^52 magic constant (|hi / p| < 2^51), r = fma(-q, p, hi) is exact because it
//     is an integer below 2^32; result r + lo is congruent to a*b and lies within +-(p/2 + 2^-13 p + |lo|);
//   * the linear layers need NO reductions: |S-box output| <= 2^30.01, one external layer grows it 35x (2^35.13), the
//     initial layer of the NEXT permutation of a sponge another <= 35x (2^40.3), and an S-box takes inputs up to 2^40.4
//     (|t*t| <= 2^80.8: q < 2^50; |x2*t| <= 2^70.9);
//   * the state is kept in CANONICAL form (not Montgomery): the S-box is then a plain cube and only the words that
//     enter or leave the sponge are converted (one modular product by R^-1 or R each).
// Reference algorithm: SURVEY A.3 (crates/primitives/src/lib.rs:563-1121, crates/recursion/core/include/
// poseidon2.hpp:21-71, poseidon2_constants.hpp:1083-1100).
#pragma once
#include "../../zkmips_b200/csrc/kb31.cuh"
#include "../../zkmips_b200/csrc/kb31_host.h"
#include "../../include/zk_poseidon2_rc.h"

namespace p2d {

constexpr double PD = 2130706433.0;               // p
constexpr double PINV = 1.0 / 2130706433.0;        // RN(1/p)
constexpr double MAGIC = 6755399441055744.0;       // 1.5 * 2^52: x + MAGIC - MAGIC = rint(x) for |x| < 2^51
constexpr double TWO52 = 4503599627370496.0;
// centred representatives of R^-1 and R (R = 2^32) mod p: entering / leaving Montgomery form
constexpr int64_t centred(uint32_t c) { return c > kbh::P / 2 ? (int64_t)c - (int64_t)kbh::P : (int64_t)c; }
constexpr double RINV_C = (double)centred(kbh::from_monty(1u));            // 2^-32 mod p
constexpr double R_C = (double)centred(kbh::ONE);                          // 2^32 mod p

struct RcTables {
  double ext[8][16];
  double in[13];
};
constexpr RcTables make_rc() {
  constexpr uint32_t e[8][16] = ZK_P2_EXT_RC_MONTY;
  constexpr uint32_t n[13] = ZK_P2_INT_RC_MONTY;
  RcTables t{};
  for (int r = 0; r < 8; r++)
    for (int i = 0; i < 16; i++) t.ext[r][i] = (double)kbh::from_monty(e[r][i]);
  for (int r = 0; r < 13; r++) t.in[r] = (double)kbh::from_monty(n[r]);
  return t;
}
static __constant__ RcTables RC = make_rc();

__device__ __forceinline__ double rint_div_p(double x) { return __dadd_rn(__fma_rn(x, PINV, MAGIC), -MAGIC); }
// integer x, |x| < 2^82 -> congruent value within +-(p/2 + p * 2^-13)
__device__ __forceinline__ double red(double x) { return __fma_rn(-rint_div_p(x), PD, x); }
// a * b mod p for integers with |a * b| < 2^82
__device__ __forceinline__ double mulmod(double a, double b) {
  double hi = __dmul_rn(a, b);
  double lo = __fma_rn(a, b, -hi);
  double r = __fma_rn(-rint_div_p(hi), PD, hi);
  return __dadd_rn(r, lo);
}
__device__ __forceinline__ double sbox(double x, double rc) {
  double t = __dadd_rn(x, rc);
  return mulmod(mulmod(t, t), t);
}
// x * 2^-K mod p: x = xr * 2^K + xl with |xl| <= 2^(K-1);  2^-K = -(p-1)/2^K (mod p), so x / 2^K = xr - xl * ((p-1) >> K)
template <int K>
__device__ __forceinline__ double div2k(double x) {
  constexpr double inv = 1.0 / (double)(1u << K), pw = (double)(1u << K), c = (double)((kbh::P - 1) >> K);
  double xr = __dadd_rn(__fma_rn(x, inv, MAGIC), -MAGIC);
  double xl = __fma_rn(xr, -pw, x);
  return __fma_rn(xl, -c, xr);
}

__device__ __forceinline__ void m4(double& x0, double& x1, double& x2, double& x3) {
  double t01 = __dadd_rn(x0, x1), t23 = __dadd_rn(x2, x3);
  double t0123 = __dadd_rn(t01, t23);
  double t01123 = __dadd_rn(t0123, x1), t01233 = __dadd_rn(t0123, x3);
  double n3 = __fma_rn(2.0, x0, t01233), n1 = __fma_rn(2.0, x2, t01123);
  double n0 = __dadd_rn(t01123, t01), n2 = __dadd_rn(t01233, t23);
  x0 = n0; x1 = n1; x2 = n2; x3 = n3;
}
__device__ __forceinline__ void external_layer(double (&s)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) m4(s[i], s[i + 1], s[i + 2], s[i + 3]);
  double sums[4];
#pragma unroll
  for (int k = 0; k < 4; k++) sums[k] = __dadd_rn(__dadd_rn(s[k], s[4 + k]), __dadd_rn(s[8 + k], s[12 + k]));
#pragma unroll
  for (int j = 0; j < 16; j++) s[j] = __dadd_rn(s[j], sums[j & 3]);
}
// s[i] = V[i] * s[i] + sum, V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/8, 2^-24, -2^-8, -1/8, -1/16, -2^-24].
// The sum is reduced (so that the lanes with |V| = 1 and the 2^-k lanes stay small); the lanes with |V| >= 2 grow by
// that factor per round and are reduced every second round (`shrink`): nothing exceeds 2^38.
__device__ __forceinline__ void internal_layer(double (&s)[16], bool shrink) {
  if (shrink) {
    s[2] = red(s[2]); s[4] = red(s[4]); s[5] = red(s[5]); s[7] = red(s[7]); s[8] = red(s[8]);
  }
  double a0 = __dadd_rn(s[0], s[1]), a1 = __dadd_rn(s[2], s[3]), a2 = __dadd_rn(s[4], s[5]), a3 = __dadd_rn(s[6], s[7]);
  double a4 = __dadd_rn(s[8], s[9]), a5 = __dadd_rn(s[10], s[11]), a6 = __dadd_rn(s[12], s[13]), a7 = __dadd_rn(s[14], s[15]);
  double sum = red(__dadd_rn(__dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3)), __dadd_rn(__dadd_rn(a4, a5), __dadd_rn(a6, a7))));
  s[0] = __fma_rn(-2.0, s[0], sum);
  s[1] = __dadd_rn(s[1], sum);
  s[2] = __fma_rn(2.0, s[2], sum);
  s[3] = __dadd_rn(sum, div2k<1>(s[3]));
  s[4] = __fma_rn(3.0, s[4], sum);
  s[5] = __fma_rn(4.0, s[5], sum);
  s[6] = __dadd_rn(sum, -div2k<1>(s[6]));
  s[7] = __fma_rn(-3.0, s[7], sum);
  s[8] = __fma_rn(-4.0, s[8], sum);
  s[9] = __dadd_rn(sum, div2k<8>(s[9]));
  s[10] = __dadd_rn(sum, div2k<3>(s[10]));
  s[11] = __dadd_rn(sum, div2k<24>(s[11]));
  s[12] = __dadd_rn(sum, -div2k<8>(s[12]));
  s[13] = __dadd_rn(sum, -div2k<3>(s[13]));
  s[14] = __dadd_rn(sum, -div2k<4>(s[14]));
  s[15] = __dadd_rn(sum, -div2k<24>(s[15]));
}

// In: canonical integers, |s[i]| <= 2^35.2 (e.g. the unreduced output of a previous call, or fresh words < p).
// Out: congruent to the permutation's output, |s[i]| <= 2^35.13.  Loops rolled like the integer version.
__device__ __forceinline__ void permute(double (&s)[16]) {
  external_layer(s);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = sbox(s[i], RC.ext[half * 4 + r][i]);
      external_layer(s);
    }
    if (half == 0) {
#pragma unroll 1
      for (int r = 0; r < 13; r++) {
        s[0] = sbox(s[0], RC.in[r]);
        internal_layer(s, (r & 1) != 0);
      }
    }
  }
}

// ---- words entering / leaving the sponge ------------------------------------------------------------------------
__device__ __forceinline__ double u32_to_double(uint32_t x) {
  return __dadd_rn(__hiloint2double(0x43300000, (int)x), -TWO52);  // bit pattern of 2^52 + x, minus 2^52: exact
}
// Montgomery word (< p) -> canonical value (centred), and back to a canonical Montgomery word in [0, p)
__device__ __forceinline__ double from_monty_word(uint32_t w) { return mulmod(u32_to_double(w), RINV_C); }
__device__ __forceinline__ uint32_t to_monty_word(double x) {
  double r = mulmod(red(x), R_C);                                // within (-p, p)
  uint32_t v = (uint32_t)__double2loint(__dadd_rn(r, MAGIC));    // two's complement of the integer r
  return min(v, v + kb::P);
}

}  // namespace p2d
