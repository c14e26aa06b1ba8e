// kb31.cuh -- KoalaBear (p = 2^31 - 2^24 + 1) device arithmetic for sm_100a.
//
// Values are canonical Montgomery residues (R = 2^32), bit-compatible with a Rust `KoalaBear` and with the
// reference's native class (crates/core/machine/include/kb31_t.hpp:27-34; host monty_reduce :495-503).
// The reference's device branch (kb31_t.hpp:208-222) uses mad.lo.cc/madc.hi with M = -p^{-1}; here the
// subtractive form with MU = p^{-1} is used because ptxas fuses the final correction into one
// VIADDMNMX.U32 on sm_100a:  IMAD.WIDE.U32, IMAD, IMAD.HI.U32 (fma pipe) + IADD3, VIADDMNMX (alu pipe).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kb {

constexpr uint32_t P = 0x7f000001u;
constexpr uint32_t MU = 0x81000001u;   // p^{-1} mod 2^32
constexpr uint32_t ONE = 0x01fffffeu;  // R mod p
constexpr uint32_t RR = 0x17f7efe4u;   // R^2 mod p
constexpr uint32_t GEN = 0x05fffffau;  // 3 in Montgomery form (3 * ONE mod p)
constexpr uint32_t HALF = 0x00ffffffu; // 1/2 in Montgomery form (ONE / 2)

// [0, 2p) -> [0, p)
__device__ __forceinline__ uint32_t reduce(uint32_t s) { return min(s, s - P); }
__device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return reduce(a + b); }
__device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
  uint32_t d = a - b;
  return min(d, d + P);
}
__device__ __forceinline__ uint32_t neg(uint32_t a) { return a ? P - a : 0u; }
__device__ __forceinline__ uint32_t dbl(uint32_t a) { return reduce(a << 1); }
__device__ __forceinline__ uint32_t halve(uint32_t a) { return (a >> 1) + ((a & 1u) ? (P + 1) / 2 : 0u); }

// Montgomery product; inputs need a * b < p * 2^32 (one operand may be any u32 if the other is < p).
__device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * MU;
  uint32_t u = __umulhi(m, P);
  uint32_t r = (uint32_t)(t >> 32) - u;
  return min(r, r + P);
}
// Montgomery product without the final correction: result in (0, 2p) when a * b < p * p.
__device__ __forceinline__ uint32_t mul_lazy(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * MU;
  uint32_t u = __umulhi(m, P);
  return (uint32_t)(t >> 32) - u + P;
}
__device__ __forceinline__ uint32_t sqr(uint32_t a) { return mul(a, a); }
__device__ __forceinline__ uint32_t cube(uint32_t a) { return mul(mul_lazy(a, a), a); }

__device__ __forceinline__ uint32_t pow(uint32_t a, uint64_t e) {
  uint32_t r = ONE;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ uint32_t inv(uint32_t a) { return pow(a, (uint64_t)P - 2); }
__device__ __forceinline__ uint32_t to_monty(uint32_t canon) { return mul(canon, RR); }
__device__ __forceinline__ uint32_t from_monty(uint32_t m) { return mul(m, 1u); }

// ---- degree-4 extension F_p[X]/(X^4 - 3)  (crates/stark/src/air/extension.rs:53-75) ----
struct Ext {
  uint32_t c[4];
};
__device__ __forceinline__ Ext ext_zero() { return Ext{{0u, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_one() { return Ext{{ONE, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_from_base(uint32_t a) { return Ext{{a, 0u, 0u, 0u}}; }
__device__ __forceinline__ Ext ext_add(Ext a, Ext b) {
  return Ext{{add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2]), add(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Ext ext_sub(Ext a, Ext b) {
  return Ext{{sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2]), sub(a.c[3], b.c[3])}};
}
__device__ __forceinline__ Ext ext_neg(Ext a) { return Ext{{neg(a.c[0]), neg(a.c[1]), neg(a.c[2]), neg(a.c[3])}}; }
__device__ __forceinline__ Ext ext_mul_base(Ext a, uint32_t b) {
  return Ext{{mul(a.c[0], b), mul(a.c[1], b), mul(a.c[2], b), mul(a.c[3], b)}};
}
__device__ __forceinline__ Ext ext_add_base(Ext a, uint32_t b) {
  a.c[0] = add(a.c[0], b);
  return a;
}
__device__ __forceinline__ Ext ext_sub_base(Ext a, uint32_t b) {
  a.c[0] = sub(a.c[0], b);
  return a;
}
// 64-bit accumulation of partial products, one Montgomery reduction per output coefficient.
// Each product is < p^2 < 2^62; sums of up to 4 products times (1 or 3) need care: reduce the
// high-degree half first.
__device__ __forceinline__ uint32_t mont_reduce64(uint64_t t) {  // t < p * 2^32
  uint32_t m = (uint32_t)t * MU;
  uint32_t u = __umulhi(m, P);
  uint32_t r = (uint32_t)(t >> 32) - u;
  return min(r, r + P);
}
__device__ __forceinline__ Ext ext_mul(Ext a, Ext b) {
  // c_k = sum_{i+j=k} a_i b_j + 3 * sum_{i+j=k+4} a_i b_j
  // each a_i b_j < p^2 ~ 2^62; two of them fit in u64 (< 2^63), so pairs are reduced separately.
  auto pr = [](uint32_t x, uint32_t y) { return (uint64_t)x * y; };
  uint32_t h0 = add(mont_reduce64(pr(a.c[1], b.c[3]) + pr(a.c[3], b.c[1])), mul(a.c[2], b.c[2]));  // deg 4
  uint32_t h1 = mont_reduce64(pr(a.c[2], b.c[3]) + pr(a.c[3], b.c[2]));                            // deg 5
  uint32_t h2 = mul(a.c[3], b.c[3]);                                                               // deg 6
  uint32_t l0 = mul(a.c[0], b.c[0]);
  uint32_t l1 = mont_reduce64(pr(a.c[0], b.c[1]) + pr(a.c[1], b.c[0]));
  uint32_t l2 = add(mont_reduce64(pr(a.c[0], b.c[2]) + pr(a.c[2], b.c[0])), mul(a.c[1], b.c[1]));
  uint32_t l3 = add(mont_reduce64(pr(a.c[0], b.c[3]) + pr(a.c[3], b.c[0])),
                    mont_reduce64(pr(a.c[1], b.c[2]) + pr(a.c[2], b.c[1])));
  auto times3 = [](uint32_t x) { return add(dbl(x), x); };
  return Ext{{add(l0, times3(h0)), add(l1, times3(h1)), add(l2, times3(h2)), l3}};
}
__device__ __forceinline__ Ext ext_sqr(Ext a) { return ext_mul(a, a); }
__device__ __forceinline__ Ext ext_pow(Ext a, uint64_t e) {
  Ext r = ext_one();
  while (e) {
    if (e & 1) r = ext_mul(r, a);
    a = ext_sqr(a);
    e >>= 1;
  }
  return r;
}
// inverse through the norm to F_p[Y]/(Y^2 - 3), Y = X^2 (unique field element; any method is exact)
__device__ __forceinline__ Ext ext_inv(Ext a) {
  auto times3 = [](uint32_t x) { return add(dbl(x), x); };
  uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
  uint32_t n0 = sub(add(sqr(a0), times3(sqr(a2))), times3(dbl(mul(a1, a3))));
  uint32_t n1 = sub(dbl(mul(a0, a2)), add(sqr(a1), times3(sqr(a3))));
  uint32_t d = inv(sub(sqr(n0), times3(sqr(n1))));
  uint32_t i0 = mul(n0, d), i1 = neg(mul(n1, d));
  uint32_t A0 = add(mul(a0, i0), times3(mul(a2, i1)));
  uint32_t A1 = add(mul(a0, i1), mul(a2, i0));
  uint32_t B0 = add(mul(a1, i0), times3(mul(a3, i1)));
  uint32_t B1 = add(mul(a1, i1), mul(a3, i0));
  return Ext{{A0, neg(B0), A1, neg(B1)}};
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

}  // namespace kb
