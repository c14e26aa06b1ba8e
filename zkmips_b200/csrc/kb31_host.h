// kb31_host.h -- host-side KoalaBear scalars for launch parameters (generators, coset shifts, 1/n).
// Same representation as the device code: canonical Montgomery residues, R = 2^32
// (reference constants: crates/core/machine/include/kb31_t.hpp:27-34).  Only O(log n) scalar work
// happens here; every per-element operation of the proving path runs in the CUDA kernels.
#pragma once
#include <cstdint>

namespace kbh {

constexpr uint32_t P = 0x7f000001u;
constexpr uint32_t MU = 0x81000001u;
constexpr uint32_t ONE = 0x01fffffeu;
constexpr uint32_t RR = 0x17f7efe4u;
constexpr uint32_t GEN = 0x05fffffau;  // 3
constexpr unsigned TWO_ADICITY = 24;

constexpr uint32_t reduce64(uint64_t t) {
  uint32_t m = (uint32_t)t * MU;
  uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
  uint32_t r = (uint32_t)(t >> 32) - u;
  return (uint32_t)(t >> 32) < u ? r + P : r;
}
constexpr uint32_t mul(uint32_t a, uint32_t b) { return reduce64((uint64_t)a * b); }
constexpr uint32_t add(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P ? s - P : s; }
constexpr uint32_t sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }
constexpr uint32_t neg(uint32_t a) { return a ? P - a : 0; }
constexpr uint32_t pow(uint32_t a, uint64_t e) {
  uint32_t r = ONE;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
constexpr uint32_t inv(uint32_t a) { return pow(a, (uint64_t)P - 2); }
constexpr uint32_t to_monty(uint32_t c) { return mul(c % P, RR); }
constexpr uint32_t from_monty(uint32_t m) { return reduce64((uint64_t)m); }
// generator of the subgroup of order 2^bits: (3^127)^(2^(24-bits))   
constexpr uint32_t two_adic_generator(unsigned bits) {
  uint32_t g = pow(GEN, 127);
  for (unsigned i = bits; i < TWO_ADICITY; i++) g = mul(g, g);
  return g;
}
constexpr uint32_t bitrev(uint32_t x, unsigned bits) {
  uint32_t r = 0;
  for (unsigned i = 0; i < bits; i++) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}
constexpr unsigned log2_exact(uint64_t n) {
  unsigned k = 0;
  while ((1ull << k) < n) k++;
  return k;
}

}  // namespace kbh
