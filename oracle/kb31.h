/* oracle/kb31.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * KoalaBear base field F_p (p = 2^31 - 2^24 + 1) in Montgomery form (R = 2^32) and its degree-4
 * binomial extension F_p[X]/(X^4 - 3).  Plain-C restatement of the reference semantics:
 *   - Montgomery constants and monty_reduce: crates/core/machine/include/kb31_t.hpp:27-34,495-503
 *   - add / sub / mul / reciprocal (x^(p-2)): kb31_t.hpp:522-598
 *   - extension multiplication with W = 3:   crates/stark/src/air/extension.rs:53-75
 *     (same W at crates/recursion/gnark-ffi/go/zkm/koalabear/koalabear.go:247-268)
 *   - two-adic generator chain g_k = (3^127)^(2^(24-k)): SURVEY.md A.1 (Plonky3 monty-31, pinned fork
 *     ProjectZKM/Plonky3@faa24ca, absent from /root/reference; constant 0x6ac49f88 = 3^127 is checked
 *     by tests/test_oracle_field.py).
 * Values are canonical Montgomery residues (< p), so equality of field elements is equality of words.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this directory.
 */
#ifndef ORACLE_KB31_H
#define ORACLE_KB31_H
#include <stdint.h>
#include <stddef.h>

#define KB_P 0x7f000001u
#define KB_MU 0x81000001u /* p^{-1} mod 2^32, kb31_t.hpp:33 */
#define KB_ONE 0x01fffffeu /* R mod p, kb31_t.hpp:29 */
#define KB_RR 0x17f7efe4u  /* R^2 mod p, kb31_t.hpp:32 */
#define KB_TWO_ADICITY 24

typedef uint32_t kb_t;                 /* Montgomery residue */
typedef struct { kb_t c[4]; } kb4_t;   /* c[0] + c[1] X + c[2] X^2 + c[3] X^3, X^4 = 3 */

/* kb31_t.hpp:495-503 */
static inline kb_t kb_monty_reduce(uint64_t x) {
  uint64_t t = (x * (uint64_t)KB_MU) & 0xffffffffull;
  uint64_t u = t * (uint64_t)KB_P;
  uint64_t d = x - u;
  uint32_t hi = (uint32_t)(d >> 32);
  return (x < u) ? hi + KB_P : hi;
}
static inline kb_t kb_to_monty(uint32_t canon) { return (kb_t)((((uint64_t)canon) << 32) % KB_P); }
static inline uint32_t kb_from_monty(kb_t x) { return kb_monty_reduce((uint64_t)x); }
static inline kb_t kb_add(kb_t a, kb_t b) { uint32_t s = a + b; return s >= KB_P ? s - KB_P : s; }
static inline kb_t kb_sub(kb_t a, kb_t b) { return a >= b ? a - b : a + KB_P - b; }
static inline kb_t kb_neg(kb_t a) { return a ? KB_P - a : 0; }
static inline kb_t kb_mul(kb_t a, kb_t b) { return kb_monty_reduce((uint64_t)a * (uint64_t)b); }
static inline kb_t kb_dbl(kb_t a) { return kb_add(a, a); }
static inline kb_t kb_pow(kb_t a, uint64_t e) {
  kb_t r = KB_ONE;
  while (e) { if (e & 1) r = kb_mul(r, a); a = kb_mul(a, a); e >>= 1; }
  return r;
}
/* kb31_t.hpp:579-598 computes x^(p-2) by an addition chain; the value is the same. */
static inline kb_t kb_inv(kb_t a) { return kb_pow(a, (uint64_t)KB_P - 2); }
static inline kb_t kb_from_u32(uint32_t x) { return kb_to_monty(x % KB_P); }
/* multiplicative generator 3 (recursion/circuit/src/fri.rs:140) */
static inline kb_t kb_generator(void) { return kb_to_monty(3); }
/* g_k: generator of the subgroup of order 2^k; g_24 = 3^127 */
static inline kb_t kb_two_adic_generator(unsigned k) {
  kb_t g = kb_pow(kb_to_monty(3), 127);
  for (unsigned i = k; i < KB_TWO_ADICITY; i++) g = kb_mul(g, g);
  return g;
}

/* ---- extension ---- */
static inline kb4_t kb4_zero(void) { kb4_t r = {{0, 0, 0, 0}}; return r; }
static inline kb4_t kb4_one(void) { kb4_t r = {{KB_ONE, 0, 0, 0}}; return r; }
static inline kb4_t kb4_from_base(kb_t a) { kb4_t r = {{a, 0, 0, 0}}; return r; }
static inline int kb4_eq(kb4_t a, kb4_t b) {
  return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3];
}
static inline kb4_t kb4_add(kb4_t a, kb4_t b) {
  kb4_t r; for (int i = 0; i < 4; i++) r.c[i] = kb_add(a.c[i], b.c[i]); return r;
}
static inline kb4_t kb4_sub(kb4_t a, kb4_t b) {
  kb4_t r; for (int i = 0; i < 4; i++) r.c[i] = kb_sub(a.c[i], b.c[i]); return r;
}
static inline kb4_t kb4_neg(kb4_t a) {
  kb4_t r; for (int i = 0; i < 4; i++) r.c[i] = kb_neg(a.c[i]); return r;
}
static inline kb4_t kb4_mul_base(kb4_t a, kb_t b) {
  kb4_t r; for (int i = 0; i < 4; i++) r.c[i] = kb_mul(a.c[i], b); return r;
}
static inline kb4_t kb4_add_base(kb4_t a, kb_t b) { a.c[0] = kb_add(a.c[0], b); return a; }
static inline kb4_t kb4_sub_base(kb4_t a, kb_t b) { a.c[0] = kb_sub(a.c[0], b); return a; }
/* schoolbook product, wrap coefficient W = 3 (air/extension.rs:57-75) */
static inline kb4_t kb4_mul(kb4_t a, kb4_t b) {
  kb_t w = kb_to_monty(3);
  kb_t t[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) t[i + j] = kb_add(t[i + j], kb_mul(a.c[i], b.c[j]));
  kb4_t r;
  r.c[0] = kb_add(t[0], kb_mul(w, t[4]));
  r.c[1] = kb_add(t[1], kb_mul(w, t[5]));
  r.c[2] = kb_add(t[2], kb_mul(w, t[6]));
  r.c[3] = t[3];
  return r;
}
static inline kb4_t kb4_sqr(kb4_t a) { return kb4_mul(a, a); }
static inline kb4_t kb4_pow(kb4_t a, uint64_t e) {
  kb4_t r = kb4_one();
  while (e) { if (e & 1) r = kb4_mul(r, a); a = kb4_sqr(a); e >>= 1; }
  return r;
}
/* Inverse through the norm to the quadratic subfield F_p[Y]/(Y^2 - 3), Y = X^2:
 * a = A + X B with A = a0 + a2 Y, B = a1 + a3 Y;  a * (A - X B) = A^2 - Y B^2 =: N (in the subfield);
 * N^{-1} by the conjugate again.  Any correct inverse gives the same (unique) field element. */
static inline kb4_t kb4_inv(kb4_t a) {
  kb_t w = kb_to_monty(3);
  kb_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
  /* A^2 = (a0^2 + 3 a2^2) + (2 a0 a2) Y ;  B^2 = (a1^2 + 3 a3^2) + (2 a1 a3) Y
   * Y B^2 = 3 (2 a1 a3) + (a1^2 + 3 a3^2) Y */
  kb_t n0 = kb_sub(kb_add(kb_mul(a0, a0), kb_mul(w, kb_mul(a2, a2))), kb_mul(w, kb_dbl(kb_mul(a1, a3))));
  kb_t n1 = kb_sub(kb_dbl(kb_mul(a0, a2)), kb_add(kb_mul(a1, a1), kb_mul(w, kb_mul(a3, a3))));
  /* (n0 + n1 Y)^{-1} = (n0 - n1 Y) / (n0^2 - 3 n1^2) */
  kb_t d = kb_inv(kb_sub(kb_mul(n0, n0), kb_mul(w, kb_mul(n1, n1))));
  kb_t i0 = kb_mul(n0, d), i1 = kb_neg(kb_mul(n1, d));
  /* a^{-1} = (A - X B) * (i0 + i1 Y):  A' = a0 + a2 Y, B' = -(a1 + a3 Y) */
  kb_t A0 = kb_add(kb_mul(a0, i0), kb_mul(w, kb_mul(a2, i1)));
  kb_t A1 = kb_add(kb_mul(a0, i1), kb_mul(a2, i0));
  kb_t B0 = kb_add(kb_mul(a1, i0), kb_mul(w, kb_mul(a3, i1)));
  kb_t B1 = kb_add(kb_mul(a1, i1), kb_mul(a3, i0));
  kb4_t r = {{A0, kb_neg(B0), A1, kb_neg(B1)}};
  return r;
}

static inline uint32_t bitrev32(uint32_t x, unsigned bits) {
  uint32_t r = 0;
  for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static inline unsigned log2_exact(uint64_t n) { unsigned k = 0; while ((1ull << k) < n) k++; return k; }

#endif
