// quotient.cuh -- runtime pieces of the generated quotient kernels (csrc/gen/airs_gen.cuh).
//
// Replaces `quotient_values` (crates/stark/src/quotient.rs:19-171) + `ProverConstraintFolder`
// (crates/stark/src/folder.rs:19-149).  One thread = one point x_i = GENERATOR * g_{n+lqd}^i of the quotient
// domain.  The trace LDEs are NOT re-materialised (the reference copies three matrices per chip through
// get_evaluations_on_domain, crates/stark/src/prover.rs:437-445): natural index i lives at row
// bitrev_{n+lqd}(i) of the committed, bit-reversed LDE, and "next" is natural index i + 2^lqd (quotient.rs:44-45,61).
// Selectors are the unnormalised ones of TwoAdicMultiplicativeCoset::selectors_on_coset (mirror:
// crates/recursion/circuit/src/domain.rs:46-64):  Z_H = x^N - 1, first = Z_H/(x-1), last = Z_H/(x-g^-1),
// transition = x - g^-1, inv_zeroifier = 1/Z_H.
// Output: quotient chunk c (rows i = c mod 2^lqd, crates/stark/src/prover.rs:477-488) as an N x 4 base matrix.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kb31.cuh"

namespace quot {

struct Args {
  const uint32_t* prep;
  const uint32_t* main;
  const uint32_t* perm;
  const uint32_t* alpha_pows;  // alpha_pows[k] = alpha^(n_constraints-1-k), 4 words each
  const uint32_t* chal;        // permutation challenges, 4 words each
  const uint32_t* pvs;         // public values
  const uint32_t* lcs;         // local cumulative sum (4)
  const uint32_t* gcs;         // global cumulative sum (14)
  uint32_t* out;               // 2^lqd chunk matrices of N x 4, back to back
  uint32_t wp, wm, wq;         // row pitches in words (wq = 4 * perm_width)
  uint32_t log_n, lqd;
  uint32_t g_q;                // g_{n+lqd}
  uint32_t g_n_inv;            // g_n^-1
  uint32_t shift_pow_n;        // GENERATOR^N
  uint32_t w_lqd;              // g_{n+lqd}^N: primitive 2^lqd-th root of unity
};

struct Row {
  const uint32_t *p0, *p1, *m0, *m1, *q0, *q1;
  uint32_t is_first, is_last, is_trans, inv_zh;
};

__device__ __forceinline__ kb::Ext ld_ext(const uint32_t* p) {
  uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  return kb::Ext{{v.x, v.y, v.z, v.w}};
}

__device__ __forceinline__ void prologue(const Args& A, uint32_t i, Row& R) {
  const uint32_t bits = A.log_n + A.lqd;
  const uint32_t mask = (1u << bits) - 1;
  const uint32_t r0 = kb::bitrev(i, bits);
  const uint32_t r1 = kb::bitrev((i + (1u << A.lqd)) & mask, bits);
  R.p0 = A.prep + (size_t)r0 * A.wp;
  R.p1 = A.prep + (size_t)r1 * A.wp;
  R.m0 = A.main + (size_t)r0 * A.wm;
  R.m1 = A.main + (size_t)r1 * A.wm;
  R.q0 = A.perm + (size_t)r0 * A.wq;
  R.q1 = A.perm + (size_t)r1 * A.wq;
  uint32_t x = kb::mul(kb::GEN, kb::pow(A.g_q, i));
  // x^N = GENERATOR^N * (g_{n+lqd}^N)^i
  uint32_t xn = kb::mul(A.shift_pow_n, kb::pow(A.w_lqd, i & ((1u << A.lqd) - 1)));
  uint32_t zh = kb::sub(xn, kb::ONE);
  uint32_t a = kb::sub(x, kb::ONE), b = kb::sub(x, A.g_n_inv);
  uint32_t inv_abz = kb::inv(kb::mul(kb::mul(a, b), zh));  // one inversion for the three denominators
  uint32_t inv_ab = kb::mul(inv_abz, zh);
  R.inv_zh = kb::mul(inv_abz, kb::mul(a, b));
  R.is_first = kb::mul(zh, kb::mul(inv_ab, b));
  R.is_last = kb::mul(zh, kb::mul(inv_ab, a));
  R.is_trans = b;
}

// ProverConstraintFolder::assert_zero / assert_zero_ext (folder.rs:79-84,94-102)
__device__ __forceinline__ kb::Ext fold_b(kb::Ext acc, const uint32_t* ap, uint32_t c) {
  return kb::ext_add(acc, kb::ext_mul_base(ld_ext(ap), c));
}
__device__ __forceinline__ kb::Ext fold_e(kb::Ext acc, const uint32_t* ap, kb::Ext c) {
  return kb::ext_add(acc, kb::ext_mul(ld_ext(ap), c));
}

// quotient = accumulator * inv_zeroifier (quotient.rs:160), written into chunk i mod 2^lqd (prover.rs:477-488)
__device__ __forceinline__ void epilogue(const Args& A, uint32_t i, const Row& R, kb::Ext acc, bool first_part) {
  kb::Ext q = kb::ext_mul_base(acc, R.inv_zh);
  uint32_t c = i & ((1u << A.lqd) - 1), j = i >> A.lqd;
  uint4* o = reinterpret_cast<uint4*>(A.out + (((size_t)c << A.log_n) + j) * 4);
  if (!first_part) {
    uint4 v = *o;
    q = kb::ext_add(q, kb::Ext{{v.x, v.y, v.z, v.w}});
  }
  *o = make_uint4(q.c[0], q.c[1], q.c[2], q.c[3]);
}

}  // namespace quot
