// logup.cuh -- runtime pieces of the generated LogUp permutation-trace kernels (csrc/gen/airs_gen.cuh).
//
// Replaces `generate_permutation_trace` / `populate_local_permutation_row`
// (crates/stark/src/permutation.rs:29-69,102-196), which the reference runs on the CPU between
// commit(main) and commit(permutation) (crates/stark/src/prover.rs:341-364).  Per row and per batch of
// 2^log_quotient_degree lookups:  entry = sum_i (+-mult_i) / (alpha + kind_i + sum_k beta^(k+1) value_ik);
// the last column is the inclusive prefix sum over rows of the row totals (rayon_scan in the reference; a
// three-kernel block scan here); the local cumulative sum is its last element.  One extension inversion
// per batch (Montgomery's trick over the batch) instead of one per lookup; field results are unique.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kb31.cuh"

namespace logup {

struct Args {
  const uint32_t* prep;
  const uint32_t* main;
  const uint32_t* chal;  // alpha, beta (4 words each)
  uint32_t* perm;        // h x wq words
  uint32_t* rowsum;      // h x 4 words
  uint32_t wp, wm, wq;   // row pitches in words (wq = 4 * permutation width)
  uint64_t h;
};

__device__ __forceinline__ kb::Ext ld_ext(const uint32_t* p) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  return kb::Ext{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void st_ext(uint32_t* p, kb::Ext e) {
  *reinterpret_cast<uint4*>(p) = make_uint4(e.c[0], e.c[1], e.c[2], e.c[3]);
}

// sum_i m_i / d_i for a batch of N lookups with ONE inversion: prefix/suffix products of the denominators
template <int N>
__device__ __forceinline__ kb::Ext batch_entry(const kb::Ext (&d)[N], const uint32_t (&m)[N]) {
  kb::Ext pre[N], suf[N];
  pre[0] = kb::ext_one();
#pragma unroll
  for (int i = 1; i < N; i++) pre[i] = kb::ext_mul(pre[i - 1], d[i - 1]);
  suf[N - 1] = kb::ext_one();
#pragma unroll
  for (int i = N - 2; i >= 0; i--) suf[i] = kb::ext_mul(suf[i + 1], d[i + 1]);
  kb::Ext inv = kb::ext_inv(kb::ext_mul(pre[N - 1], d[N - 1]));
  kb::Ext num = kb::ext_zero();
#pragma unroll
  for (int i = 0; i < N; i++) num = kb::ext_add(num, kb::ext_mul_base(kb::ext_mul(pre[i], suf[i]), m[i]));
  return kb::ext_mul(num, inv);
}

// ---- inclusive prefix sum of h extension elements, written into column `col4` (word offset) of perm ---
constexpr int SCAN_T = 1024;
__global__ void __launch_bounds__(SCAN_T) scan_block_kernel(const uint32_t* __restrict__ rowsum, uint64_t h,
                                                            uint32_t* __restrict__ perm, uint32_t wq, uint32_t col4,
                                                            uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t buf[2][SCAN_T * 4];
  uint64_t r = (uint64_t)blockIdx.x * SCAN_T + threadIdx.x;
  kb::Ext v = r < h ? ld_ext(rowsum + 4 * r) : kb::ext_zero();
  int cur = 0;
  st_ext(&buf[0][4 * threadIdx.x], v);
  __syncthreads();
  for (int off = 1; off < SCAN_T; off <<= 1) {
    kb::Ext x = ld_ext(&buf[cur][4 * threadIdx.x]);
    if ((int)threadIdx.x >= off) x = kb::ext_add(x, ld_ext(&buf[cur][4 * (threadIdx.x - off)]));
    st_ext(&buf[cur ^ 1][4 * threadIdx.x], x);
    cur ^= 1;
    __syncthreads();
  }
  kb::Ext s = ld_ext(&buf[cur][4 * threadIdx.x]);
  if (r < h) st_ext(perm + r * wq + col4, s);
  if (threadIdx.x == SCAN_T - 1) st_ext(blocksum + 4 * blockIdx.x, s);
}
// exclusive scan of the block totals (single block, sequential per thread chunk then the same block scan)
__global__ void __launch_bounds__(SCAN_T) scan_totals_kernel(uint32_t* __restrict__ blocksum, uint32_t nblocks) {
  __shared__ uint32_t buf[2][SCAN_T * 4];
  uint32_t per = (nblocks + SCAN_T - 1) / SCAN_T;
  uint32_t b0 = threadIdx.x * per;
  kb::Ext acc = kb::ext_zero();
  for (uint32_t i = 0; i < per && b0 + i < nblocks; i++) acc = kb::ext_add(acc, ld_ext(blocksum + 4 * (b0 + i)));
  int cur = 0;
  st_ext(&buf[0][4 * threadIdx.x], acc);
  __syncthreads();
  for (int off = 1; off < SCAN_T; off <<= 1) {
    kb::Ext x = ld_ext(&buf[cur][4 * threadIdx.x]);
    if ((int)threadIdx.x >= off) x = kb::ext_add(x, ld_ext(&buf[cur][4 * (threadIdx.x - off)]));
    st_ext(&buf[cur ^ 1][4 * threadIdx.x], x);
    cur ^= 1;
    __syncthreads();
  }
  // exclusive prefix of this thread's chunk
  kb::Ext run = threadIdx.x ? ld_ext(&buf[cur][4 * (threadIdx.x - 1)]) : kb::ext_zero();
  for (uint32_t i = 0; i < per && b0 + i < nblocks; i++) {
    kb::Ext t = ld_ext(blocksum + 4 * (b0 + i));
    st_ext(blocksum + 4 * (b0 + i), run);
    run = kb::ext_add(run, t);
  }
}
__global__ void __launch_bounds__(SCAN_T) scan_add_kernel(uint32_t* __restrict__ perm, uint64_t h, uint32_t wq, uint32_t col4,
                                                          const uint32_t* __restrict__ blocksum, uint32_t* __restrict__ last) {
  uint64_t r = (uint64_t)blockIdx.x * SCAN_T + threadIdx.x;
  if (r >= h) return;
  kb::Ext s = kb::ext_add(ld_ext(perm + r * wq + col4), ld_ext(blocksum + 4 * blockIdx.x));
  st_ext(perm + r * wq + col4, s);
  if (r == h - 1) st_ext(last, s);
}

}  // namespace logup
