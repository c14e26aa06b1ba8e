// fri.cuh -- kernels of TwoAdicFriPcs::open: device-resident duplex challenger, opening reduction
// (barycentric evaluation + alpha-reduced quotients), FRI folding, proof-of-work grind and query gather.
//
// Replaces the CPU loops of Plonky3 `TwoAdicFriPcs::open` / `p3_fri::prover::{commit_phase, answer_query}`
// / `DuplexChallenger` behind the call at crates/stark/src/prover.rs:546-556.  Semantics: SURVEY A.6 and
// A.10, pinned by the in-repo verifier crates/recursion/circuit/src/fri.rs:34-405 and
// crates/recursion/circuit/src/challenger.rs:90-233.
//
// The Fiat-Shamir transcript lives in device memory (Chal) and is advanced by one-thread kernels, so the
// whole commit phase (fold -> Merkle commit -> observe root -> sample beta -> fold ...) is enqueued on one
// stream without a host round trip per layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kb31_host.h"
#include "launch.cuh"
#include "poseidon2.cuh"

__host__ __device__ inline uint64_t mmcs_layer_off(uint32_t L, uint32_t l);

namespace fri {

struct Chal {  // same 34-word image as zk_challenger in include/zkgpu.h
  uint32_t state[16];
  uint32_t in[8];
  uint32_t n_in;
  uint32_t out[8];
  uint32_t n_out;
};

__device__ __forceinline__ void ch_duplex(Chal& c) {  // challenger.rs:221-232
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = (i < 8 && (uint32_t)i < c.n_in) ? c.in[i] : c.state[i];
  c.n_in = 0;
  p2::permute(s);
#pragma unroll
  for (int i = 0; i < 16; i++) c.state[i] = s[i];
#pragma unroll
  for (int i = 0; i < 8; i++) c.out[i] = s[i];
  c.n_out = 8;
}
__device__ __forceinline__ void ch_observe(Chal& c, uint32_t v) {  // challenger.rs:90-98
  c.n_out = 0;
  c.in[c.n_in++] = v;
  if (c.n_in == 8) ch_duplex(c);
}
__device__ __forceinline__ uint32_t ch_sample(Chal& c) {  // challenger.rs:100-106
  if (c.n_in != 0 || c.n_out == 0) ch_duplex(c);
  return c.out[--c.n_out];
}

__global__ void ch_observe_kernel(Chal* ch, const uint32_t* vals, uint32_t n) {
  if (threadIdx.x | blockIdx.x) return;
  Chal c = *ch;
  for (uint32_t i = 0; i < n; i++) ch_observe(c, vals[i]);
  *ch = c;
}
// samples n_ext extension elements (4 base samples each, coefficient 0 first: challenger.rs:201-207)
__global__ void ch_sample_ext_kernel(Chal* ch, uint32_t* out, uint32_t n_ext) {
  if (threadIdx.x | blockIdx.x) return;
  Chal c = *ch;
  for (uint32_t i = 0; i < 4 * n_ext; i++) out[i] = ch_sample(c);
  *ch = c;
}
__global__ void ch_sample_bits_kernel(Chal* ch, uint32_t bits, uint32_t n, uint64_t* out) {
  if (threadIdx.x | blockIdx.x) return;
  Chal c = *ch;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t v = kb::from_monty(ch_sample(c));
    out[i] = bits >= 32 ? v : (v & ((1u << bits) - 1));
  }
  *ch = c;
}
// Proof of work: found = min canonical w in [base, base + threads) with check_witness(bits, w).
// (Plonky3 takes any hit; the smallest one makes runs reproducible.)
__global__ void __launch_bounds__(256) grind_kernel(const Chal* ch, uint32_t bits, uint32_t base, uint32_t count,
                                                    uint32_t* found) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  uint32_t w = base + t;
  if (w >= kb::P) return;
  Chal c = *ch;
  ch_observe(c, kb::to_monty(w));
  uint32_t v = kb::from_monty(ch_sample(c));
  if ((v & ((1u << bits) - 1)) == 0) atomicMin(found, w);
}
// Applies the witness (found by grind, or injected) to the real transcript: check_witness (challenger.rs:209-219).
// status[0] |= 1 when the witness does not satisfy the proof-of-work condition.
__global__ void ch_witness_kernel(Chal* ch, uint32_t bits, const uint32_t* found, int64_t inject, uint32_t* slot,
                                  uint32_t* status) {
  if (threadIdx.x | blockIdx.x) return;
  uint32_t wm = inject >= 0 ? (uint32_t)inject : kb::to_monty(*found);
  if (inject < 0 && *found == 0xffffffffu) status[0] |= 2u;
  Chal c = *ch;
  ch_observe(c, wm);
  uint32_t v = kb::from_monty(ch_sample(c));
  if ((v & ((1u << bits) - 1)) != 0) status[0] |= 1u;
  *ch = c;
  *slot = wm;
}

// ---- extension helpers ---------------------------------------------------------------------------
__device__ __forceinline__ kb::Ext ld_ext(const uint32_t* p) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  return kb::Ext{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void st_ext(uint32_t* p, kb::Ext e) {
  *reinterpret_cast<uint4*>(p) = make_uint4(e.c[0], e.c[1], e.c[2], e.c[3]);
}

// ---- lazy dot products --------------------------------------------------------------------------------
// sum_i v_i * w_i with the weights split into 16-bit halves: acc_lo += v*lo(w), acc_hi += v*hi(w) are single
// IMAD.WIDE accumulations (terms < 2^47, so 2^16 of them fit in 64 bits) instead of a Montgomery product
// and a modular addition per term.  reduce_acc() returns (acc_hi * 2^16 + acc_lo) * R^-1 mod p, i.e. the sum
// of the Montgomery products.
constexpr uint32_t TWO16_MONTY = kbh::to_monty(1u << 16);
__device__ __forceinline__ uint32_t reduce_acc(uint64_t lo, uint64_t hi) {
  constexpr uint32_t TWO16 = TWO16_MONTY;
  uint64_t fl = (lo >> 32) * (uint64_t)kb::ONE + (lo & 0xffffffffull);  // == lo (mod p), < 2^58
  uint64_t fh = (hi >> 32) * (uint64_t)kb::ONE + (hi & 0xffffffffull);
  return kb::add(kb::mont_reduce64(fl), kb::mul(kb::mont_reduce64(fh), TWO16));
}
__device__ __forceinline__ void st_split(uint32_t* p, kb::Ext e) {  // 8 words: low halves, then high halves
  reinterpret_cast<uint4*>(p)[0] = make_uint4(e.c[0] & 0xffffu, e.c[1] & 0xffffu, e.c[2] & 0xffffu, e.c[3] & 0xffffu);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(e.c[0] >> 16, e.c[1] >> 16, e.c[2] >> 16, e.c[3] >> 16);
}
struct Acc4 {  // one extension accumulator: 4 coefficients x (lo, hi)
  uint64_t lo[4], hi[4];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < 4; k++) lo[k] = hi[k] = 0;
  }
  __device__ __forceinline__ void fma(uint4 wl, uint4 wh, uint32_t v) {
    lo[0] += (uint64_t)wl.x * v; lo[1] += (uint64_t)wl.y * v; lo[2] += (uint64_t)wl.z * v; lo[3] += (uint64_t)wl.w * v;
    hi[0] += (uint64_t)wh.x * v; hi[1] += (uint64_t)wh.y * v; hi[2] += (uint64_t)wh.z * v; hi[3] += (uint64_t)wh.w * v;
  }
  __device__ __forceinline__ kb::Ext reduce() const {
    return kb::Ext{{reduce_acc(lo[0], hi[0]), reduce_acc(lo[1], hi[1]), reduce_acc(lo[2], hi[2]), reduce_acc(lo[3], hi[3])}};
  }
};

// out[j] = alpha^j, j < n
__global__ void ext_powers_kernel(const uint32_t* alpha, uint32_t* out, uint32_t* out_split, uint32_t n) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  kb::Ext e = kb::ext_pow(ld_ext(alpha), j);
  st_ext(out + 4 * (size_t)j, e);
  st_split(out_split + 8 * (size_t)j, e);
}

// x_r = GENERATOR * g_L^{bitrev_L(r)}: the point of row r of a bit-reversed LDE of height 2^L
__device__ __forceinline__ uint32_t lde_point(uint32_t r, uint32_t L, uint32_t gL) {
  return kb::mul(kb::GEN, kb::pow(gL, kb::bitrev(r, L)));
}

// rowred[r] = sum_j alpha^j m[r][j]   (Matrix::dot_ext_powers).  One thread per row; split alpha powers in smem.
constexpr int RR_CHUNK = 512;  // alpha powers staged per pass
__global__ void __launch_bounds__(256) row_reduce_kernel(const uint32_t* __restrict__ mat, uint64_t H, uint32_t w,
                                                         uint32_t pitch, const uint32_t* __restrict__ apow_split,
                                                         uint32_t* __restrict__ rowred) {
  __shared__ uint4 sp[RR_CHUNK * 2];
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t* row = mat + (r < H ? r : 0) * pitch;
  Acc4 acc;
  acc.zero();
  for (uint32_t c0 = 0; c0 < w; c0 += RR_CHUNK) {
    uint32_t nc = min(w - c0, (uint32_t)RR_CHUNK);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nc * 2; i += blockDim.x)
      sp[i] = reinterpret_cast<const uint4*>(apow_split)[(size_t)c0 * 2 + i];
    __syncthreads();
    if (r < H) {
#pragma unroll 4
      for (uint32_t j = 0; j < nc; j++) acc.fma(sp[2 * j], sp[2 * j + 1], __ldg(row + c0 + j));
    }
  }
  if (r < H) st_ext(rowred + 4 * r, acc.reduce());
}

// Same for wide matrices: one WARP per R consecutive rows, lanes stride the columns (coalesced), partial sums combined
// with shuffles.  The thread-per-row form above leaves most of the machine idle when the LDE has few, long rows.
// R = 4: the split alpha powers of a column (two 16-byte loads) are fetched once and used for four rows, whose four
// data loads are in flight together -- with one row per warp the kernel had 128 bytes in flight per warp and as much
// L1 traffic for the powers as eight times its data (1.7 TB/s); 16-byte data loads with a lane owning four adjacent
// columns were tried and are 3x SLOWER (every load of the powers then touches 32 lines instead of 8).
template <int R>
__global__ void __launch_bounds__(256) row_reduce_warp_kernel(const uint32_t* __restrict__ mat, uint64_t H, uint32_t w,
                                                              uint32_t pitch, const uint32_t* __restrict__ apow_split,
                                                              uint32_t* __restrict__ rowred) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t r0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R;
  if (r0 >= H) return;
  const uint32_t* row = mat + r0 * pitch;
  const uint4* A = reinterpret_cast<const uint4*>(apow_split);
  Acc4 acc[R];
#pragma unroll
  for (int k = 0; k < R; k++) acc[k].zero();
  for (uint32_t j = lane; j < w; j += 32) {
    uint32_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = __ldg(row + (size_t)k * pitch + j);
    const uint4 al = __ldg(A + 2 * j), ah = __ldg(A + 2 * j + 1);
#pragma unroll
    for (int k = 0; k < R; k++) acc[k].fma(al, ah, v[k]);
  }
#pragma unroll
  for (int k = 0; k < R; k++) {
    kb::Ext e = acc[k].reduce();
#ifdef ZK_EMU  // the test-only emulator has no lock-step lanes: combine through shared memory instead
    __shared__ uint32_t sh[256 * 4];
    __syncthreads();
    st_ext(sh + 4 * threadIdx.x, e);
    __syncthreads();
    if (lane == 0)
      for (int l = 1; l < 32; l++) e = kb::ext_add(e, ld_ext(sh + 4 * (threadIdx.x + l)));
#else
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
      for (int c = 0; c < 4; c++) e.c[c] = kb::add(e.c[c], __shfl_xor_sync(0xffffffffu, e.c[c], off));
    }
#endif
    if (lane == 0) st_ext(rowred + 4 * (r0 + k), e);
  }
}

// Denominators of one opening point over a whole LDE domain, computed ONCE per (height, point) and shared by every
// matrix of that height opened there (a chip's main trace, permutation trace and quotient chunks all open at zeta; the
// per-matrix kernels used to redo a 22-step pow and a 31-squaring inversion per row and point):
//   inv[r] = 1 / (z - x_r),  x_r = GENERATOR * g_L^{bitrev_L(r)},  r < 2^L;
//   wts[r] = x_r / (z - x_r) in split form for the rows of the low coset, r < 2^n (its points are the first 2^n of the LDE's).
// A thread walks INV_BATCH consecutive exponents e (x advances by one multiplication) and inverts its batch with
// Montgomery's trick: three extension products per element and one extension inversion per batch.
constexpr int INV_BATCH = 8;
__global__ void __launch_bounds__(256) inv_den_kernel(const uint32_t* __restrict__ pt, uint32_t L, uint32_t n, uint32_t gL,
                                                      uint32_t* __restrict__ inv, uint32_t* __restrict__ wts) {
  const uint64_t H = 1ull << L;
  const uint64_t e0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * INV_BATCH;
  if (e0 >= H) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)INV_BATCH, H - e0);
  const kb::Ext z = ld_ext(pt);
  uint32_t x = kb::mul(kb::GEN, kb::pow(gL, e0));
  uint32_t xs[INV_BATCH];
  kb::Ext d[INV_BATCH], pre[INV_BATCH], run = kb::ext_one();
#pragma unroll
  for (int i = 0; i < INV_BATCH; i++) {
    if ((uint32_t)i < cnt) {
      xs[i] = x;
      d[i] = kb::ext_sub_base(z, x);
      run = i == 0 ? d[0] : kb::ext_mul(run, d[i]);
      pre[i] = run;
      x = kb::mul(x, gL);
    }
  }
  kb::Ext acc = kb::ext_inv(run);
#pragma unroll
  for (int i = INV_BATCH - 1; i >= 0; i--) {
    if ((uint32_t)i < cnt) {
      kb::Ext v = i == 0 ? acc : kb::ext_mul(acc, pre[i - 1]);
      if (i > 0) acc = kb::ext_mul(acc, d[i]);
      const uint64_t r = kb::bitrev((uint32_t)(e0 + i), L);
      st_ext(inv + 4 * r, v);
      if (r < (1ull << n)) st_split(wts + 8 * r, kb::ext_mul_base(v, xs[i]));
    }
  }
}

// partial[chunk][p][c] = sum over the rows of the chunk of wts[p][r] * m[r][c].
// Block = CW column lanes x (256 / CW) row lanes, CW = min(32, columns / CPL) rounded up to a power of two, so
// narrow matrices (w = 2, 4: Fibonacci, quotient chunks) still use every thread; a lane owns CPL adjacent
// columns (64-bit loads when CPL == 2); row lanes stride the chunk and are reduced in shared memory.
constexpr int BARY_ROWS = 2048;  // rows per chunk
constexpr int BARY_STAGE = 128;  // rows of weights staged in shared memory at a time
template <int CPL>
__global__ void __launch_bounds__(256) bary_partial_kernel(const uint32_t* __restrict__ mat, uint32_t n, uint32_t w,
                                                           uint32_t pitch, const uint32_t* __restrict__ wts0,
                                                           const uint32_t* __restrict__ wts1, uint32_t npts,
                                                           uint32_t log_cw, uint32_t* __restrict__ partial) {
  // The weights of a row are the same for every column: they are staged through shared memory (coalesced loads, then
  // broadcast reads) instead of four 16-byte global loads per row and thread, and the matrix loads of four rows are
  // issued before their products -- the first version had one 8-byte load in flight per thread and ran at 0.5 TB/s.
  __shared__ uint32_t red[2 * CPL * 4][256];
  __shared__ uint4 sw[2][BARY_STAGE * 2];
  const uint32_t cw = 1u << log_cw, nrl = 256u >> log_cw;
  const uint32_t lane = threadIdx.x & (cw - 1), rl = threadIdx.x >> log_cw;
  const uint32_t ntile = (w + cw * CPL - 1) / (cw * CPL);
  const uint32_t tile = blockIdx.x % ntile, chunk = blockIdx.x / ntile;
  const uint32_t col = (tile * cw + lane) * CPL;
  const uint64_t N = 1ull << n;
  uint64_t r0 = (uint64_t)chunk * BARY_ROWS, r1 = min(N, r0 + BARY_ROWS);
  Acc4 acc[2][CPL];
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int c = 0; c < CPL; c++) acc[p][c].zero();
  const uint4* W[2] = {reinterpret_cast<const uint4*>(wts0), reinterpret_cast<const uint4*>(wts1)};  // per point: N x 8 words
  const bool active = col < w;
  for (uint64_t s0 = r0; s0 < r1; s0 += BARY_STAGE) {
    const uint32_t ns = (uint32_t)min((uint64_t)BARY_STAGE, r1 - s0);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 2 * ns * npts; i += 256) {
      const uint32_t p = i / (2 * ns), k = i - p * 2 * ns;
      sw[p][k] = __ldg(W[p] + 2 * s0 + k);
    }
    __syncthreads();
    if (active) {
      // this thread's rows of the stage: rl, rl + nrl, ... ; four at a time
      uint32_t t = rl;
      for (; t + 3 * nrl < ns; t += 4 * nrl) {
        uint32_t v[4][CPL];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t* src = mat + (s0 + t + u * nrl) * pitch + col;
          if constexpr (CPL == 2) {
            uint2 x = __ldg(reinterpret_cast<const uint2*>(src));
            v[u][0] = x.x;
            v[u][1] = x.y;
          } else {
            v[u][0] = __ldg(src);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int p = 0; p < 2; p++)
            if ((uint32_t)p < npts) {
              const uint4 wl = sw[p][2 * (t + u * nrl)], wh = sw[p][2 * (t + u * nrl) + 1];
#pragma unroll
              for (int c = 0; c < CPL; c++) acc[p][c].fma(wl, wh, v[u][c]);
            }
      }
      for (; t < ns; t += nrl) {
        uint32_t v[CPL];
        const uint32_t* src = mat + (s0 + t) * pitch + col;
        if constexpr (CPL == 2) {
          uint2 x = __ldg(reinterpret_cast<const uint2*>(src));
          v[0] = x.x;
          v[1] = x.y;
        } else {
          v[0] = __ldg(src);
        }
#pragma unroll
        for (int p = 0; p < 2; p++)
          if ((uint32_t)p < npts) {
            const uint4 wl = sw[p][2 * t], wh = sw[p][2 * t + 1];
#pragma unroll
            for (int c = 0; c < CPL; c++) acc[p][c].fma(wl, wh, v[c]);
          }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      kb::Ext e = acc[p][c].reduce();
#pragma unroll
      for (int k = 0; k < 4; k++) red[(p * CPL + c) * 4 + k][threadIdx.x] = e.c[k];
    }
  __syncthreads();
  for (uint32_t s = nrl >> 1; s >= 1; s >>= 1) {
    if (rl < s) {
#pragma unroll
      for (int q = 0; q < 2 * CPL * 4; q++)
        red[q][threadIdx.x] = kb::add(red[q][threadIdx.x], red[q][threadIdx.x + (s << log_cw)]);
    }
    __syncthreads();
  }
  if (rl == 0 && col < w) {
    for (uint32_t p = 0; p < npts; p++)
      for (int c = 0; c < CPL; c++) {
        if (col + c >= w) continue;
        kb::Ext e;
        for (int k = 0; k < 4; k++) e.c[k] = red[(p * CPL + c) * 4 + k][threadIdx.x];
        st_ext(partial + 4 * (((size_t)chunk * npts + p) * w + col + c), e);
      }
  }
}

// ys[p][c] = scale_p * sum_chunks partial;  scale_p = (z^N - s^N) / (N s^N), s = GENERATOR.
// One block per column: the chunk partials are summed by 128 threads.  Writes the opened values into the
// proof (point-major, width ext each).
__global__ void __launch_bounds__(128) bary_final_kernel(const uint32_t* __restrict__ partial, uint32_t nchunks,
                                                         uint32_t w, uint32_t n, const uint32_t* __restrict__ pts,
                                                         uint32_t npts, uint32_t* __restrict__ ys) {
  __shared__ uint32_t red[4][128];
  const uint32_t c = blockIdx.x;
  for (uint32_t p = 0; p < npts; p++) {
    kb::Ext acc = kb::ext_zero();
    for (uint32_t k = threadIdx.x; k < nchunks; k += 128) acc = kb::ext_add(acc, ld_ext(partial + 4 * (((size_t)k * npts + p) * w + c)));
    __syncthreads();
    for (int k = 0; k < 4; k++) red[k][threadIdx.x] = acc.c[k];
    __syncthreads();
    for (uint32_t s = 64; s >= 1; s >>= 1) {
      if (threadIdx.x < s)
        for (int k = 0; k < 4; k++) red[k][threadIdx.x] = kb::add(red[k][threadIdx.x], red[k][threadIdx.x + s]);
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      acc = kb::Ext{{red[0][0], red[1][0], red[2][0], red[3][0]}};
      kb::Ext z = ld_ext(pts + 4 * p);
      kb::Ext zN = z;
      uint32_t sN = kb::GEN;
      for (uint32_t i = 0; i < n; i++) {
        zN = kb::ext_sqr(zN);
        sN = kb::sqr(sN);
      }
      uint32_t Nm = kb::to_monty(1u << n);  // n <= 22
      kb::Ext scale = kb::ext_mul_base(kb::ext_sub_base(zN, sN), kb::inv(kb::mul(Nm, sN)));
      st_ext(ys + 4 * ((size_t)p * w + c), kb::ext_mul(acc, scale));
    }
  }
}

// red[p] = sum_j alpha^j ys[p][j];  aoff[p] = alpha^(offset + p * w).   One block.
__global__ void __launch_bounds__(256) reduce_ys_kernel(const uint32_t* __restrict__ ys, const uint32_t* __restrict__ apow,
                                                        const uint32_t* __restrict__ alpha, uint32_t w, uint32_t npts,
                                                        uint64_t offset, uint32_t* __restrict__ red_ys,
                                                        uint32_t* __restrict__ aoff) {
  __shared__ uint32_t sm[256 * 4];
  for (uint32_t p = 0; p < npts; p++) {
    kb::Ext acc = kb::ext_zero();
    for (uint32_t j = threadIdx.x; j < w; j += blockDim.x)
      acc = kb::ext_add(acc, kb::ext_mul(ld_ext(apow + 4 * (size_t)j), ld_ext(ys + 4 * ((size_t)p * w + j))));
    __syncthreads();
    st_ext(sm + 4 * threadIdx.x, acc);
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) st_ext(sm + 4 * threadIdx.x, kb::ext_add(ld_ext(sm + 4 * threadIdx.x), ld_ext(sm + 4 * (threadIdx.x + s))));
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      st_ext(red_ys + 4 * p, ld_ext(sm));
      st_ext(aoff + 4 * p, kb::ext_pow(ld_ext(alpha), offset + (uint64_t)p * w));
    }
  }
}

// ro[r] += sum_p aoff[p] * (rowred[r] - red_ys[p]) / (x_r - z_p);  inv_p[r] = 1 / (z_p - x_r) from inv_den_kernel
__global__ void __launch_bounds__(256) ro_accumulate_kernel(uint32_t* __restrict__ ro, const uint32_t* __restrict__ rowred,
                                                            uint32_t L, const uint32_t* __restrict__ inv0,
                                                            const uint32_t* __restrict__ inv1, uint32_t npts,
                                                            const uint32_t* __restrict__ red_ys,
                                                            const uint32_t* __restrict__ aoff) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (1u << L)) return;
  kb::Ext rr = ld_ext(rowred + 4 * (size_t)r);
  kb::Ext acc = ld_ext(ro + 4 * (size_t)r);
  for (uint32_t p = 0; p < npts; p++) {
    kb::Ext inv_den = kb::ext_neg(ld_ext((p == 0 ? inv0 : inv1) + 4 * (size_t)r));  // 1 / (x - z)
    kb::Ext t = kb::ext_mul(kb::ext_sub(rr, ld_ext(red_ys + 4 * p)), inv_den);
    acc = kb::ext_add(acc, kb::ext_mul(ld_ext(aoff + 4 * p), t));
  }
  st_ext(ro + 4 * (size_t)r, acc);
}

// FRI fold (TwoAdicFriGenericConfig::fold_matrix + the beta^2 roll-in of commit_phase; mirror fri.rs:308-351):
// out[k] = (1/2 + beta/(2 x_k)) in[2k] + (1/2 - beta/(2 x_k)) in[2k+1] + beta^2 * ro_next[k],
// x_k = g_L^{bitrev_{L-1}(k)}, L = log2(len(in)).
__global__ void __launch_bounds__(256) fold_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t L,
                                                   uint32_t gL_inv, const uint32_t* __restrict__ beta,
                                                   const uint32_t* __restrict__ ro_next) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (1u << (L - 1))) return;
  kb::Ext b = ld_ext(beta);
  uint32_t inv_x = kb::pow(gL_inv, kb::bitrev(k, L - 1));
  kb::Ext pw = kb::ext_mul_base(b, kb::mul(kb::HALF, inv_x));
  kb::Ext ca = kb::ext_add_base(pw, kb::HALF);
  kb::Ext cb = kb::ext_neg(kb::ext_sub_base(pw, kb::HALF));
  kb::Ext e0 = ld_ext(in + 8 * (size_t)k), e1 = ld_ext(in + 8 * (size_t)k + 4);
  kb::Ext f = kb::ext_add(kb::ext_mul(ca, e0), kb::ext_mul(cb, e1));
  if (ro_next) f = kb::ext_add(f, kb::ext_mul(kb::ext_sqr(b), ld_ext(ro_next + 4 * (size_t)k)));
  st_ext(out + 4 * (size_t)k, f);
}

// status |= 4 when the final folded vector is not constant; writes final_poly.
__global__ void final_poly_kernel(const uint32_t* __restrict__ folded, uint32_t len, uint32_t* __restrict__ slot,
                                  uint32_t* status) {
  if (threadIdx.x | blockIdx.x) return;
  kb::Ext f0 = ld_ext(folded);
  for (uint32_t k = 1; k < len; k++) {
    kb::Ext f = ld_ext(folded + 4 * k);
    if (f.c[0] != f0.c[0] || f.c[1] != f0.c[1] || f.c[2] != f0.c[2] || f.c[3] != f0.c[3]) status[0] |= 4u;
  }
  st_ext(slot, f0);
}

// The LAST layers of the commit phase in ONE launch.  A layer of <= 1024 leaf rows is a chain of single-permutation-deep
// steps -- leaf hash, log2 tree levels, observe the root, sample beta, fold -- that cost 5 launches and ~100 us each as
// separate kernels (profiles/r2_launches_shard_prove.csv); one CTA walks all of them with block barriers instead, then
// checks that the last folded vector is constant, writes final_poly and observes it.  Same transcript order as the
// per-layer path: commit -> observe -> sample beta -> fold (fri.rs:308-351 mirror).
__device__ __forceinline__ void st_digest(uint32_t* out, const uint32_t (&s)[16]) {
  reinterpret_cast<uint4*>(out)[0] = make_uint4(s[0], s[1], s[2], s[3]);
  reinterpret_cast<uint4*>(out)[1] = make_uint4(s[4], s[5], s[6], s[7]);
}
struct TailLayer {
  uint32_t* cur;           // 2^L extension elements = 2^(L-1) leaf rows of 8 words
  uint32_t* nxt;           // 2^(L-1) extension elements
  uint32_t* digests;       // 8 * (2^L - 1) words, leaves first (mmcs_layer_off(L - 1, l))
  const uint32_t* ro_next; // reduced openings of height 2^(L-1) rolled in with beta^2, or null
  uint32_t* commit_slot;   // 8 words in the proof
  uint32_t L, gL_inv;
};
constexpr int FRI_TAIL_MAX_LOG = 11;  // layers whose input has <= 2^11 elements (<= 1024 leaf rows)
__global__ void __launch_bounds__(1024) fri_tail_kernel(const TailLayer* __restrict__ layers, uint32_t n_layers, Chal* ch,
                                                        uint32_t final_len, uint32_t* __restrict__ final_slot,
                                                        uint32_t* __restrict__ status) {
  __shared__ uint32_t s_beta[4];
  Chal c;
  if (threadIdx.x == 0) c = *ch;
  const uint32_t* last = nullptr;
  for (uint32_t li = 0; li < n_layers; li++) {
    const TailLayer ly = layers[li];
    const uint32_t hh = 1u << (ly.L - 1);
    // leaf layer: digest of row r = sponge of its 8 words (one permutation)
    for (uint32_t r = threadIdx.x; r < hh; r += blockDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(ly.cur + 8 * (size_t)r);
      uint4 a = p[0], b = p[1];
      uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0, 0, 0, 0, 0, 0, 0, 0};
      p2::permute(s);
      st_digest(ly.digests + 8 * (size_t)r, s);
    }
    __syncthreads();
    uint64_t off = 0;
    for (uint32_t len = hh; len > 1; len >>= 1) {
      const uint32_t half = len >> 1;
      const uint64_t nx = off + (uint64_t)len * 8;
      for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(ly.digests + off + (uint64_t)i * 16);
        uint4 a = p[0], b = p[1], cc = p[2], d = p[3];
        uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y, cc.z, cc.w, d.x, d.y, d.z, d.w};
        p2::permute(s);
        st_digest(ly.digests + nx + (uint64_t)i * 8, s);
      }
      __syncthreads();
      off = nx;
    }
    if (threadIdx.x == 0) {
      const uint32_t* root = ly.digests + off;
      for (int k = 0; k < 8; k++) {
        const uint32_t v = root[k];
        ly.commit_slot[k] = v;
        ch_observe(c, v);
      }
      for (int k = 0; k < 4; k++) s_beta[k] = ch_sample(c);
    }
    __syncthreads();
    const kb::Ext b = kb::Ext{{s_beta[0], s_beta[1], s_beta[2], s_beta[3]}};
    for (uint32_t k = threadIdx.x; k < hh; k += blockDim.x) {
      uint32_t inv_x = kb::pow(ly.gL_inv, kb::bitrev(k, ly.L - 1));
      kb::Ext pw = kb::ext_mul_base(b, kb::mul(kb::HALF, inv_x));
      kb::Ext ca = kb::ext_add_base(pw, kb::HALF);
      kb::Ext cb = kb::ext_neg(kb::ext_sub_base(pw, kb::HALF));
      kb::Ext e0 = ld_ext(ly.cur + 8 * (size_t)k), e1 = ld_ext(ly.cur + 8 * (size_t)k + 4);
      kb::Ext f = kb::ext_add(kb::ext_mul(ca, e0), kb::ext_mul(cb, e1));
      if (ly.ro_next) f = kb::ext_add(f, kb::ext_mul(kb::ext_sqr(b), ld_ext(ly.ro_next + 4 * (size_t)k)));
      st_ext(ly.nxt + 4 * (size_t)k, f);
    }
    __syncthreads();
    last = ly.nxt;
  }
  if (threadIdx.x == 0) {
    kb::Ext f0 = ld_ext(last);
    for (uint32_t k = 1; k < final_len; k++) {
      kb::Ext f = ld_ext(last + 4 * k);
      if (f.c[0] != f0.c[0] || f.c[1] != f0.c[1] || f.c[2] != f0.c[2] || f.c[3] != f0.c[3]) status[0] |= 4u;
    }
    st_ext(final_slot, f0);
    for (int k = 0; k < 4; k++) ch_observe(c, f0.c[k]);
    *ch = c;
  }
}

// answer_query for one commit-phase layer (p3_fri::prover::answer_query): block q = query q.
// Writes sibling value (4 words) then the path of the pair leaf into proof[q * query_stride + off ...].
__global__ void fri_layer_query_kernel(const uint32_t* __restrict__ leaves /* ext pairs */, const uint32_t* __restrict__ digests,
                                       uint32_t log_h /* of the layer tree */,
                                       uint32_t layer_i, const uint64_t* __restrict__ indices, uint32_t* __restrict__ proof,
                                       uint64_t query_stride, uint64_t off) {
  uint64_t index_i = indices[blockIdx.x] >> layer_i;
  uint64_t pair = index_i >> 1;
  uint32_t* o = proof + blockIdx.x * query_stride + off;
  for (uint32_t t = threadIdx.x; t < 4 + log_h * 8; t += blockDim.x) {
    if (t < 4) {
      o[t] = leaves[pair * 8 + 4 * ((index_i ^ 1) & 1) + t];
    } else {
      uint32_t l = (t - 4) >> 3, k = (t - 4) & 7;
      o[t] = digests[mmcs_layer_off(log_h, l) + (((pair >> l) ^ 1) << 3) + k];
    }
  }
}

}  // namespace fri
