// ntt.cuh -- batched multi-column NTT over KoalaBear for row-major matrices (sm_100a).
//
// Replaces Plonky3 `Radix2DitParallel` behind `TwoAdicSubgroupDft::{dft_batch, coset_lde_batch}`
// (type alias crates/stark/src/kb31_poseidon2.rs:179; called from TwoAdicFriPcs::commit, call sites
// crates/stark/src/prover.rs:277,403,497).  The transform result is mathematically unique, so the
// decomposition below is free to differ from the CPU one while staying bit-exact.
//
// Layout: a matrix is h rows x w columns, row-major, u32 Montgomery words.  Every column is one
// polynomial; all columns share the butterfly schedule, so a warp always touches a contiguous run of
// columns of one row (coalesced) and twiddles are warp-uniform.
//
// Decomposition (decimation in frequency, natural order in, bit-reversed order out), "four-step" at
// two levels so that almost all twiddles are compile-time constants:
//   * the n stages are cut into ceil(n/10) passes of 6..10 stages each, as even as possible (n = 16 -> 8 + 8,
//     n = 20 -> 10 + 10, n = 22 -> 7 + 7 + 8); sizes below 2^6 use one register-only pass.  A pass tile is
//     2^k rows of stride 2^(n-s0-k); after its local size-2^k DFT each element is multiplied by the pass
//     twiddle g_n^(lo * 2^s0 * bitrev_k(i)), which makes the remaining stages independent smaller DFTs;
//   * inside a tile (k = 5 + B) each thread keeps 32 elements of a column in registers: a size-32 DFT with
//     constant twiddles (Shoup multiplication by immediates), one twiddle multiply from a per-CTA shared-memory
//     table, ONE shared-memory exchange, then 2^(5-B) size-2^B DFTs.  One HBM read + one HBM write per pass.
// Fused into the first pass of a transform: bit-reversed row gather (hands the bit-reversed output of
// the inverse transform to the forward one without a separate permutation kernel) and the coset scale
// sigma^j / n.  Kernels live in ntt_kernels.cuh, instantiated per direction in ntt_fwd.cu / ntt_inv.cu.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kb31_host.h"

namespace ntt {

constexpr int DIR_FWD = 0;
constexpr int DIR_INV = 1;
constexpr int TILE_COLS = 16;

struct PassArgs {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* tw;     // tw[e] = g_L^(+-e) for this direction, e < 2^(L-1)
  const uint32_t* scale;  // optional: element of natural row j is multiplied by scale[j] on load (register passes)
  uint32_t ws, wd;        // row pitch (words) of src / dst
  uint32_t c0s, c0d;      // first column inside src / dst
  uint32_t nc;            // number of columns transformed
  uint32_t log_n;         // transform size
  uint32_t s0;            // stages done by earlier passes
  uint32_t log_L;         // size of the twiddle table's group
  uint32_t src_bitrev;    // read natural row j from memory row bitrev_n(j)
};

struct PassExtra {   // coset scale of a FIRST shared-memory pass, derived in the kernel from three scalars
  uint32_t dq[32];   // sigma^(q * 2^(rem + B)), Montgomery
  uint32_t sigma;    // coset shift of this block (Montgomery)
  uint32_t hinv;     // 1 / n (Montgomery)
};

// A column range of a row-major matrix: `nc` columns starting at column c0 of rows of pitch w words.
struct Cols {
  uint32_t* ptr;
  uint32_t w, c0;
};

struct CosetScale {
  const uint32_t* vec = nullptr;  // sigma^j / n for every natural row j (needed when the first pass is a register pass)
  uint32_t sigma = 0, hinv = 0;   // the same as scalars (Montgomery)
};

// Pass plan: ceil(n/10) passes of 6..10 stages each, as even as possible (shared-memory kernel); sizes below
// 2^6 and the odd 11 = 6 + 5 use a register-only pass (k <= 5).
inline uint32_t plan_passes(uint32_t log_n, uint32_t* ks) {
  uint32_t np = 0;
  if (log_n <= 5) {
    ks[np++] = log_n;
    return np;
  }
  uint32_t n = (log_n + 9) / 10;
  if (log_n == 11) {
    ks[np++] = 6;
    ks[np++] = 5;
    return np;
  }
  uint32_t base = log_n / n, extra = log_n - base * n;  // `extra` passes get one more stage
  for (uint32_t i = 0; i < n; i++) ks[np++] = base + (i >= n - extra ? 1 : 0);
  return np;
}
// does the first pass of a transform of this size derive the coset scale itself (shared-memory pass)?
inline bool first_pass_is_smem(uint32_t log_n) {
  uint32_t ks[8];
  plan_passes(log_n, ks);
  return ks[0] >= 6;
}

// implemented in ntt_fwd.cu / ntt_inv.cu (one translation unit per direction, compiled in parallel)
cudaError_t run_pass_fwd(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st);
cudaError_t run_pass_inv(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st);
cudaError_t configure_fwd();
cudaError_t configure_inv();
// persistent TMA-fed variant of the shared-memory pass (ntt_tma.cuh; ntt_tma_fwd.cu / ntt_tma_inv.cu): sets *handled when
// the pass qualified (size, alignment, enough tiles to stream) and was launched
cudaError_t run_pass_tma_fwd(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st, bool* handled);
cudaError_t run_pass_tma_inv(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st, bool* handled);
std::atomic<uint64_t>& tma_pass_counter();  // defined in zkgpu.cu (the emulator build never increments it)
cudaError_t configure_tma_fwd();
cudaError_t configure_tma_inv();
inline cudaError_t configure_device() {
  cudaError_t e = configure_fwd();
  if (e == cudaSuccess) e = configure_inv();
  if (e == cudaSuccess) e = configure_tma_fwd();
  if (e == cudaSuccess) e = configure_tma_inv();
  return e;
}

// Full transform of `nc` columns of a 2^log_n-row matrix: natural-order rows in (optionally gathered
// through a bit reversal and scaled), bit-reversed rows out.  dst may alias src only when src_bitrev == 0.
inline cudaError_t transform(Cols src, Cols dst, uint32_t nc, uint32_t log_n, int dir, const uint32_t* tw,
                             uint32_t log_L, const CosetScale* cs, bool src_bitrev, cudaStream_t st) {
  if (nc == 0) return cudaSuccess;
  uint32_t ks[8];
  uint32_t npass = plan_passes(log_n, ks);
  uint32_t s0 = 0;
  for (uint32_t p = 0; p < npass; p++) {
    uint32_t k = ks[p];
    PassArgs A;
    A.src = p == 0 ? src.ptr : dst.ptr;
    A.ws = p == 0 ? src.w : dst.w;
    A.c0s = p == 0 ? src.c0 : dst.c0;
    A.dst = dst.ptr;
    A.wd = dst.w;
    A.c0d = dst.c0;
    A.nc = nc;
    A.tw = tw;
    A.scale = nullptr;
    A.log_n = log_n;
    A.s0 = s0;
    A.log_L = log_L;
    A.src_bitrev = (p == 0 && src_bitrev) ? 1u : 0u;
    PassExtra X{};
    const bool first = p == 0 && cs != nullptr;  // coset scale (always together with the bit-reversed gather)
    if (first) {
      if (!src_bitrev) return cudaErrorInvalidValue;
      if (k >= 6) {
        X.sigma = cs->sigma;
        X.hinv = cs->hinv;
        uint32_t d = kbh::pow(cs->sigma, 1ull << (log_n - 5)), acc = kbh::ONE;  // sigma^(2^(rem + B)), rem + B = n - 5
        for (int q = 0; q < 32; q++) {
          X.dq[q] = acc;
          acc = kbh::mul(acc, d);
        }
      } else {
        if (!cs->vec) return cudaErrorInvalidValue;  // scale vector was not prepared
        A.scale = cs->vec;
      }
    } else if (p == 0 && src_bitrev) {
      return cudaErrorInvalidValue;  // gather without scale is not used
    }
    bool handled = false;
    cudaError_t e = cudaSuccess;
    if (k >= 6) e = dir == DIR_FWD ? run_pass_tma_fwd(A, k, first, X, st, &handled) : run_pass_tma_inv(A, k, first, X, st, &handled);
    if (e == cudaSuccess && !handled) e = dir == DIR_FWD ? run_pass_fwd(A, k, first, X, st) : run_pass_inv(A, k, first, X, st);
    if (e != cudaSuccess) return e;
    s0 += k;
  }
  return cudaSuccess;
}

}  // namespace ntt
