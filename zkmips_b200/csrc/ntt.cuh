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
//   * the n stages are cut into passes: floor(n/10) passes of k = 10 stages (ntt_pass10) preceded by one
//     or two register-only passes (k <= 5, ntt_pass_reg) for the remaining n mod 10 stages.  A pass tile is
//     2^k rows of stride 2^(n-s0-k); after its local size-2^k DFT each element is multiplied by the pass
//     twiddle g_n^(lo * 2^s0 * bitrev_k(i)), which makes the remaining stages independent smaller DFTs;
//   * inside a k=10 tile each thread keeps 32 elements of a column in registers: a size-32 DFT with constant
//     twiddles (Shoup multiplication by immediates), one twiddle multiply from a per-CTA shared-memory table,
//     ONE shared-memory exchange, a second size-32 DFT.  One HBM read + one HBM write per pass.
// Fused into the first pass of a transform: bit-reversed row gather (hands the bit-reversed output of
// the inverse transform to the forward one without a separate permutation kernel) and the coset scale
// sigma^j / n.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kb31.cuh"
#include "kb31_host.h"
#include "launch.cuh"

namespace ntt {

constexpr int DIR_FWD = 0;
constexpr int DIR_INV = 1;
constexpr int TILE_COLS = 16;

// Butterfly twiddles are compile-time constants: plain (non-Montgomery) powers of w_64 with their Shoup
// companions floor(w * 2^32 / p), so x*w mod p = x*w - floor(x*w'/2^32)*p costs IMAD.HI + 2 IMAD + VIADDMNMX
// and needs no Montgomery correction (the data stays in Montgomery form: (xR)*w = (xw)R).
struct Tw64 {
  uint32_t w[2][32];   // w_64^(+-e), e < 32
  uint32_t ws[2][32];  // floor(w * 2^32 / p)
};
constexpr Tw64 make_tw64() {
  Tw64 t{};
  uint32_t g6 = kbh::two_adic_generator(6);
  uint32_t g[2] = {g6, kbh::inv(g6)};
  for (int d = 0; d < 2; d++)
    for (int e = 0; e < 32; e++) {
      uint32_t w = kbh::from_monty(kbh::pow(g[d], e));
      t.w[d][e] = w;
      t.ws[d][e] = (uint32_t)(((uint64_t)w << 32) / kbh::P);
    }
  return t;
}
__device__ constexpr Tw64 TW64 = make_tw64();

__device__ __forceinline__ uint32_t shoup_mul(uint32_t x, uint32_t w, uint32_t ws) {
  uint32_t q = __umulhi(x, ws);
  uint32_t r = x * w - q * kb::P;  // in [0, 2p)
  return min(r, r - kb::P);
}

struct PassArgs {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* tw;     // tw[e] = g_L^(+-e) for this direction, e < 2^(L-1)
  const uint32_t* scale;  // optional: element of natural row j is multiplied by scale[j] on load
  uint32_t ws, wd;        // row pitch (words) of src / dst
  uint32_t c0s, c0d;      // first column inside src / dst
  uint32_t nc;            // number of columns transformed
  uint32_t log_n;         // transform size
  uint32_t s0;            // stages done by earlier passes
  uint32_t log_L;         // size of the twiddle table's group
  uint32_t src_bitrev;    // read natural row j from memory row bitrev_n(j)
};

// g_n^(+-E) from the table of g_L powers (E < 2^n); the upper half of the circle is the negated lower.
__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ tw, uint32_t log_L, uint32_t log_n,
                                             uint32_t E) {
  uint32_t idx = E << (log_L - log_n);
  uint32_t half = 1u << (log_L - 1);
  if (idx >= half) return kb::P - __ldg(tw + (idx - half));
  return __ldg(tw + idx);
}

// N / 2^LG independent size-2^LG DIF transforms (LG <= 6) on consecutive groups of v.
template <int LG, int DIR, int N>
__device__ __forceinline__ void dif_groups(uint32_t (&v)[N]) {
#pragma unroll
  for (int t = 0; t < LG; t++) {
    const int half = 1 << (LG - 1 - t);
#pragma unroll
    for (int x = 0; x < N; x++) {
      if ((x & half) == 0) {
        const int e64 = ((x & (half - 1)) << t) << (6 - LG);
        uint32_t u = v[x], z = v[x + half];
        v[x] = kb::add(u, z);
        if (e64 == 0)
          v[x + half] = kb::sub(u, z);
        else
          v[x + half] = shoup_mul(u - z + kb::P, TW64.w[DIR][e64], TW64.ws[DIR][e64]);
      }
    }
  }
}

__host__ __device__ constexpr int brev5(int q) {
  return ((q & 1) << 4) | ((q & 2) << 2) | (q & 4) | ((q & 8) >> 2) | ((q & 16) >> 4);
}

// ---- pass with k = 10 stages -------------------------------------------------------------------
// Tile = 1024 rows (stride 2^rem) x 16 columns; each thread keeps 32 elements per column in registers:
// size-32 DFT, twiddle, ONE shared-memory exchange (padded, conflict free), size-32 DFT, pass twiddle.
// (The first version of this kernel, one column per thread with global twiddle/scale loads, needed 85-104
// instructions per element; this one 51-69: profiles/README.md.)
//   * two adjacent columns per thread (64-bit global and shared accesses, twiddles shared by both);
//   * butterfly twiddles are compile-time constants multiplied with Shoup's method (TW64 above);
//   * the pass twiddle g_n^(base*bitrev_10(i)) is split as g^(base*bitrev_5(i>>5)) * g^(32*base*bitrev_5(i&31)):
//     the first factor, the inner twiddle w_1024^(tau*kappa) and the coset scale sigma^(lo + tau*2^rem)/h are
//     merged into ONE per-CTA shared-memory table F[tau][kappa]; the second factor is a 32-entry table G;
//     the remaining coset factor sigma^(q*32*2^rem) comes from the kernel arguments (constant bank).
//   No per-element global twiddle or scale loads remain.  Requires even pitches, offsets and column count.
struct Pass10Extra {
  uint32_t dq[32];  // sigma^(q * 32 * 2^rem), Montgomery (FIRST passes only)
  uint32_t sigma;   // coset shift of this block (Montgomery)
  uint32_t hinv;    // 1 / n (Montgomery)
};

constexpr int P10_ROWS = 1024, P10_FSTRIDE = 33;
constexpr size_t P10_SMEM = ((size_t)(P10_ROWS + P10_ROWS / 32) * TILE_COLS + 32 * P10_FSTRIDE + 96) * 4;

// CPT = columns per thread: 2 -> 256 threads, 64-bit accesses, 128 registers (16 warps/SM);
//                            1 -> 512 threads, 32-bit accesses, 64 registers (32 warps/SM).
template <int DIR, bool FIRST, bool PASSTW, int CPT>
__global__ void __launch_bounds__(512 / CPT, 2) ntt_pass10(PassArgs A, Pass10Extra X) {
  constexpr int C = TILE_COLS;
  constexpr uint32_t NT = 512 / CPT, CSH = CPT == 2 ? 3 : 4;  // threads, log2(threads per tau)
  ZK_DYN_SMEM(sm);
  uint32_t* sdat = sm;
  uint32_t* F = sm + (P10_ROWS + P10_ROWS / 32) * C;
  uint32_t* G = F + 32 * P10_FSTRIDE;
  uint32_t* gk = G + 32;
  uint32_t* ct = gk + 32;

  const uint32_t cp = threadIdx.x & ((1u << CSH) - 1), tau = threadIdx.x >> CSH;
  const uint32_t ncg = (A.nc + C - 1) / C;
  const uint32_t cg = blockIdx.x % ncg, tile = blockIdx.x / ncg;
  const uint32_t n = A.log_n, rem = n - A.s0 - 10;
  const uint32_t lo = tile & ((1u << rem) - 1), hi = tile >> rem;
  const uint32_t lc = cp * CPT;  // column inside the tile
  const uint32_t col = cg * C + lc;
  const bool ok = col < A.nc;  // nc is even when CPT == 2
  const uint32_t jbase = (hi << (n - A.s0)) + lo;
  const uint32_t base = lo << A.s0;

  if (threadIdx.x < 32) {
    gk[threadIdx.x] = PASSTW ? root_pow(A.tw, A.log_L, n, base * threadIdx.x) : kb::ONE;
  } else if (threadIdx.x < 64) {
    uint32_t k = threadIdx.x - 32;
    G[k] = PASSTW ? root_pow(A.tw, A.log_L, n, (base * k) << 5) : kb::ONE;
  } else if (threadIdx.x < 96) {
    uint32_t t = threadIdx.x - 64;
    ct[t] = FIRST ? kb::mul(kb::pow(X.sigma, jbase + (t << rem)), X.hinv) : kb::ONE;
  }

  uint32_t v[CPT][32];
#pragma unroll
  for (int q = 0; q < 32; q++) {
    uint32_t i = ((uint32_t)q << 5) + tau;
    uint32_t j = jbase + (i << rem);
    uint32_t srow = FIRST ? (__brev(j) >> (32 - n)) : j;
    const uint32_t* sp = A.src + (size_t)srow * A.ws + A.c0s + col;
    if constexpr (CPT == 2) {
      uint2 x = make_uint2(0u, 0u);
      if (ok) x = __ldg(reinterpret_cast<const uint2*>(sp));
      v[0][q] = x.x;
      v[1][q] = x.y;
    } else {
      v[0][q] = ok ? __ldg(sp) : 0u;
    }
    if (FIRST && q > 0) {
#pragma unroll
      for (int c = 0; c < CPT; c++) v[c][q] = kb::mul(v[c][q], X.dq[q]);
    }
  }
  __syncthreads();  // gk, ct ready
  for (uint32_t e = threadIdx.x; e < 1024; e += NT) {
    uint32_t t = e >> 5, k = e & 31;
    uint32_t f = root_pow(A.tw, A.log_L, 10, t * k);
    if (PASSTW) f = kb::mul(f, gk[k]);
    if (FIRST) f = kb::mul(f, ct[t]);
    F[t * P10_FSTRIDE + k] = f;
  }
#pragma unroll
  for (int c = 0; c < CPT; c++) dif_groups<5, DIR, 32>(v[c]);
  __syncthreads();  // F ready
#pragma unroll
  for (int q = 0; q < 32; q++) {
    if (q > 0 || FIRST || PASSTW) {
      uint32_t f = F[tau * P10_FSTRIDE + brev5(q)];
#pragma unroll
      for (int c = 0; c < CPT; c++) v[c][q] = kb::mul(v[c][q], f);
    }
    uint32_t i = ((uint32_t)q << 5) + tau;
    uint32_t* dp = sdat + (i + (i >> 5)) * C + lc;
    if constexpr (CPT == 2)
      *reinterpret_cast<uint2*>(dp) = make_uint2(v[0][q], v[1][q]);
    else
      *dp = v[0][q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 32; q++) {
    uint32_t i = tau * 32 + q;
    const uint32_t* dp = sdat + (i + (i >> 5)) * C + lc;
    if constexpr (CPT == 2) {
      uint2 x = *reinterpret_cast<const uint2*>(dp);
      v[0][q] = x.x;
      v[1][q] = x.y;
    } else {
      v[0][q] = *dp;
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; c++) dif_groups<5, DIR, 32>(v[c]);
#pragma unroll
  for (int q = 0; q < 32; q++) {
    uint32_t i = tau * 32 + q;
    if (PASSTW && q > 0) {
      uint32_t g = G[brev5(q)];
#pragma unroll
      for (int c = 0; c < CPT; c++) v[c][q] = kb::mul(v[c][q], g);
    }
    uint32_t* op = A.dst + (size_t)(jbase + (i << rem)) * A.wd + A.c0d + col;
    if (ok) {
      if constexpr (CPT == 2)
        *reinterpret_cast<uint2*>(op) = make_uint2(v[0][q], v[1][q]);
      else
        *op = v[0][q];
    }
  }
}

// ---- pass with k <= 5 stages: registers only --------------------------------------------------
template <int K, int DIR>
__global__ void __launch_bounds__(256) ntt_pass_reg(PassArgs A, uint64_t total /* tiles * w */) {
  uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const uint32_t col = (uint32_t)(gid % A.nc);
  const uint32_t tile = (uint32_t)(gid / A.nc);
  const uint32_t n = A.log_n, rem = n - A.s0 - K;
  const uint32_t lo = tile & ((1u << rem) - 1), hi = tile >> rem;
  const uint32_t jbase = (hi << (n - A.s0)) + lo;
  constexpr int R = 1 << K;
  uint32_t v[R];
#pragma unroll
  for (int i = 0; i < R; i++) {
    uint32_t j = jbase + ((uint32_t)i << rem);
    uint32_t srow = (A.src_bitrev && n > 0) ? (__brev(j) >> (32 - n)) : j;
    uint32_t x = __ldg(A.src + (size_t)srow * A.ws + A.c0s + col);
    if (A.scale) x = kb::mul(x, __ldg(A.scale + j));
    v[i] = x;
  }
  if constexpr (K > 0) dif_groups<K, DIR, R>(v);
#pragma unroll
  for (int i = 0; i < R; i++) {
    uint32_t x = v[i];
    if (K > 0 && rem > 0) {
      uint32_t E = (lo << A.s0) * (__brev((uint32_t)i) >> (32 - (K > 0 ? K : 1)));
      x = kb::mul(x, root_pow(A.tw, A.log_L, n, E));
    }
    A.dst[(size_t)(jbase + ((uint32_t)i << rem)) * A.wd + A.c0d + col] = x;
  }
}

// ---- host-side launch -------------------------------------------------------------------------
template <int K, int DIR>
inline cudaError_t launch_reg(const PassArgs& A, cudaStream_t st) {
  uint64_t total = (uint64_t)A.nc << (A.log_n - K);
  unsigned blocks = (unsigned)((total + 255) / 256);
  auto kfn = ntt_pass_reg<K, DIR>;
  ZK_LAUNCH(kfn, blocks, 256, 0, st, A, total);
  return cudaGetLastError();
}

template <int DIR, bool FIRST, bool PASSTW, int CPT>
inline cudaError_t launch_pass10_cpt(const PassArgs& A, const Pass10Extra& X, cudaStream_t st) {
  uint32_t ncg = (A.nc + TILE_COLS - 1) / TILE_COLS;
  uint64_t blocks = (1ull << (A.log_n - 10)) * ncg;
  auto kfn = ntt_pass10<DIR, FIRST, PASSTW, CPT>;
  ZK_LAUNCH_COOP(kfn, (unsigned)blocks, 512 / CPT, P10_SMEM, st, A, X);
  return cudaGetLastError();
}

// Function attributes are per device: every context calls this once for its device (a process may drive
// several GPUs, one context each), so nothing is cached in process-wide statics.
template <int DIR, bool FIRST, bool PASSTW>
inline cudaError_t configure_pass10() {
  cudaError_t e = cudaFuncSetAttribute(ntt_pass10<DIR, FIRST, PASSTW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P10_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ntt_pass10<DIR, FIRST, PASSTW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P10_SMEM);
}
inline cudaError_t configure_device() {
  cudaError_t e;
  if ((e = configure_pass10<DIR_FWD, false, false>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_FWD, false, true>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_FWD, true, false>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_FWD, true, true>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_INV, false, false>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_INV, false, true>()) != cudaSuccess) return e;
  if ((e = configure_pass10<DIR_INV, true, false>()) != cudaSuccess) return e;
  return configure_pass10<DIR_INV, true, true>();
}

// columns per thread of the k=10 pass: 2 needs every access 8-byte aligned
inline bool pass10_aligned(const PassArgs& A) {
  return ((A.ws | A.wd | A.c0s | A.c0d | A.nc) & 1u) == 0 && ((uintptr_t)A.src % 8) == 0 && ((uintptr_t)A.dst % 8) == 0;
}
inline int& pass10_cpt_pref() {
  static int pref = [] {
    const char* e = getenv("ZK_NTT_CPT");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  return pref;
}

template <int DIR, bool FIRST, bool PASSTW>
inline cudaError_t launch_pass10(const PassArgs& A, const Pass10Extra& X, cudaStream_t st) {
  if (pass10_cpt_pref() == 2 && pass10_aligned(A)) return launch_pass10_cpt<DIR, FIRST, PASSTW, 2>(A, X, st);
  return launch_pass10_cpt<DIR, FIRST, PASSTW, 1>(A, X, st);
}

template <int DIR>
inline cudaError_t launch_pass(const PassArgs& A, uint32_t k, cudaStream_t st) {
  switch (k) {
    case 0: return launch_reg<0, DIR>(A, st);
    case 1: return launch_reg<1, DIR>(A, st);
    case 2: return launch_reg<2, DIR>(A, st);
    case 3: return launch_reg<3, DIR>(A, st);
    case 4: return launch_reg<4, DIR>(A, st);
    case 5: return launch_reg<5, DIR>(A, st);
    case 6: return launch_reg<6, DIR>(A, st);
  }
  return cudaErrorInvalidValue;
}

// A column range of a row-major matrix: `nc` columns starting at column c0 of rows of pitch w words.
struct Cols {
  uint32_t* ptr;
  uint32_t w, c0;
};

// Full transform of `nc` columns of a 2^log_n-row matrix: natural-order rows in (optionally gathered
// through a bit reversal and scaled), bit-reversed rows out.  dst may alias src only when src_bitrev == 0.
// Pass plan: as many k=10 passes (shared-memory kernel) as fit, preceded by one or two register-only passes
// (k <= 6) for the remaining log_n mod 10 stages.  Register passes are plain streaming kernels.
inline uint32_t plan_passes(uint32_t log_n, uint32_t* ks) {
  uint32_t n10 = log_n / 10, r = log_n - 10 * n10, np = 0;
  if (log_n == 0) {
    ks[np++] = 0;
    return np;
  }
  if (r > 6) {
    ks[np++] = r - 5;
    ks[np++] = 5;
  } else if (r > 0) {
    ks[np++] = r;
  }
  for (uint32_t i = 0; i < n10; i++) ks[np++] = 10;
  return np;
}

struct CosetScale {
  const uint32_t* vec = nullptr;  // sigma^j / n for every natural row j (needed when the first pass is not a k=10 pass)
  uint32_t sigma = 0, hinv = 0;   // the same as scalars (Montgomery)
};

inline cudaError_t transform(Cols src, Cols dst, uint32_t nc, uint32_t log_n, int dir, const uint32_t* tw,
                             uint32_t log_L, const CosetScale* cs, bool src_bitrev, cudaStream_t st) {
  if (nc == 0) return cudaSuccess;
  const uint32_t* scale = cs ? cs->vec : nullptr;
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
    A.scale = p == 0 ? scale : nullptr;
    A.log_n = log_n;
    A.s0 = s0;
    A.log_L = log_L;
    A.src_bitrev = (p == 0 && src_bitrev) ? 1u : 0u;
    cudaError_t e;
    const bool first = p == 0 && cs != nullptr && src_bitrev;  // coset scale + bit-reversed gather come together
    const bool plain = A.scale == nullptr && A.src_bitrev == 0;
    if (k == 10 && (first || plain)) {
      Pass10Extra X{};
      const bool passtw = log_n - s0 - 10 > 0;
      if (first) {
        X.sigma = cs->sigma;
        X.hinv = cs->hinv;
        uint32_t d = kbh::pow(cs->sigma, 32ull << (log_n - 10)), acc = kbh::ONE;
        for (int q = 0; q < 32; q++) {
          X.dq[q] = acc;
          acc = kbh::mul(acc, d);
        }
        A.scale = nullptr;
        if (dir == DIR_FWD)
          e = passtw ? launch_pass10<DIR_FWD, true, true>(A, X, st) : launch_pass10<DIR_FWD, true, false>(A, X, st);
        else
          e = passtw ? launch_pass10<DIR_INV, true, true>(A, X, st) : launch_pass10<DIR_INV, true, false>(A, X, st);
      } else if (dir == DIR_FWD) {
        e = passtw ? launch_pass10<DIR_FWD, false, true>(A, X, st) : launch_pass10<DIR_FWD, false, false>(A, X, st);
      } else {
        e = passtw ? launch_pass10<DIR_INV, false, true>(A, X, st) : launch_pass10<DIR_INV, false, false>(A, X, st);
      }
    } else {
      if (p == 0 && cs != nullptr && A.scale == nullptr) return cudaErrorInvalidValue;  // scale vector was not prepared
      e = dir == DIR_FWD ? launch_pass<DIR_FWD>(A, k, st) : launch_pass<DIR_INV>(A, k, st);
    }
    if (e != cudaSuccess) return e;
    s0 += k;
  }
  return cudaSuccess;
}

// tw[e] = base^e for e < count (base = g_L or its inverse); one thread per entry, square-and-multiply.
__global__ void powers_kernel(uint32_t* out, uint64_t count, uint32_t base, uint32_t init) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  uint32_t r = init, b = base;
  uint64_t k = e;
  while (k) {
    if (k & 1) r = kb::mul(r, b);
    b = kb::mul(b, b);
    k >>= 1;
  }
  out[e] = r;
}

}  // namespace ntt
