// ntt_kernels.cuh -- device code of the batched NTT (see ntt.cuh for the decomposition) and the per-direction
// launch dispatch.  Included by ntt_fwd.cu and ntt_inv.cu only.
#pragma once
#include "ntt.cuh"
#include "kb31.cuh"
#include "launch.cuh"

namespace ntt {

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

// ---- shared-memory pass with k = 5 + B stages (B in 1..5) --------------------------------------------------
// Tile = 2^k rows (stride 2^rem) x 16 columns; a thread keeps 32 elements per column in registers:
//   round A: size-32 DFT over q (positions i = q*2^B + tau), immediate twiddles;
//   one multiply from the per-CTA table F[tau][kappa] (kappa = bitrev_5(q)) that merges
//       the inner twiddle w_{2^k}^(tau*kappa), the first factor of the pass twiddle g^(base*(kappa mod 2^B))
//       (base = lo*2^s0; bitrev_k(i) = bitrev_5(i&31)*2^B + bitrev_B(i>>5) and bitrev_B(i>>5) = kappa mod 2^B)
//       and the coset scale sigma^(jbase + tau*2^rem)/n of a FIRST pass;
//   ONE shared-memory exchange (row padding 1/32: conflict free for 32- and 64-bit accesses);
//   round B: 2^(5-B) size-2^B DFTs on the thread's 32 consecutive rows;
//   second factor of the pass twiddle G[bitrev_5(i&31)] = g^(base*2^B*...), store.
// CPT = columns per thread: 2 -> 64-bit accesses, 128 registers; 1 -> any alignment, 64 registers.
// (Three resident CTAs per SM instead of two -- __launch_bounds__(256, 3): 80 registers, 48 B of spill -- were
// measured slower on B200: inverse transform 1.44 -> 1.66 ms, coset transforms 3.17 -> 3.27 ms per 2^20 x 256 LDE.)
// No per-element global twiddle or scale loads.  (First version of this kernel: 85-104 instructions per
// element; this one 49-62: profiles/README.md.)
constexpr int FSTRIDE = 33;
template <int B, int C = TILE_COLS>
constexpr size_t pass_smem_bytes() {
  constexpr int ROWS = 1 << (5 + B);
  return ((size_t)(ROWS + ROWS / 32) * C + (size_t)(1 << B) * FSTRIDE + 96) * 4;
}
// Narrow matrices (<= 4 columns: Fibonacci-like chips, permutation traces and quotient chunks, which are 4 base-field
// columns each) use 4-column tiles: with 16-column tiles three quarters of every CTA's lanes carried zeros, and the
// eighteen 4-column transforms of a shard proof cost as much as its 1024-column one.
constexpr int NARROW_COLS = 4;

template <int B, int DIR, bool FIRST, bool PASSTW, int CPT, int C = TILE_COLS>
__global__ void __launch_bounds__((C / CPT) << B, (CPT == 2 ? 512 : 1024) / ((C / CPT) << B)) ntt_pass_smem(PassArgs A,
                                                                                                          PassExtra X) {
  constexpr int K = 5 + B, ROWS = 1 << K;
  constexpr uint32_t NTAU = 1u << B, NT = NTAU * (C / CPT), CSH = (C / CPT) == 8 ? 3 : (C / CPT) == 16 ? 4 : (C / CPT) == 2 ? 1 : 2;
  ZK_DYN_SMEM(sm);
  uint32_t* sdat = sm;
  uint32_t* F = sm + (ROWS + ROWS / 32) * C;
  uint32_t* G = F + NTAU * FSTRIDE;
  uint32_t* gk = G + 32;
  uint32_t* ct = gk + 32;

  const uint32_t cp = threadIdx.x & ((1u << CSH) - 1), tau = threadIdx.x >> CSH;
  const uint32_t ncg = (A.nc + C - 1) / C;
  const uint32_t cg = blockIdx.x % ncg, tile = blockIdx.x / ncg;
  const uint32_t n = A.log_n, rem = n - A.s0 - K;
  const uint32_t lo = tile & ((1u << rem) - 1), hi = tile >> rem;
  const uint32_t lc = cp * CPT;  // column inside the tile
  const uint32_t col = cg * C + lc;
  const bool ok = col < A.nc;  // nc is even when CPT == 2
  const uint32_t jbase = (hi << (n - A.s0)) + lo;
  const uint32_t base = lo << A.s0;

  for (uint32_t t = threadIdx.x; t < 96; t += NT) {
    if (t < 32) {
      gk[t] = (PASSTW && t < NTAU) ? root_pow(A.tw, A.log_L, n, base * t) : kb::ONE;
    } else if (t < 64) {
      uint32_t k = t - 32;
      G[k] = PASSTW ? root_pow(A.tw, A.log_L, n, (base * k) << B) : kb::ONE;
    } else {
      uint32_t u = t - 64;
      ct[u] = (FIRST && u < NTAU) ? kb::mul(kb::pow(X.sigma, jbase + (u << rem)), X.hinv) : kb::ONE;
    }
  }

  uint32_t v[CPT][32];
#pragma unroll
  for (int q = 0; q < 32; q++) {
    uint32_t i = ((uint32_t)q << B) + tau;
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
  for (uint32_t e = threadIdx.x; e < NTAU * 32; e += NT) {
    uint32_t t = e >> 5, k = e & 31;
    uint32_t f = root_pow(A.tw, A.log_L, K, t * k);
    if (PASSTW) f = kb::mul(f, gk[k & (NTAU - 1)]);
    if (FIRST) f = kb::mul(f, ct[t]);
    F[t * FSTRIDE + k] = f;
  }
#pragma unroll
  for (int c = 0; c < CPT; c++) dif_groups<5, DIR, 32>(v[c]);
  __syncthreads();  // F ready
#pragma unroll
  for (int q = 0; q < 32; q++) {
    if (q > 0 || FIRST || PASSTW) {
      uint32_t f = F[tau * FSTRIDE + brev5(q)];
#pragma unroll
      for (int c = 0; c < CPT; c++) v[c][q] = kb::mul(v[c][q], f);
    }
    uint32_t i = ((uint32_t)q << B) + tau;
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
  for (int c = 0; c < CPT; c++) dif_groups<B, DIR, 32>(v[c]);
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

// ---- per-direction launch dispatch ---------------------------------------------------------------------
template <int K, int DIR>
inline cudaError_t launch_reg(const PassArgs& A, cudaStream_t st) {
  uint64_t total = (uint64_t)A.nc << (A.log_n - K);
  unsigned blocks = (unsigned)((total + 255) / 256);
  auto kfn = ntt_pass_reg<K, DIR>;
  ZK_LAUNCH(kfn, blocks, 256, 0, st, A, total);
  return cudaGetLastError();
}

// columns per thread: 2 needs every access 8-byte aligned (env ZK_NTT_CPT=1 forces the scalar variant).  Letting the
// two-column kernel fall back to 32-bit accesses per side at run time (for odd chip widths) was measured: the extra
// uniform branches in the 32-deep load / store loops cost the ALIGNED case 1.44 -> 1.71 ms (inverse) and 3.17 -> 5.0 ms
// (coset) per 2^20 x 256 LDE, and the odd-width execution shard got slower too (NTT 20.3 -> 28.6 ms): reverted.  A
// compile-time two-column variant with 32-bit global accesses (no run-time branches) left the aligned case untouched but
// was still slower than one column per thread on the odd widths (20.3 -> 22.0 ms): it is the 64-bit accesses that pay.
inline bool pass_aligned(const PassArgs& A) {
  return ((A.ws | A.wd | A.c0s | A.c0d | A.nc) & 1u) == 0 && ((uintptr_t)A.src % 8) == 0 && ((uintptr_t)A.dst % 8) == 0;
}
inline int cpt_pref() {
  static int pref = [] {
    const char* e = getenv("ZK_NTT_CPT");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  return pref;
}

template <int B, int DIR, bool FIRST, bool PASSTW, int CPT, int C = TILE_COLS>
inline cudaError_t launch_smem_cpt(const PassArgs& A, const PassExtra& X, cudaStream_t st) {
  uint32_t ncg = (A.nc + C - 1) / C;
  uint64_t blocks = (1ull << (A.log_n - (5 + B))) * ncg;
  auto kfn = ntt_pass_smem<B, DIR, FIRST, PASSTW, CPT, C>;
  ZK_LAUNCH_COOP(kfn, (unsigned)blocks, (C / CPT) << B, (pass_smem_bytes<B, C>()), st, A, X);
  return cudaGetLastError();
}
// Ragged widths (EXPERIMENT, off by default, ZK_NTT_SPLIT=1 enables it): a 16-column tile whose column group holds only a
// few real columns still runs all its lanes, and real chips are rarely a multiple of 16 columns wide (19, 36, 44, 62, 66,
// 71 ...; permutation traces 8, 20, 36, 56 ...).  Sending the last column group of <= 8 (<= 4) columns out as its own
// launch on 8- (4-) column tiles is bit-exact (tests/test_commit_parity.py::test_ntt_ragged_widths ran green with it on)
// and did NOT pay: core shard 24.93 -> 25.26 ms, recursion 18.96 -> 19.02, mixed 19.99 -> 19.96
// (profiles/r2b_ntt_split_ab.txt) -- at these heights (2^16..2^19 rows) a pass is bound by launch and memory latency,
// not by the idle lanes' arithmetic, and the extra launch costs what the narrower tile saves.
constexpr int HALF_COLS = 8;
template <int B, int DIR, int C>
inline cudaError_t launch_smem_c(const PassArgs& A, bool first, bool passtw, const PassExtra& X, cudaStream_t st) {
  if (first) return passtw ? launch_smem_cpt<B, DIR, true, true, 2, C>(A, X, st) : launch_smem_cpt<B, DIR, true, false, 2, C>(A, X, st);
  return passtw ? launch_smem_cpt<B, DIR, false, true, 2, C>(A, X, st) : launch_smem_cpt<B, DIR, false, false, 2, C>(A, X, st);
}
inline bool split_pref() {
  static bool on = [] {
    const char* e = getenv("ZK_NTT_SPLIT");
    return e && e[0] == '1';
  }();
  return on;
}
template <int B, int DIR>
inline cudaError_t launch_smem(const PassArgs& A, bool first, const PassExtra& X, cudaStream_t st) {
  const bool passtw = A.log_n - A.s0 - (5 + B) > 0;
  const bool two = cpt_pref() == 2 && pass_aligned(A);
  if (!two) {
    if (first) return passtw ? launch_smem_cpt<B, DIR, true, true, 1>(A, X, st) : launch_smem_cpt<B, DIR, true, false, 1>(A, X, st);
    return passtw ? launch_smem_cpt<B, DIR, false, true, 1>(A, X, st) : launch_smem_cpt<B, DIR, false, false, 1>(A, X, st);
  }
  if constexpr (B >= 3) {
    const uint32_t rem = A.nc % TILE_COLS, main = A.nc - rem;
    const bool split = split_pref() ? rem != 0 && rem <= HALF_COLS : (main == 0 && rem <= NARROW_COLS && B >= 4);
    if (split) {
      if (main) {
        PassArgs M = A;
        M.nc = main;
        cudaError_t e = launch_smem_c<B, DIR, TILE_COLS>(M, first, passtw, X, st);
        if (e != cudaSuccess) return e;
      }
      PassArgs R = A;
      R.c0s += main;
      R.c0d += main;
      R.nc = rem;
      if constexpr (B >= 4)
        if (rem <= NARROW_COLS) return launch_smem_c<B, DIR, NARROW_COLS>(R, first, passtw, X, st);
      return launch_smem_c<B, DIR, HALF_COLS>(R, first, passtw, X, st);
    }
  }
  return launch_smem_c<B, DIR, TILE_COLS>(A, first, passtw, X, st);
}

template <int DIR>
inline cudaError_t run_pass(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st) {
  switch (k) {
    case 0: return launch_reg<0, DIR>(A, st);
    case 1: return launch_reg<1, DIR>(A, st);
    case 2: return launch_reg<2, DIR>(A, st);
    case 3: return launch_reg<3, DIR>(A, st);
    case 4: return launch_reg<4, DIR>(A, st);
    case 5: return launch_reg<5, DIR>(A, st);
    case 6: return launch_smem<1, DIR>(A, first, X, st);
    case 7: return launch_smem<2, DIR>(A, first, X, st);
    case 8: return launch_smem<3, DIR>(A, first, X, st);
    case 9: return launch_smem<4, DIR>(A, first, X, st);
    case 10: return launch_smem<5, DIR>(A, first, X, st);
  }
  return cudaErrorInvalidValue;
}

// Function attributes are per device: every context calls this once for its device (a process may drive
// several GPUs, one context each), so nothing is cached in process-wide statics.
template <int B, int DIR>
inline cudaError_t configure_b() {
  cudaError_t e = cudaSuccess;
#define ZK_NTT_ATTR(F, T, CPTV)                                                                                   \
  if (e == cudaSuccess)                                                                                           \
    e = cudaFuncSetAttribute(ntt_pass_smem<B, DIR, F, T, CPTV>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             (int)pass_smem_bytes<B>());                                                          \
  if constexpr (B >= 4 && CPTV == 2)                                                                              \
    if (e == cudaSuccess)                                                                                         \
      e = cudaFuncSetAttribute(ntt_pass_smem<B, DIR, F, T, 2, NARROW_COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               (int)pass_smem_bytes<B, NARROW_COLS>());                                         \
  if constexpr (B >= 3 && CPTV == 2)                                                                              \
    if (e == cudaSuccess)                                                                                         \
      e = cudaFuncSetAttribute(ntt_pass_smem<B, DIR, F, T, 2, HALF_COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               (int)pass_smem_bytes<B, HALF_COLS>());
  ZK_NTT_ATTR(false, false, 1) ZK_NTT_ATTR(false, true, 1) ZK_NTT_ATTR(true, false, 1) ZK_NTT_ATTR(true, true, 1)
  ZK_NTT_ATTR(false, false, 2) ZK_NTT_ATTR(false, true, 2) ZK_NTT_ATTR(true, false, 2) ZK_NTT_ATTR(true, true, 2)
#undef ZK_NTT_ATTR
  return e;
}
template <int DIR>
inline cudaError_t configure_dir() {
  cudaError_t e;
  if ((e = configure_b<1, DIR>()) != cudaSuccess) return e;
  if ((e = configure_b<2, DIR>()) != cudaSuccess) return e;
  if ((e = configure_b<3, DIR>()) != cudaSuccess) return e;
  if ((e = configure_b<4, DIR>()) != cudaSuccess) return e;
  return configure_b<5, DIR>();
}

}  // namespace ntt
