// ntt_tma.cuh -- persistent, TMA-fed variant of the shared-memory NTT pass (sm_100a).
//
// Same arithmetic as ntt_pass_smem (ntt_kernels.cuh): size-32 DFT in registers, one table multiply, ONE exchange,
// 2^(5-B) size-2^B DFTs, second table multiply, 64-bit stores.  What changes is how the tile gets on chip.  The
// plain kernel is a load phase, an arithmetic phase and a store phase per CTA, and because every CTA of a wave starts
// at the same time the whole GPU swings between "all SMs load" and "all SMs compute": a pass took the SUM of its HBM
// time and its integer-pipe time (0.68-0.93 ms for 2^28 elements; either alone is 0.33 / 0.45 ms).  Here
//   * one CTA per SM stays resident and walks over its share of the tiles;
//   * a producer warp streams the NEXT tiles into a ring of three 64 KB shared-memory buffers with
//     cp.async.bulk.tensor (TMA, 4-D tensor map: column, interleaved tile, row-in-tile, block) signalling mbarriers;
//   * two groups of 256 threads take the landed tiles alternately, each with its own named barrier, so one group's
//     barrier waits and stores overlap the other group's arithmetic, and neither ever waits for HBM latency;
//   * the exchange happens IN PLACE in the landed buffer (a thread writes its round-A results back to the slots it
//     read), so a tile costs 64 KB of shared memory, not 64 + 66.
// A "super-tile" is always 1024 row slots x 16 columns: 2^(5-B) tiles of 2^(5+B) rows that are neighbours in memory
// (consecutive `lo`: their rows interleave; in the last pass of a transform consecutive blocks), so the thread
// layout, the buffer size and the table sizes do not depend on B.
//
// Constraints of the TMA path (otherwise the plain kernel runs): transform size >= 2^10, source pitch and first
// column multiples of 4 words, 16-byte aligned base, even destination pitch.
#pragma once
#include <cuda.h>

#include "ntt_kernels.cuh"

namespace ntt {

constexpr int TMA_GROUP = 256;                 // threads per compute group
constexpr int TMA_NGROUP = 2;
constexpr int TMA_THREADS = TMA_GROUP * TMA_NGROUP + 128;  // + producer warpgroup (one lane works; setmaxnreg is per warpgroup)
constexpr int TMA_REGS_COMPUTE = 112, TMA_REGS_PRODUCER = 24;  // the CTA keeps what it was launched with (640 x 96): 4 x 128 x 112 + 128 x 24 <= 61440 (120 deadlocks)
constexpr int TMA_NBUF = 3;
constexpr uint32_t TMA_TILE_WORDS = 1024 * TILE_COLS;     // 64 KB
constexpr uint32_t TMA_TAB_WORDS = 32 * FSTRIDE + 16 * 32 + 32 + 32;  // F, G (per sub-tile), gk, ct
constexpr size_t TMA_SMEM_BYTES = (size_t)TMA_NBUF * TMA_TILE_WORDS * 4 + (size_t)TMA_NGROUP * TMA_TAB_WORDS * 4 + 64;

struct TmaArgs {
  uint32_t total;    // super-tiles = 2^(n-10) * column groups
  uint32_t ncg;      // column groups of 16
  uint32_t rem;      // n - s0 - K
  uint32_t lay_b;    // 1: last pass of a transform (rem == 0): sub-tiles are consecutive blocks
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, uint32_t c0, uint32_t c1,
                                            uint32_t c2, uint32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void group_sync(uint32_t g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(TMA_GROUP) : "memory");
}

template <int B, int DIR, bool FIRST, bool PASSTW>
__global__ void __launch_bounds__(TMA_THREADS, 1)
    ntt_pass_tma(const __grid_constant__ CUtensorMap tmap, PassArgs A, PassExtra X, TmaArgs T) {
  constexpr int K = 5 + B, LS = 5 - B;
  constexpr uint32_t NTAU = 1u << B, S = 1u << LS, C = TILE_COLS;
  // TMA destinations need 128-byte alignment: the declaration's alignment is honoured for the dynamic window (no static
  // shared memory in this kernel).  No integer round trip on the pointer: that loses the shared address space and every
  // access turns into a generic LD.E / ST.E (seen in the first ncu capture: long-scoreboard and lg-throttle stalls).
  extern __shared__ __align__(1024) uint32_t sm[];
  uint32_t* bufs = sm;
  uint32_t* tabs = sm + TMA_NBUF * TMA_TILE_WORDS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tabs + TMA_NGROUP * TMA_TAB_WORDS);  // full[3], empty[3]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TMA_NBUF);

  if (threadIdx.x == 0) {
    for (int b = 0; b < TMA_NBUF; b++) {
      mbar_init(full0 + 8 * b, 1);
      mbar_init(empty0 + 8 * b, TMA_GROUP);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t n = A.log_n, rem = T.rem;
  // this CTA's super-tiles are CONSECUTIVE (column group fastest): the twiddle tables depend on the row tile only, so a
  // group recomputes them once per row tile, not once per column group (16 tiles share them at 256 columns)
  const uint32_t per = T.total / gridDim.x, extra = T.total % gridDim.x;
  const uint32_t t_begin = blockIdx.x * per + min(blockIdx.x, extra);
  const uint32_t nseq = per + (blockIdx.x < extra ? 1u : 0u);

  if (threadIdx.x >= TMA_GROUP * TMA_NGROUP) {
    // ---------------- producer warpgroup: one lane streams this CTA's super-tiles into the ring ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TMA_REGS_PRODUCER));
    if (threadIdx.x == TMA_GROUP * TMA_NGROUP) {
      for (uint32_t s = 0; s < nseq; s++) {
        const uint32_t b = s % TMA_NBUF, use = s / TMA_NBUF;
        if (use > 0) mbar_wait(empty0 + 8 * b, (use - 1) & 1);
        const uint32_t Tt = t_begin + s;
        const uint32_t cg = Tt % T.ncg, st = Tt / T.ncg;
        const uint32_t dst = smem_u32(bufs + b * TMA_TILE_WORDS), bar = full0 + 8 * b;
        mbar_expect_tx(bar, TMA_TILE_WORDS * 4);
        const uint32_t c0 = cg * C;
        if (FIRST) {
          // natural row lo + i*2^rem sits at memory row bitrev_rem(lo)*2^K + bitrev_K(i): one contiguous block per sub-tile
          const uint32_t lo0 = st << LS;
          if (K >= 8) {
            constexpr uint32_t cps = 1u << (K >= 8 ? K - 8 : 0);
#pragma unroll
            for (uint32_t m = 0; m < 4; m++) {
              const uint32_t sub = m / cps, lo = lo0 + sub;
              const uint32_t row = ((rem ? (__brev(lo) >> (32 - rem)) : 0u) << K) + (m % cps) * 256;
              tma_load_4d(dst + m * 16384, &tmap, bar, c0, row, 0, 0);
            }
          } else {
#pragma unroll
            for (uint32_t sub = 0; sub < S; sub++) {
              const uint32_t lo = lo0 + sub;
              const uint32_t row = (rem ? (__brev(lo) >> (32 - rem)) : 0u) << K;
              tma_load_4d(dst + sub * ((1u << K) * C * 4), &tmap, bar, c0, row, 0, 0);
            }
          }
        } else if (!T.lay_b) {
          const uint32_t t0 = st << LS;
          const uint32_t lo0 = t0 & ((1u << rem) - 1), hi = t0 >> rem;
#pragma unroll
          for (uint32_t m = 0; m < 4; m++) tma_load_4d(dst + m * 16384, &tmap, bar, c0, lo0, m * (256u >> LS), hi);
        } else {
          const uint32_t t0 = st << LS;  // rem == 0: tile index = hi
          if (K >= 8) {
            constexpr uint32_t cps = 1u << (K >= 8 ? K - 8 : 0);
#pragma unroll
            for (uint32_t m = 0; m < 4; m++) tma_load_4d(dst + m * 16384, &tmap, bar, c0, 0, (m % cps) * 256, t0 + m / cps);
          } else {
            constexpr uint32_t spc = 1u << (K < 8 ? 8 - K : 0);
#pragma unroll
            for (uint32_t m = 0; m < 4; m++) tma_load_4d(dst + m * 16384, &tmap, bar, c0, 0, 0, t0 + m * spc);
          }
        }
      }
    }
    return;
  }

  // ---------------- compute groups ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TMA_REGS_COMPUTE));
  const uint32_t g = threadIdx.x / TMA_GROUP, gt = threadIdx.x % TMA_GROUP;
  const uint32_t cp = gt & 7, tau = gt >> 3;
  const bool lay_b = T.lay_b != 0;
  const uint32_t sub = lay_b ? (tau >> B) : (tau & (S - 1));
  const uint32_t tb = lay_b ? (tau & (NTAU - 1)) : (tau >> LS);
  const uint32_t lc = cp * 2;
  // slot (row of the 1024 x 16 buffer) of element i of this thread's sub-tile: base + q * stride for both rounds
  const uint32_t a_base = lay_b ? (sub << K) + tb : (tb << LS) + sub, a_str = lay_b ? NTAU : 32u;   // i = q*2^B + tb
  const uint32_t b_base = lay_b ? (sub << K) + tb * 32 : ((tb * 32) << LS) + sub, b_str = lay_b ? 1u : S;  // i = tb*32 + q
  const uint32_t f_base = (FIRST ? (sub << K) + ((B ? (__brev(tb) >> (32 - (B ? B : 1))) : 0u) << 5) : 0u);
  uint32_t* F = tabs + g * TMA_TAB_WORDS;
  uint32_t* G = F + 32 * FSTRIDE;
  uint32_t* gk = G + 16 * 32;
  uint32_t* ct = gk + 32;
  uint32_t cur_st = 0xffffffffu;

  for (uint32_t s = g; s < nseq; s += TMA_NGROUP) {
    const uint32_t b = s % TMA_NBUF, use = s / TMA_NBUF;
    const uint32_t Tt = t_begin + s;
    const uint32_t cg = Tt % T.ncg, st = Tt / T.ncg;
    const uint32_t col = cg * C + lc;
    const bool ok = col < A.nc;
    uint32_t* buf = bufs + b * TMA_TILE_WORDS;

    // per sub-tile: tile = st*S + sub'; lo / hi / jbase as in ntt_pass_smem
    auto tile_lo = [&](uint32_t sb) { return rem ? ((st << LS) + sb) & ((1u << rem) - 1) : 0u; };
    auto tile_hi = [&](uint32_t sb) { return ((st << LS) + sb) >> rem; };
    // ---- twiddle tables of this row tile (uniform branch: the whole group works on one super-tile)
    if (st != cur_st) {
    cur_st = st;
    group_sync(g);  // the previous tile's readers of G are done
    for (uint32_t t = gt; t < 64 + S * 32; t += TMA_GROUP) {
      if (t < 32) {
        const uint32_t sb = t >> B, u = t & (NTAU - 1);
        gk[t] = PASSTW ? root_pow(A.tw, A.log_L, n, (tile_lo(sb) << A.s0) * u) : kb::ONE;
      } else if (t < 64) {
        const uint32_t e = t - 32, sb = e >> B, u = e & (NTAU - 1);
        const uint32_t jb = (tile_hi(sb) << (n - A.s0)) + tile_lo(sb);
        ct[e] = FIRST ? kb::mul(kb::pow(X.sigma, jb + (u << rem)), X.hinv) : kb::ONE;
      } else {
        const uint32_t e = t - 64, sb = e >> 5, k = e & 31;
        G[e] = PASSTW ? root_pow(A.tw, A.log_L, n, ((tile_lo(sb) << A.s0) * k) << B) : kb::ONE;
      }
    }
    group_sync(g);  // gk, ct ready
    for (uint32_t e = gt; e < 32 * 32; e += TMA_GROUP) {
      const uint32_t fi = e >> 5, k = e & 31;  // fi = sub' * NTAU + tb'
      const uint32_t sb = fi >> B, t = fi & (NTAU - 1);
      uint32_t f = root_pow(A.tw, A.log_L, K, t * k);
      if (PASSTW) f = kb::mul(f, gk[(sb << B) + (k & (NTAU - 1))]);
      if (FIRST) f = kb::mul(f, ct[fi]);
      F[fi * FSTRIDE + k] = f;
    }
    }

    // ---- the tile has landed?
    mbar_wait(full0 + 8 * b, use & 1);
    uint32_t v[2][32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const uint32_t slot = FIRST ? f_base + brev5(q) : a_base + q * a_str;
      uint2 x = *reinterpret_cast<const uint2*>(buf + slot * C + lc);
      v[0][q] = x.x;
      v[1][q] = x.y;
      if (FIRST && q > 0) {
        v[0][q] = kb::mul(v[0][q], X.dq[q]);
        v[1][q] = kb::mul(v[1][q], X.dq[q]);
      }
    }
    group_sync(g);  // F ready; FIRST: every input has been read, the slots may be overwritten in natural order
#pragma unroll
    for (int c = 0; c < 2; c++) dif_groups<5, DIR, 32>(v[c]);
    const uint32_t fidx = ((sub << B) + tb) * FSTRIDE;
#pragma unroll
    for (int q = 0; q < 32; q++) {
      if (q > 0 || FIRST || PASSTW) {
        const uint32_t f = F[fidx + brev5(q)];
        v[0][q] = kb::mul(v[0][q], f);
        v[1][q] = kb::mul(v[1][q], f);
      }
      *reinterpret_cast<uint2*>(buf + (a_base + q * a_str) * C + lc) = make_uint2(v[0][q], v[1][q]);
    }
    group_sync(g);
#pragma unroll
    for (int q = 0; q < 32; q++) {
      uint2 x = *reinterpret_cast<const uint2*>(buf + (b_base + q * b_str) * C + lc);
      v[0][q] = x.x;
      v[1][q] = x.y;
    }
    // this thread is done with the buffer: order its generic-proxy accesses before the next TMA write, release it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(empty0 + 8 * b);
#pragma unroll
    for (int c = 0; c < 2; c++) dif_groups<B, DIR, 32>(v[c]);
    const uint32_t jbase = (tile_hi(sub) << (n - A.s0)) + tile_lo(sub);
    uint32_t* orow = A.dst + A.c0d + col;
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const uint32_t i = tb * 32 + q;
      if (PASSTW && q > 0) {
        const uint32_t gg = G[(sub << 5) + brev5(q)];
        v[0][q] = kb::mul(v[0][q], gg);
        v[1][q] = kb::mul(v[1][q], gg);
      }
      if (ok) *reinterpret_cast<uint2*>(orow + (size_t)(jbase + (i << rem)) * A.wd) = make_uint2(v[0][q], v[1][q]);
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

inline bool tma_pass_ok(const PassArgs& A, uint32_t k) {
  if (k < 6 || k > 10 || A.log_n < 10) return false;
  if ((A.ws & 3u) || (A.c0s & 3u) || ((uintptr_t)A.src & 15u)) return false;           // TMA strides / base: 16 bytes
  if ((A.wd & 1u) || (A.c0d & 1u) || ((uintptr_t)A.dst & 7u) || (A.nc & 1u)) return false;  // 64-bit stores
  const char* e = getenv("ZK_NTT_TMA");  // "0": plain kernels only (A/B measurements; read per call)
  return !(e && e[0] == '0') && encode_tiled_fn() != nullptr;
}

template <int B, int DIR, bool FIRST, bool PASSTW>
inline cudaError_t launch_tma_one(const PassArgs& A, const PassExtra& X, cudaStream_t st, int sms) {
  constexpr uint32_t K = 5 + B, LS = 5 - B;
  const uint32_t n = A.log_n, rem = n - A.s0 - K;
  TmaArgs T;
  T.ncg = (A.nc + TILE_COLS - 1) / TILE_COLS;
  T.total = (1u << (n - 10)) * T.ncg;
  T.rem = rem;
  T.lay_b = (rem == 0) ? 1u : 0u;
  if (rem != 0 && rem + B < 5) return cudaErrorNotSupported;  // rem < 5 - B: sub-tiles would straddle blocks
  CUtensorMap map;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  const cuuint64_t pitch = (cuuint64_t)A.ws * 4;
  if (FIRST) {
    dims[0] = A.nc; dims[1] = 1ull << n; dims[2] = 1; dims[3] = 1;
    strides[0] = pitch; strides[1] = pitch << n; strides[2] = pitch << n;
    box[0] = TILE_COLS; box[1] = K >= 8 ? 256 : (1u << K); box[2] = 1; box[3] = 1;
  } else {
    dims[0] = A.nc; dims[1] = 1ull << rem; dims[2] = 1ull << K; dims[3] = 1ull << A.s0;
    strides[0] = pitch; strides[1] = pitch << rem; strides[2] = pitch << (n - A.s0);
    if (!T.lay_b) {
      box[0] = TILE_COLS; box[1] = 1u << LS; box[2] = 256u >> LS; box[3] = 1;
    } else if (K >= 8) {
      box[0] = TILE_COLS; box[1] = 1; box[2] = 256; box[3] = 1;
    } else {
      box[0] = TILE_COLS; box[1] = 1; box[2] = 1u << K; box[3] = 256u >> K;
    }
  }
  CUresult r = encode_tiled_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)(A.src + A.c0s), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::min<uint32_t>((uint32_t)sms, (T.total + 1) / 2);
  ntt_pass_tma<B, DIR, FIRST, PASSTW><<<grid ? grid : 1, TMA_THREADS, TMA_SMEM_BYTES, st>>>(map, A, X, T);
  return cudaGetLastError();
}

template <int B, int DIR>
inline cudaError_t launch_tma(const PassArgs& A, bool first, const PassExtra& X, cudaStream_t st, int sms) {
  const bool passtw = A.log_n - A.s0 - (5 + B) > 0;
  if constexpr (DIR == DIR_INV) {
    if (first) return cudaErrorNotSupported;
    return passtw ? launch_tma_one<B, DIR, false, true>(A, X, st, sms) : launch_tma_one<B, DIR, false, false>(A, X, st, sms);
  } else {
    if (first) return passtw ? launch_tma_one<B, DIR, true, true>(A, X, st, sms) : launch_tma_one<B, DIR, true, false>(A, X, st, sms);
    return passtw ? launch_tma_one<B, DIR, false, true>(A, X, st, sms) : launch_tma_one<B, DIR, false, false>(A, X, st, sms);
  }
}

template <int DIR>
inline cudaError_t run_pass_tma_dir(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st, int sms) {
  switch (k) {
    case 6: return launch_tma<1, DIR>(A, first, X, st, sms);
    case 7: return launch_tma<2, DIR>(A, first, X, st, sms);
    case 8: return launch_tma<3, DIR>(A, first, X, st, sms);
    case 9: return launch_tma<4, DIR>(A, first, X, st, sms);
    case 10: return launch_tma<5, DIR>(A, first, X, st, sms);
  }
  return cudaErrorNotSupported;
}

template <int B, int DIR>
inline cudaError_t configure_tma_b() {
  cudaError_t e = cudaSuccess;
#define ZK_TMA_ATTR(F, T)                                                                                             \
  if (e == cudaSuccess)                                                                                               \
    e = cudaFuncSetAttribute(ntt_pass_tma<B, DIR, F, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES);
  ZK_TMA_ATTR(false, false) ZK_TMA_ATTR(false, true)
  if constexpr (DIR == DIR_FWD) { ZK_TMA_ATTR(true, false) ZK_TMA_ATTR(true, true) }
#undef ZK_TMA_ATTR
  return e;
}
template <int DIR>
inline cudaError_t configure_tma_dir() {
  cudaError_t e;
  if ((e = configure_tma_b<1, DIR>()) != cudaSuccess) return e;
  if ((e = configure_tma_b<2, DIR>()) != cudaSuccess) return e;
  if ((e = configure_tma_b<3, DIR>()) != cudaSuccess) return e;
  if ((e = configure_tma_b<4, DIR>()) != cudaSuccess) return e;
  return configure_tma_b<5, DIR>();
}

}  // namespace ntt
