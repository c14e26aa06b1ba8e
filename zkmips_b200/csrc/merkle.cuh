// merkle.cuh -- Poseidon2 Merkle kernels: leaf sponge over matrix rows, 2-to-1 compression layers with
// the mixed-height injection rule of Plonky3's MerkleTreeMmcs.
//
// Replaces `MerkleTreeMmcs<_, _, MyHash, MyCompress, 8>::commit` (crates/stark/src/kb31_poseidon2.rs:173-177);
// tree layout pinned by the in-repo verifier crates/recursion/circuit/src/fri.rs:363-405:
//   leaf[r]      = H(concat over the tallest matrices, in input order, of row r)
//   node         = C(left, right); when shorter matrices have exactly this layer's height:
//   node         = C(node, H(concat of their rows at this index))
//
// One thread owns one sponge (one row): the 16-word state stays in registers for the whole row and
// the next 32-byte chunk is fetched while the current permutation runs.  These kernels are bound by
// the integer pipes, not by HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "poseidon2.cuh"

namespace mk {

struct MatDesc {
  const uint32_t* ptr;
  uint32_t w;
  uint32_t pitch;  // row stride in words
};

__device__ __forceinline__ void store_digest(uint32_t* out, const uint32_t (&s)[16]) {
  uint4* o = reinterpret_cast<uint4*>(out);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]);
  o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// Single matrix, width a multiple of 8, 32-byte aligned rows: vector loads, software prefetch.
__global__ void __launch_bounds__(256) hash_rows_w8(const uint32_t* __restrict__ mat, uint32_t w, uint64_t h,
                                                    uint32_t* __restrict__ out) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h) return;
  const uint4* row = reinterpret_cast<const uint4*>(mat + r * w);
  const uint32_t nchunk = w >> 3;
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = 0;
  uint4 a = __ldg(row), b = __ldg(row + 1);
  for (uint32_t c = 0; c < nchunk; c++) {
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if (c + 1 < nchunk) {
      a = __ldg(row + 2 * (c + 1));
      b = __ldg(row + 2 * (c + 1) + 1);
    }
    p2::permute(s);
  }
  store_digest(out + r * 8, s);
}

// Resumable leaf sponge for the streaming commit: absorbs columns [c0, c0 + 8*nchunk) of every row.  The full
// 16-word sponge state of row r lives in state[k*h + r], k < 4 (coalesced uint4 per thread) between slabs.
__global__ void __launch_bounds__(256) hash_rows_slab(const uint32_t* __restrict__ mat, uint32_t pitch, uint32_t c0,
                                                      uint32_t nchunk, uint64_t h, uint4* __restrict__ state, int first,
                                                      int last, uint32_t* __restrict__ out) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h) return;
  const uint4* row = reinterpret_cast<const uint4*>(mat + r * pitch + c0);
  uint32_t s[16];
  if (first) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = 0;
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint4 v = state[k * h + r];
      s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w;
    }
  }
  uint4 a = __ldg(row), b = __ldg(row + 1);
  for (uint32_t c = 0; c < nchunk; c++) {
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    if (c + 1 < nchunk) {
      a = __ldg(row + 2 * (c + 1));
      b = __ldg(row + 2 * (c + 1) + 1);
    }
    p2::permute(s);
  }
  if (last) {
    store_digest(out + r * 8, s);
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) state[k * h + r] = make_uint4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
  }
}

// Resumable sponge at ANY alignment: absorbs columns [c0, c0 + nc) of every row of one matrix into the sponge of
// the row's height class, which may already hold k0 = (words absorbed so far) mod 8 pending words in its rate
// (a class of several matrices is the concatenation of their rows in input order, so a matrix of width 2 in front
// shifts every later chunk boundary by 2).  State layout as hash_rows_slab.  `last` finalises: a pending partial
// block is permuted once (PaddingFreeSponge) and the digest written.  nc == 0 is allowed (zero-width member).
__global__ void __launch_bounds__(256) hash_rows_slab_any(const uint32_t* __restrict__ mat, uint32_t pitch, uint32_t c0,
                                                          uint32_t nc, uint64_t h, uint4* __restrict__ state, uint32_t k0,
                                                          int first, int last, uint32_t* __restrict__ out) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h) return;
  const uint32_t* row = mat + r * pitch + c0;
  uint32_t s[16];
  if (first) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = 0;
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint4 v = state[k * h + r];
      s[4 * k] = v.x; s[4 * k + 1] = v.y; s[4 * k + 2] = v.z; s[4 * k + 3] = v.w;
    }
  }
  uint32_t c = 0, k = k0;  // k0 is uniform over the grid: the branches below do not diverge
  do {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i >= (int)k && c < nc) {
        s[i] = __ldg(row + c);
        c++;
        k = i + 1;
      }
    }
    if (k == 8 || (last && c == nc && k > 0)) {
      p2::permute(s);
      k = 0;
    }
  } while (c < nc);
  if (last) {
    store_digest(out + r * 8, s);
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) state[q * h + r] = make_uint4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
  }
}

// General case: the row is the concatenation of the rows of several matrices of any widths.
__global__ void __launch_bounds__(256) hash_rows_multi(const MatDesc* __restrict__ gm, uint32_t gn, uint64_t h,
                                                       uint32_t* __restrict__ out) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h) return;
  uint32_t s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = 0;
  uint32_t mi = 0, c = 0;
  while (mi < gn && gm[mi].w == 0) mi++;
  const uint32_t* row = mi < gn ? gm[mi].ptr + r * gm[mi].pitch : nullptr;
  while (mi < gn) {
    uint32_t got = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (mi < gn) {
        s[i] = __ldg(row + c);
        got++;
        c++;
        if (c == gm[mi].w) {
          c = 0;
          mi++;
          while (mi < gn && gm[mi].w == 0) mi++;
          if (mi < gn) row = gm[mi].ptr + r * gm[mi].pitch;
        }
      }
    }
    if (got) p2::permute(s);
  }
  store_digest(out + r * 8, s);
}

// cur[i] = C(prev[2i], prev[2i+1]), then (optionally) cur[i] = C(cur[i], inject[i]).
__global__ void __launch_bounds__(256) compress_layer(const uint32_t* __restrict__ prev, uint32_t* __restrict__ cur,
                                                      uint64_t n, const uint32_t* __restrict__ inject) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* p = reinterpret_cast<const uint4*>(prev + i * 16);
  uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
  p2::permute(s);
  if (inject) {
    const uint4* q = reinterpret_cast<const uint4*>(inject + i * 8);
    uint4 e = __ldg(q), f = __ldg(q + 1);
    s[8] = e.x; s[9] = e.y; s[10] = e.z; s[11] = e.w;
    s[12] = f.x; s[13] = f.y; s[14] = f.z; s[15] = f.w;
    p2::permute(s);
  }
  store_digest(cur + i * 8, s);
}

// Top of the tree in ONE launch: starting from a layer of `len0` <= 1024 digests at word offset off0 of the
// digest buffer (layers stored back to back, halving), a single CTA compresses layer after layer down to
// the root.  Replaces up to 10 launch-latency-bound compress_layer launches per tree (FRI commits 20 trees).
__global__ void __launch_bounds__(512) compress_top(uint32_t* __restrict__ digests, uint64_t off0, uint32_t len0) {
  uint64_t off = off0;
  for (uint32_t len = len0; len > 1; len >>= 1) {
    uint32_t half = len >> 1;
    uint64_t nxt = off + (uint64_t)len * 8;
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(digests + off + (uint64_t)i * 16);
      uint4 a = p[0], b = p[1], c = p[2], d = p[3];
      uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
      p2::permute(s);
      store_digest(digests + nxt + (uint64_t)i * 8, s);
    }
    __syncthreads();
    off = nxt;
  }
}

// Tail of a tree in ONE launch for up to 2^16 digests: CTA b compresses its segment of 1024 digests down to one (ten
// layers, written to their places in the back-to-back layer buffer), and the CTA that finishes LAST (device counter)
// compresses the gridDim.x remaining digests down to the root.  A tree of 2^16 leaves used to cost six compress_layer
// launches of 11-13 us each (one permutation deep, launch-latency bound) plus compress_top; every FRI layer commits a tree.
__global__ void __launch_bounds__(512) compress_tail(uint32_t* __restrict__ digests, uint64_t off0, uint32_t len0,
                                                     unsigned int* __restrict__ counter) {
  __shared__ unsigned int s_last;
  const uint32_t seg = len0 < 1024u ? len0 : 1024u;
  uint64_t off = off0;
  uint32_t len = len0, cnt = seg, base = blockIdx.x * seg;
  while (cnt > 1) {
    const uint32_t half = cnt >> 1;
    const uint64_t nxt = off + (uint64_t)len * 8;
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(digests + off + (uint64_t)(base + 2 * i) * 8);
      uint4 a = p[0], b = p[1], c = p[2], d = p[3];
      uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
      p2::permute(s);
      store_digest(digests + nxt + (uint64_t)((base >> 1) + i) * 8, s);
    }
    __syncthreads();
    off = nxt;
    len >>= 1;
    cnt = half;
    base >>= 1;
  }
  if (gridDim.x == 1) return;
  __threadfence();  // this CTA's digest is visible device-wide before its ticket is
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) *counter = 0u;  // ready for the next tree on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (len = gridDim.x; len > 1; len >>= 1) {
    const uint32_t half = len >> 1;
    const uint64_t nxt = off + (uint64_t)len * 8;
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
      const uint4* p = reinterpret_cast<const uint4*>(digests + off + (uint64_t)i * 16);
      uint4 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2), d = __ldcg(p + 3);  // written by other SMs: not through L1
      uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
      p2::permute(s);
      store_digest(digests + nxt + (uint64_t)i * 8, s);
    }
    __syncthreads();
    off = nxt;
  }
}

// Raw permutation of n independent 16-word states (unit entry point / known-answer tests).
__global__ void __launch_bounds__(256) permute_states(uint32_t* st, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[16];
#pragma unroll
  for (int k = 0; k < 16; k++) s[k] = st[i * 16 + k];
  p2::permute(s);
#pragma unroll
  for (int k = 0; k < 16; k++) st[i * 16 + k] = s[k];
}

}  // namespace mk
