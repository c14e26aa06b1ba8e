// zkgpu_internal.cuh -- shared declarations of the libzkgpu translation units (not part of the ABI).
#pragma once
#include <algorithm>
#include <map>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/zkgpu.h"
#include "kb31_host.h"
#include "launch.cuh"

int32_t zk_fail(int32_t code, const std::string& msg);

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      return zk_fail(ZK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                                      ":" + std::to_string(__LINE__) + ")");                              \
  } while (0)

struct zk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaMemPool_t pool = nullptr;         // private stream-ordered allocator of this context
  cudaStream_t copy_stream = nullptr;   // H2D of trace slabs, overlapped with compute on `stream`
  // upload slabs, shared by all matrices / calls (a third buffer was measured: no gain, the copy stream is not
  // held back by buffer reuse but by the PCIe rate of odd-pitch rows)
  static constexpr int NSLAB = 2;
  uint32_t* slab_buf[NSLAB] = {nullptr, nullptr};
  uint64_t slab_cap = 0;
  uint64_t slab_seq = 0;
  bool slab_used[NSLAB] = {false, false};
  cudaEvent_t slab_up[NSLAB] = {nullptr, nullptr}, slab_free[NSLAB] = {nullptr, nullptr};
  // Upload helper (zk_ctx_set_upload_helper): an IDLE peer GPU whose PCIe link carries the second half of the rows of
  // every slab; the half lands in a staging buffer there and is forwarded over NVLink (cudaMemcpyPeerAsync).
  int helper_dev = -1;
  cudaStream_t helper_stream[NSLAB] = {nullptr, nullptr};
  cudaEvent_t helper_done[NSLAB] = {nullptr, nullptr};
  uint32_t* helper_stage[NSLAB] = {nullptr, nullptr};
  uint64_t helper_cap = 0;
  uint64_t helper_min_bytes = 16ull << 20;  // smaller slabs go up directly
  uint32_t slab_cols = 0;               // fixed columns per slab (multiple of 16; env ZK_SLAB_COLS); 0 = by size
  uint64_t slab_bytes = 256ull << 20;   // target slab size of the streaming commit (env ZK_SLAB_MB)
  uint64_t hash_vec_min_rows = 1ull << 19;  // streamed sponge: vector-load kernel from this many rows (env ZK_HASH_VEC_MIN_ROWS)
  bool even_pitch = true;                   // committed LDEs of odd width get a padding column (env ZK_EVEN_PITCH=0: dense)
  uint64_t stream_min_bytes = 8ull << 20;  // smaller matrices go up in one piece (env ZK_STREAM_MIN_BYTES)
  // side streams of the opening reduction (fri.cu): the per-matrix chains of small kernels (row reduction, barycentric
  // partial sums, their finish, the reduced openings) of different matrices overlap; created on first use
  static constexpr int NSIDE = 4;
  cudaStream_t side[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t side_prod[NSIDE] = {nullptr, nullptr, nullptr, nullptr}, side_cons[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t side_fork = nullptr;
  int open_streams = 4;  // env ZK_OPEN_STREAMS (1 = everything on the context's stream)
  std::mutex mu;
  uint32_t log_L = 22;                    // twiddle table group; NTT sizes up to 2^22 rows (MAX_CPU_LOG_DEGREE, crates/core/machine/src/cpu/mod.rs:8)
  uint32_t* tw[2] = {nullptr, nullptr};   // g_L^(+e), g_L^(-e), e < 2^(L-1)
  unsigned int* tail_counter = nullptr;   // mk::compress_tail / fri tail kernels: "last CTA finishes" tickets
  bool keep_traces = false;
  bool prof = false;
  struct Rec {
    std::string name;
    cudaEvent_t a, b;
    uint64_t launches;
  };
  std::vector<Rec> recs;
  uint64_t launches = 0;
  // prover data handles hold a pointer to their context: zk_ctx_destroy with handles still alive only marks the
  // context and the LAST zk_pdata_free tears it down (bindings free handles from destructors in any order)
  uint64_t live_pdata = 0;
  bool destroy_requested = false;
};

// device allocations of one entry point, freed (stream-ordered) on every return path
struct DevScope {
  zk_ctx* c;
  std::vector<void*> ptrs;
  explicit DevScope(zk_ctx* c) : c(c) {}
  template <class T>
  int32_t alloc(T** p, uint64_t bytes);
  void release(void* p) {  // hand a pointer over to the caller
    for (auto& q : ptrs)
      if (q == p) q = nullptr;
  }
  ~DevScope();
};

struct ProfScope {
  zk_ctx* c;
  int idx = -1;
  ProfScope(zk_ctx* c, const char* name);
  ~ProfScope();
};

// word offset of digest layer l in a tree whose leaf layer has 2^L digests (layers stored leaves first)
__host__ __device__ inline uint64_t mmcs_layer_off(uint32_t L, uint32_t l) {
  return 8ull * ((2ull << L) - (2ull << (L - l)));
}

struct zk_open_desc {
  const uint32_t* ptr;
  uint32_t w;
  uint32_t pitch;  // row stride in words (>= w)
  uint32_t log_h;
  uint32_t off;
};

// Row pitch of a committed LDE: odd widths get one padding column, so that every row starts 8-byte aligned and the
// two-column NTT kernels, 64-bit loads of the opening reduction and >= 8-byte-aligned uploads apply to the 47-, 115-
// and 119-column chips of a real execution shard.  The padding column holds the transform of zeros (zeros).
inline uint32_t lde_pitch(uint32_t w) { return (w + 1u) & ~1u; }

struct zk_pdata {
  zk_ctx* ctx = nullptr;
  uint32_t n = 0;
  std::vector<uint64_t> heights;  // committed heights
  std::vector<uint32_t> widths;
  std::vector<uint32_t> pitches;  // row stride of mats[i] in words: lde_pitch(w) for committed LDEs, w otherwise
  std::vector<uint32_t*> mats;    // device
  std::vector<bool> owned;
  std::map<uint64_t, uint32_t*> class_digests;  // row digests of a height class hashed while its LDE streamed in
  std::vector<uint32_t*> traces;  // retained input traces (zk_ctx_keep_traces), else empty
  std::vector<bool> trace_owned;
  std::vector<uint32_t> order;    // indices by height descending, stable
  uint32_t log_max = 0;
  uint32_t sum_w = 0;
  uint32_t* digests = nullptr;    // all layers, leaves first
  std::vector<uint64_t> layer_off;  // word offset of each layer
  zk_open_desc* d_desc = nullptr;
  uint32_t root[8] = {0};
  bool counted = false;  // handed to the caller (counts towards ctx->live_pdata)
};

int32_t dev_alloc(zk_ctx* c, uint64_t bytes, void** out);
int32_t ensure_side_streams(zk_ctx* c);
int32_t dev_free(zk_ctx* c, void* p);
template <class T>
int32_t DevScope::alloc(T** p, uint64_t bytes) {
  void* q = nullptr;
  int32_t rc = dev_alloc(c, bytes, &q);
  if (rc == ZK_OK) ptrs.push_back(q);
  *p = (T*)q;
  return rc;
}
inline DevScope::~DevScope() {
  for (auto p : ptrs)
    if (p) cudaFreeAsync(p, c->stream);
}
// `out` has row pitch out_pitch >= w words (even when w is odd: see lde_pitch); `in` is dense (pitch w)
int32_t lde_dev(zk_ctx* c, const uint32_t* in, uint64_t h, uint32_t w, uint32_t log_blowup, uint32_t shift,
                uint32_t* out, uint32_t out_pitch);
int32_t mmcs_alloc(zk_ctx* c, zk_pdata* pd);
// with_open_desc == false skips the open_batch descriptor upload (FRI layers are opened by their own kernel),
// which keeps the build free of host->device copies and therefore fully asynchronous.
// progress of an incrementally built tree (streaming commit): next layer to compute, leaf state, injection scratch
struct TreeProgress {
  uint32_t next_l = 1;
  bool leaves = false;           // leaf layer complete
  bool leaves_streamed = false;  // the leaf digests are written by the streaming sponge, not by mmcs_advance
  uint32_t* inj = nullptr;
};
int32_t mmcs_advance(zk_ctx* c, zk_pdata* pd, TreeProgress& tp, const std::map<uint64_t, int>* pending);
int32_t mmcs_build(zk_ctx* c, zk_pdata* pd, bool fetch_root = true, bool leaves_done = false, bool with_open_desc = true,
                   TreeProgress* resume = nullptr);
// Mmcs::commit of one device-resident matrix; with fetch_root == false nothing is copied to the host and the
// stream is not synchronised (the root stays at pdata_root_dev()).
int32_t mmcs_commit_one_dev(zk_ctx* c, uint32_t* mat, uint64_t h, uint32_t w, bool take_ownership, bool fetch_root,
                            zk_pdata** out, bool with_open_desc = true);
inline const uint32_t* pdata_root_dev(const zk_pdata* pd) { return pd->digests + pd->layer_off[pd->log_max]; }
void pdata_release(zk_pdata* pd);
// open_batch gather for n_idx indices (each shifted right by `shift` first); query q writes its rows at
// opened + q * opened_stride and its path at proofs + q * proofs_stride (strides in words).
int32_t pdata_open_dev(zk_ctx* c, const zk_pdata* pd, uint32_t n_idx, const uint64_t* d_idx, uint32_t shift,
                       uint32_t* d_opened, uint64_t opened_stride, uint32_t* d_proofs, uint64_t proofs_stride);
