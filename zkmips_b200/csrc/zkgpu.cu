// zkgpu.cu -- libzkgpu.so: context, Pcs::commit path (coset LDE + Poseidon2 MMCS) and the C ABI of
// include/zkgpu.h.  sm_100a only; no CPU fallback.
#include "zkgpu_internal.cuh"
#include "merkle.cuh"
#include "kb31.cuh"
#include "ntt.cuh"

thread_local std::string g_last_error;

int32_t zk_fail(int32_t code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ------------------------------------------------------------------------------------------------
// small kernels local to this file
// ------------------------------------------------------------------------------------------------
namespace {

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


// out row r = in row bitrev(r)  (natural <-> bit-reversed order), one thread per word
__global__ void bitrev_rows_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t log_h,
                                   uint32_t w, uint64_t total) {
  uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  uint32_t r = (uint32_t)(gid / w), c = (uint32_t)(gid % w);
  uint32_t s = log_h ? (__brev(r) >> (32 - log_h)) : 0u;
  out[gid] = in[(size_t)s * w + c];
}

// Mmcs::open_batch gather: block b = query index b
__global__ void open_gather_kernel(const zk_open_desc* __restrict__ mats, uint32_t n_mats, uint32_t sum_w,
                                   const uint32_t* __restrict__ digests, uint32_t log_max, const uint64_t* __restrict__ indices, uint32_t shift,
                                   uint32_t* __restrict__ opened, uint64_t opened_stride,
                                   uint32_t* __restrict__ proofs, uint64_t proofs_stride) {
  uint64_t index = indices[blockIdx.x] >> shift;
  uint32_t* o = opened + (size_t)blockIdx.x * opened_stride;
  for (uint32_t m = 0; m < n_mats; m++) {
    zk_open_desc d = mats[m];
    uint64_t r = index >> (log_max - d.log_h);
    const uint32_t* row = d.ptr + r * d.pitch;
    for (uint32_t c = threadIdx.x; c < d.w; c += blockDim.x) o[d.off + c] = row[c];
  }
  uint32_t* p = proofs + (size_t)blockIdx.x * proofs_stride;
  for (uint32_t t = threadIdx.x; t < log_max * 8; t += blockDim.x) {
    uint32_t l = t >> 3, k = t & 7;
    p[t] = digests[mmcs_layer_off(log_max, l) + (((index >> l) ^ 1) << 3) + k];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" void zk_ctx_destroy(zk_ctx* c);
static int32_t ctx_init(zk_ctx* c) {
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major != 10)
    return zk_fail(ZK_ERR_CUDA, "libzkgpu is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                                    std::to_string(prop.minor));
  if (!c->stream) {
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (const char* e = getenv("ZK_SLAB_COLS")) {
    int v = atoi(e);
    if (v >= 16 && v % 16 == 0) c->slab_cols = (uint32_t)v;
  }
  if (const char* e = getenv("ZK_SLAB_MB")) {
    int v = atoi(e);
    if (v >= 1) c->slab_bytes = (uint64_t)v << 20;
  }
  if (const char* e = getenv("ZK_STREAM_MIN_BYTES")) c->stream_min_bytes = strtoull(e, nullptr, 10);
  if (const char* e = getenv("ZK_OPEN_STREAMS")) c->open_streams = std::max(1, std::min((int)zk_ctx::NSIDE, atoi(e)));
  if (const char* e = getenv("ZK_EVEN_PITCH")) c->even_pitch = atoi(e) != 0;  // 0: dense LDEs (A/B measurements)
  if (const char* e = getenv("ZK_HASH_VEC_MIN_ROWS")) c->hash_vec_min_rows = strtoull(e, nullptr, 10);
  // a private stream-ordered pool per context: contexts that prove shards concurrently on one GPU must not
  // couple their streams through cross-stream reuse of freed blocks in the device's default pool
  cudaMemPoolProps props;
  memset(&props, 0, sizeof props);
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = c->device;
  CK(cudaMemPoolCreate(&c->pool, &props));
  uint64_t thr = UINT64_MAX;
  CK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
  CK(ntt::configure_device());
  // global tables of g_L^(+-e)
  uint64_t half = 1ull << (c->log_L - 1);
  uint32_t gL = kbh::two_adic_generator(c->log_L);
  for (int d = 0; d < 2; d++) {
    CK(cudaMalloc(&c->tw[d], half * 4));
    ZK_LAUNCH(powers_kernel, (unsigned)((half + 255) / 256), 256, 0, c->stream, c->tw[d], half,
                                                                              d == 0 ? gL : kbh::inv(gL), kbh::ONE);
    CK(cudaGetLastError());
    c->launches++;
  }
  CK(cudaMalloc(&c->tail_counter, 64));  // ticket counters of the single-launch tree tails (reset by their last CTA)
  CK(cudaMemsetAsync(c->tail_counter, 0, 64, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

extern "C" int32_t zk_ctx_create_on_stream(int32_t device, void* stream, zk_ctx** out) {
  if (!out) return zk_fail(ZK_ERR_ARG, "out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return zk_fail(ZK_ERR_CUDA, std::string("no CUDA device (libzkgpu has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return zk_fail(ZK_ERR_ARG, "device index out of range");
  zk_ctx* c = new zk_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  int32_t rc = ctx_init(c);
  if (rc != ZK_OK) {
    std::string why = g_last_error;
    zk_ctx_destroy(c);  // releases whatever ctx_init got as far as creating
    g_last_error = why;
    return rc;
  }
  *out = c;
  return ZK_OK;
}
extern "C" int32_t zk_ctx_create(int32_t device, zk_ctx** out) { return zk_ctx_create_on_stream(device, nullptr, out); }

static void ctx_teardown(zk_ctx* c);
// side streams + their events (zkgpu_internal.cuh), created on first use
int32_t ensure_side_streams(zk_ctx* c) {
  if (c->side_fork) return ZK_OK;
  CK(cudaEventCreateWithFlags(&c->side_fork, cudaEventDisableTiming));
  for (int i = 0; i < zk_ctx::NSIDE; i++) {
    CK(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->side_prod[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->side_cons[i], cudaEventDisableTiming));
  }
  return ZK_OK;
}

// ---- upload helper -------------------------------------------------------------------------------------------
static void helper_release(zk_ctx* c) {
  if (c->helper_dev < 0) return;
  cudaSetDevice(c->helper_dev);
  for (int b = 0; b < zk_ctx::NSLAB; b++) {
    if (c->helper_stream[b]) {
      cudaStreamSynchronize(c->helper_stream[b]);
      cudaStreamDestroy(c->helper_stream[b]);
    }
    if (c->helper_done[b]) cudaEventDestroy(c->helper_done[b]);
    if (c->helper_stage[b]) cudaFree(c->helper_stage[b]);
    c->helper_stream[b] = nullptr;
    c->helper_done[b] = nullptr;
    c->helper_stage[b] = nullptr;
  }
  c->helper_cap = 0;
  c->helper_dev = -1;
  cudaSetDevice(c->device);
}
extern "C" int32_t zk_ctx_set_upload_helper(zk_ctx* c, int32_t device) {
  if (!c) return zk_fail(ZK_ERR_ARG, "ctx is null");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->copy_stream));
  CK(cudaStreamSynchronize(c->stream));
  helper_release(c);
  if (device < 0) return ZK_OK;
  int ndev = 0, a = 0, b = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device >= ndev || device == c->device) return zk_fail(ZK_ERR_ARG, "upload helper must be another visible GPU");
  CK(cudaDeviceCanAccessPeer(&a, c->device, device));
  CK(cudaDeviceCanAccessPeer(&b, device, c->device));
  if (!a || !b) return zk_fail(ZK_ERR_ARG, "no peer access between the context's GPU and the upload helper");
  cudaError_t e = cudaDeviceEnablePeerAccess(device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
  cudaGetLastError();
  CK(cudaSetDevice(device));
  e = cudaDeviceEnablePeerAccess(c->device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
    cudaSetDevice(c->device);
    CK(e);
  }
  cudaGetLastError();
  for (int k = 0; k < zk_ctx::NSLAB; k++) {
    cudaError_t e1 = cudaStreamCreateWithFlags(&c->helper_stream[k], cudaStreamNonBlocking);
    cudaError_t e2 = cudaEventCreateWithFlags(&c->helper_done[k], cudaEventDisableTiming);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      c->helper_dev = device;
      helper_release(c);
      return zk_fail(ZK_ERR_CUDA, "could not create the upload helper's streams");
    }
  }
  c->helper_dev = device;
  CK(cudaSetDevice(c->device));
  return ZK_OK;
}
// staging buffers on the helper GPU (grown on demand; synchronises the helper's streams)
static int32_t helper_ensure(zk_ctx* c, uint64_t bytes) {
  if (bytes <= c->helper_cap) return ZK_OK;
  cudaError_t e = cudaSetDevice(c->helper_dev);
  for (int b = 0; b < zk_ctx::NSLAB && e == cudaSuccess; b++) {
    e = cudaStreamSynchronize(c->helper_stream[b]);
    if (e == cudaSuccess && c->helper_stage[b]) e = cudaFree(c->helper_stage[b]);
    c->helper_stage[b] = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&c->helper_stage[b], bytes);
  }
  cudaSetDevice(c->device);
  CK(e);
  c->helper_cap = bytes;
  return ZK_OK;
}

extern "C" void zk_ctx_destroy(zk_ctx* c) {
  if (!c) return;
  {
    std::lock_guard<std::mutex> g(c->mu);
    if (c->live_pdata) {  // deferred: the last zk_pdata_free finishes the job
      c->destroy_requested = true;
      return;
    }
  }
  ctx_teardown(c);
}
static void ctx_teardown(zk_ctx* c) {
  cudaSetDevice(c->device);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& r : c->recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (int d = 0; d < 2; d++)
    if (c->tw[d]) cudaFree(c->tw[d]);
  if (c->tail_counter) cudaFree(c->tail_counter);
  for (int b = 0; b < zk_ctx::NSLAB; b++) {
    if (c->slab_buf[b]) cudaFree(c->slab_buf[b]);
    if (c->slab_up[b]) cudaEventDestroy(c->slab_up[b]);
    if (c->slab_free[b]) cudaEventDestroy(c->slab_free[b]);
  }
  helper_release(c);
  if (c->pool) cudaMemPoolDestroy(c->pool);
  for (int i = 0; i < zk_ctx::NSIDE; i++) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->side_prod[i]) cudaEventDestroy(c->side_prod[i]);
    if (c->side_cons[i]) cudaEventDestroy(c->side_cons[i]);
  }
  if (c->side_fork) cudaEventDestroy(c->side_fork);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" int32_t zk_ctx_sync(zk_ctx* c) {
  if (!c) return zk_fail(ZK_ERR_ARG, "ctx is null");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}
extern "C" const char* zk_last_error(void) { return g_last_error.c_str(); }
namespace ntt {
std::atomic<uint64_t>& tma_pass_counter() {
  static std::atomic<uint64_t> n{0};  // diagnostics; contexts on several host threads bump it
  return n;
}
}  // namespace ntt
extern "C" uint64_t zk_ntt_tma_passes(void) { return ntt::tma_pass_counter().load(); }
extern "C" const char* zk_build_info(void) { return "libzkgpu sm_100a " __DATE__ " " __TIME__; }

extern "C" int32_t zk_prof_enable(zk_ctx* c, int32_t enable) {
  if (!c) return zk_fail(ZK_ERR_ARG, "ctx is null");
  std::lock_guard<std::mutex> g(c->mu);
  c->prof = enable != 0;
  return ZK_OK;
}
extern "C" int32_t zk_prof_reset(zk_ctx* c) {
  if (!c) return zk_fail(ZK_ERR_ARG, "ctx is null");
  std::lock_guard<std::mutex> g(c->mu);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c->recs.clear();
  return ZK_OK;
}
extern "C" int32_t zk_prof_count(zk_ctx* c) { return c ? (int32_t)c->recs.size() : 0; }
extern "C" int32_t zk_prof_get(zk_ctx* c, int32_t i, char* name, int32_t cap, float* ms, uint64_t* launches) {
  if (!c || i < 0 || i >= (int32_t)c->recs.size()) return zk_fail(ZK_ERR_ARG, "bad profile record index");
  auto& r = c->recs[i];
  CK(cudaEventSynchronize(r.b));
  float t = 0;
  CK(cudaEventElapsedTime(&t, r.a, r.b));
  if (ms) *ms = t;
  if (launches) *launches = r.launches;
  if (name && cap > 0) {
    strncpy(name, r.name.c_str(), cap - 1);
    name[cap - 1] = 0;
  }
  return ZK_OK;
}
// start of record i on the ctx stream, in ms after the start of record 0 (a timeline of the stages)
extern "C" int32_t zk_prof_start(zk_ctx* c, int32_t i, float* ms_after_first) {
  if (!c || !ms_after_first || i < 0 || i >= (int32_t)c->recs.size()) return zk_fail(ZK_ERR_ARG, "bad profile record index");
  CK(cudaEventSynchronize(c->recs[i].a));
  CK(cudaEventElapsedTime(ms_after_first, c->recs[0].a, c->recs[i].a));
  return ZK_OK;
}
extern "C" uint64_t zk_launch_count(zk_ctx* c) { return c ? c->launches : 0; }

ProfScope::ProfScope(zk_ctx* c, const char* name) : c(c) {
  if (!c->prof) return;
  idx = (int)c->recs.size();
  zk_ctx::Rec r;
  r.name = name;
  r.launches = c->launches;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, c->stream);
  c->recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  auto& r = c->recs[idx];
  cudaEventRecord(r.b, c->stream);
  r.launches = c->launches - r.launches;
}

// ------------------------------------------------------------------------------------------------
// device memory
// ------------------------------------------------------------------------------------------------
int32_t dev_alloc(zk_ctx* c, uint64_t bytes, void** out) {
  if (bytes == 0) bytes = 4;
  CK(cudaMallocFromPoolAsync(out, bytes, c->pool, c->stream));
  return ZK_OK;
}
int32_t dev_free(zk_ctx* c, void* p) {
  if (p) CK(cudaFreeAsync(p, c->stream));
  return ZK_OK;
}
extern "C" int32_t zk_dev_alloc(zk_ctx* c, uint64_t bytes, zk_dptr* out) {
  if (!c || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  void* p = nullptr;
  int32_t rc = dev_alloc(c, bytes, &p);
  *out = (zk_dptr)p;
  return rc;
}
extern "C" int32_t zk_dev_free(zk_ctx* c, zk_dptr p) {
  if (!c) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  return dev_free(c, (void*)p);
}
extern "C" int32_t zk_h2d(zk_ctx* c, zk_dptr dst, const void* src, uint64_t bytes) {
  if (!c || (!src && bytes)) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync((void*)dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}
extern "C" int32_t zk_d2h(zk_ctx* c, void* dst, zk_dptr src, uint64_t bytes) {
  if (!c || (!dst && bytes)) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, (const void*)src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

// ------------------------------------------------------------------------------------------------
// LDE
// ------------------------------------------------------------------------------------------------
static inline uint32_t num_passes(uint32_t log_n) {
  uint32_t ks[8];
  return ntt::plan_passes(log_n, ks);
}

// coset_lde_batch(in, log_blowup, shift).bit_reverse_rows()  (SURVEY A.7).
// Block t of the output (rows [t*h, (t+1)*h)) is the size-h DFT, bit-reversed, of the coefficients
// scaled by (shift * g_{n+b}^bitrev_b(t))^i: 2^b independent coset transforms, no zero padding.

// scale vectors sigma_t^i / h of the 2^b coset blocks (shared by every column of a matrix)
static int32_t lde_scales(zk_ctx* c, uint64_t h, uint32_t log_blowup, uint32_t shift, bool aligned,
                          std::vector<ntt::CosetScale>& out) {
  uint32_t n = kbh::log2_exact(h);
  uint32_t gnb = kbh::two_adic_generator(n + log_blowup);
  uint32_t hinv = kbh::inv(kbh::to_monty((uint32_t)(h % kbh::P)));
  out.assign(1u << log_blowup, ntt::CosetScale{});
  for (uint32_t t = 0; t < (1u << log_blowup); t++) {
    out[t].sigma = kbh::mul(shift, kbh::pow(gnb, kbh::bitrev(t, log_blowup)));
    out[t].hinv = hinv;
    (void)aligned;
    if (ntt::first_pass_is_smem(n)) continue;  // a shared-memory first pass derives the scale itself
    uint32_t* v = nullptr;
    int32_t rc = dev_alloc(c, h * 4ull, (void**)&v);
    if (rc) return rc;
    out[t].vec = v;
    ZK_LAUNCH(powers_kernel, (unsigned)((h + 255) / 256), 256, 0, c->stream, v, h, out[t].sigma, hinv);
    CK(cudaGetLastError());
    c->launches++;
  }
  return ZK_OK;
}
static void free_scales(zk_ctx* c, std::vector<ntt::CosetScale>& s) {
  for (auto& x : s)
    if (x.vec) dev_free(c, const_cast<uint32_t*>(x.vec));
}

static int32_t check_lde_shape(zk_ctx* c, uint64_t h, uint32_t log_blowup) {
  if (h == 0 || (h & (h - 1))) return zk_fail(ZK_ERR_ARG, "height must be a power of two");
  uint32_t n = kbh::log2_exact(h);
  if (n > c->log_L) return zk_fail(ZK_ERR_ARG, "trace height above 2^22 is not supported");
  if (n + log_blowup > kbh::TWO_ADICITY) return zk_fail(ZK_ERR_ARG, "LDE height exceeds the two-adicity of the field");
  return ZK_OK;
}

// dst[r * pitch + col] = 0 for every row: the padding column of an odd-width matrix
__global__ void __launch_bounds__(256) zero_col_kernel(uint32_t* __restrict__ dst, uint32_t pitch, uint32_t col, uint64_t h) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < h) dst[r * pitch + col] = 0u;
}

// dst[r * dpitch + j] = j < nc ? src[r * nc + j] : 0: dense rows of an odd slab spread to the even pitch (one warp per row)
__global__ void __launch_bounds__(256) spread_rows_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                          uint32_t nc, uint32_t dpitch, uint64_t h) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < h; r += (uint64_t)gridDim.x * 8)
    for (uint32_t j = lane; j < dpitch; j += 32) dst[r * dpitch + j] = j < nc ? src[r * nc + j] : 0u;
}

// LDE of `nc` columns.  `coef` receives the (bit-reversed, unscaled) coefficients and may alias `in` (the
// inverse transform then runs in place); the 2^b blocks go to out.ptr + t*h*out.w, columns [out.c0, out.c0+nc).
// The inverse transform covers nc_inv columns and the coset transforms nc columns (nc_inv < nc when the caller has
// zeroed padding columns of `coef` itself).
static int32_t lde_cols(zk_ctx* c, ntt::Cols in, ntt::Cols coef, ntt::Cols out, uint32_t nc, uint64_t h,
                        uint32_t log_blowup, const std::vector<ntt::CosetScale>& scales, uint32_t nc_inv = ~0u) {
  if (nc == 0) return ZK_OK;
  if (nc_inv == ~0u) nc_inv = nc;
  uint32_t n = kbh::log2_exact(h);
  {
    ProfScope ps(c, "idft");
    CK(ntt::transform(in, coef, nc_inv, n, ntt::DIR_INV, c->tw[1], c->log_L, nullptr, false, c->stream));
    c->launches += num_passes(n);
  }
  for (uint32_t t = 0; t < (1u << log_blowup); t++) {
    ProfScope ps(c, "coset_dft");
    ntt::Cols blk{out.ptr + (size_t)t * h * out.w, out.w, out.c0};
    CK(ntt::transform(coef, blk, nc, n, ntt::DIR_FWD, c->tw[0], c->log_L, &scales[t], true, c->stream));
    c->launches += num_passes(n);
  }
  return ZK_OK;
}

int32_t lde_dev(zk_ctx* c, const uint32_t* in, uint64_t h, uint32_t w, uint32_t log_blowup, uint32_t shift,
                uint32_t* out, uint32_t out_pitch) {
  int32_t rc = check_lde_shape(c, h, log_blowup);
  if (rc) return rc;
  if (w == 0) return ZK_OK;
  if (out_pitch < w) return zk_fail(ZK_ERR_ARG, "output pitch below the width");
  DevScope ds(c);
  uint32_t* coef = nullptr;
  std::vector<ntt::CosetScale> scales;
  // the coefficient buffer takes the output's pitch: with an odd width the inverse transform reads the caller's dense
  // (odd-pitch) matrix one column per thread, but the 2^b coset transforms -- two thirds of the work -- run on even
  // pitches, padding column included (zero coefficients in, zeros out)
  const uint32_t ncf = std::min(out_pitch, lde_pitch(w));
  if ((rc = ds.alloc(&coef, h * (uint64_t)ncf * 4))) return rc;
  if (ncf > w) {
    ZK_LAUNCH(zero_col_kernel, (unsigned)((h + 255) / 256), 256, 0, c->stream, coef, ncf, w, h);
    CK(cudaGetLastError());
    c->launches++;
  }
  rc = lde_scales(c, h, log_blowup, shift, true, scales);
  if (rc == ZK_OK)
    rc = lde_cols(c, ntt::Cols{const_cast<uint32_t*>(in), w, 0}, ntt::Cols{coef, ncf, 0}, ntt::Cols{out, out_pitch, 0}, ncf,
                  h, log_blowup, scales, w);
  free_scales(c, scales);  // also after a partial failure of lde_scales
  return rc;
}

static int32_t hash_group(zk_ctx* c, const std::vector<mk::MatDesc>& g, uint64_t h, uint32_t* out);

// dst[r * dpitch + j] = src[r * spitch + j], j < nc: a column slab back into its place in the row-major retained trace
template <class T>
__global__ void __launch_bounds__(256) scatter_cols_kernel(const T* __restrict__ src, T* __restrict__ dst, uint32_t nc,
                                                           uint32_t spitch, uint32_t dpitch, uint64_t total) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / nc;
    uint32_t j = (uint32_t)(i - r * nc);
    dst[r * dpitch + j] = src[r * spitch + j];
  }
}

// (Re)allocates the context's slab buffers when a larger slab is needed (rare; synchronises).
static int32_t ensure_slab_bufs(zk_ctx* c, uint64_t bytes) {
  if (bytes <= c->slab_cap) return ZK_OK;
  CK(cudaStreamSynchronize(c->copy_stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int b = 0; b < zk_ctx::NSLAB; b++) {
    if (c->slab_buf[b]) CK(cudaFree(c->slab_buf[b]));
    c->slab_buf[b] = nullptr;
    CK(cudaMalloc(&c->slab_buf[b], bytes));
    c->slab_used[b] = false;
    if (!c->slab_up[b]) {
      CK(cudaEventCreateWithFlags(&c->slab_up[b], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&c->slab_free[b], cudaEventDisableTiming));
    }
  }
  c->slab_cap = bytes;
  return ZK_OK;
}

// Streaming LDE of a HOST matrix: column slabs flow  H2D (copy stream)  ||  inverse + coset transforms  ||
// resumable leaf sponge (compute stream), double buffered.  The slab buffer doubles as the coefficient
// buffer (in-place inverse transform), so no full-size staging copy of the trace exists on the device.
// When `cs` is non-null the rows of the LDE are absorbed slab by slab into the sponge of the matrix's height class
// (`last_member`: this is the last matrix of the class, so its last slab finalises the digests).

// Sponge of one height class of a streaming commit: the rows of the class are the concatenation of the rows of its
// matrices in input order; `pos` words of every row have been absorbed so far.
struct ClassStream {
  uint32_t* digests = nullptr;  // H x 8 words: the leaf layer or the injected digests of the class
  uint4* state = nullptr;       // H x 16 words between two absorb calls (allocated on demand)
  uint64_t pos = 0;
  int remaining = 0;            // matrices of the class not absorbed yet
  bool any = false;
};

static int32_t absorb_slab(zk_ctx* c, const uint32_t* lde, uint32_t pitch, uint32_t c0, uint32_t nc, uint64_t H,
                           ClassStream& cs, bool last) {
  ProfScope ps(c, "leaf_hash");
  const bool first = !cs.any;
  int32_t rc;
  if (!(first && last) && !cs.state && (rc = dev_alloc(c, H * 64, (void**)&cs.state))) return rc;
  const uint32_t k0 = (uint32_t)(cs.pos & 7);
  // fewer than ~3 resident CTAs of 256 threads per SM: 128-thread CTAs spread the rows evenly over the SMs
  const unsigned bs = H < (1ull << 19) ? 128 : 256;
  const unsigned blocks = (unsigned)((H + bs - 1) / bs);
  // the vector-load kernel needs block alignment; below 2^19 rows the scalar-load kernel is the faster one anyway
  // (5.58 vs 5.75 ms of leaf hashing on the 86 M-cell shard: fewer live registers across the permutation)
  const bool aligned = H >= c->hash_vec_min_rows && k0 == 0 && nc > 0 && nc % 8 == 0 && pitch % 8 == 0 && c0 % 8 == 0 &&
                       ((uintptr_t)lde % 32) == 0;
  if (aligned)
    ZK_LAUNCH(mk::hash_rows_slab, blocks, bs, 0, c->stream, lde, pitch, c0, nc >> 3, H, cs.state, (int)first, (int)last,
              cs.digests);
  else
    ZK_LAUNCH(mk::hash_rows_slab_any, blocks, bs, 0, c->stream, lde, pitch, c0, nc, H, cs.state, k0, (int)first, (int)last,
              cs.digests);
  CK(cudaGetLastError());
  c->launches++;
  cs.any = true;
  cs.pos += nc;
  return ZK_OK;
}

static int32_t lde_stream_host(zk_ctx* c, const uint32_t* host, uint64_t h, uint32_t w, uint32_t log_blowup, uint32_t shift,
                               uint32_t* out, uint32_t out_pitch, ClassStream* cs, bool last_member, uint32_t* keep_trace) {
  int32_t rc = check_lde_shape(c, h, log_blowup);
  if (rc) return rc;
  if (w == 0) return (cs && last_member) ? absorb_slab(c, nullptr, 0, 0, 0, h << log_blowup, *cs, true) : ZK_OK;
  if (!host) return zk_fail(ZK_ERR_ARG, "null matrix pointer");
  const uint64_t H = h << log_blowup;
  // Pageable host memory (a plain Rust Vec): strided 2-D copies from it crawl (measured 4 GB/s), so the trace goes
  // up in ONE contiguous copy and the device-resident path takes over.  Pinned / registered buffers stream.
  cudaPointerAttributes pa;
  bool pinned = cudaPointerGetAttributes(&pa, host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  cudaGetLastError();  // unregistered pointers may leave a sticky-less error code behind on old drivers
  if (!pinned && (uint64_t)h * w * 4 >= c->stream_min_bytes && c->slab_cols == 0) {
    uint32_t* stage = keep_trace;
    if (!stage && (rc = dev_alloc(c, h * w * 4ull, (void**)&stage))) return rc;
    CK(cudaMemcpyAsync(stage, host, h * w * 4ull, cudaMemcpyHostToDevice, c->stream));
    rc = lde_dev(c, stage, h, w, log_blowup, shift, out, out_pitch);
    if (rc == ZK_OK && cs) rc = absorb_slab(c, out, out_pitch, 0, w, H, *cs, last_member);
    if (!keep_trace) dev_free(c, stage);
    return rc;
  }
  // slab width: a fixed number of columns when ZK_SLAB_COLS is set, otherwise as many columns as make
  // ~slab_bytes (a multiple of 16 columns, >= 32): short-and-wide traces must not be cut into tiny slabs
  uint32_t slab = c->slab_cols;
  if (slab == 0) {
    // rows of a 2-D copy narrower than 256 B lose PCIe efficiency (measured: 64 B rows 43 ms, 128 B 24.8 ms,
    // 256 B 24.4 ms, 512 B 28.8 ms per GiB-sized trace), a trace should still be cut into >= 4 slabs to overlap
    uint64_t by_bytes = std::max<uint64_t>(64, c->slab_bytes / (h * 4) / 16 * 16);
    uint64_t quarter = std::max<uint64_t>(64, (w / 4 + 15) / 16 * 16);
    slab = (uint32_t)std::min(by_bytes, quarter);
  }
  if ((uint64_t)h * w * 4 < c->stream_min_bytes || w <= slab) slab = w;  // small matrices: one slab
  // Rows whose pitch is not a multiple of 128 bytes (47-, 115-, 119-column chips) make poor column slabs: the strided
  // copy of 224..256-byte rows out of a 476-byte pitch reaches 42 GB/s, and the log-21 execution shard spent 74 ms on
  // uploads that take 58 ms as linear copies (profiles/r2_timeline_exec_shard.txt).  Such a matrix goes up WHOLE, in one
  // linear copy, and the pipeline runs at matrix granularity instead (the two slab buffers belong to the context, so the
  // copy stream is already uploading the next matrix while this one is transformed and hashed).
  if (c->slab_cols == 0 && ((uint64_t)w * 4) % 128 != 0) slab = w;
  // Slab schedule.  The copy stream is the critical path (compute per slab is shorter than its upload), so what
  // is left after the LAST upload -- the transforms and the sponge of the last slab -- is pure tail.  The last
  // full slab is therefore cut in halves, down to 32 columns (128 B rows still copy at full PCIe rate, 64 B rows
  // do not): 256 columns go up as 64, 64, 64, 32, 32 and the tail is a 32-column slab instead of a 64-column one.
  std::vector<uint32_t> cuts;  // first column of every slab, then w
  if (c->slab_cols != 0 || slab >= w) {  // fixed slab width (ZK_SLAB_COLS), or one slab
    for (uint32_t c0 = 0; c0 < w; c0 += slab) cuts.push_back(c0);
  } else {
    // about w / slab slabs of EQUAL width (a multiple of 8 columns), never a narrow remainder: 68 columns cut as
    // 64 + 4 made the second upload a 2-D copy of 2^17 rows of 16 bytes -- 7.8 ms for 2 MB
    const uint32_t ns = std::max(1u, (w + slab / 2) / slab);
    const uint32_t base = ns > 1 ? std::max(8u, w / ns / 8 * 8) : w;
    for (uint32_t k = 0; k < ns; k++) cuts.push_back(k * base);
    uint32_t c0 = (ns - 1) * base, nc = w - c0;
    while (ns > 1 && nc >= 64 && nc % 32 == 0) {  // taper the last slab: halves, down to 32 columns
      nc /= 2;
      c0 += nc;
      cuts.push_back(c0);
    }
  }
  cuts.push_back(w);
  const uint32_t nslab = (uint32_t)cuts.size() - 1;
  std::vector<ntt::CosetScale> scales;
  const bool aligned = (w & 1u) == 0 && (slab & 1u) == 0 && ((uintptr_t)out % 8) == 0;
  if ((rc = lde_scales(c, h, log_blowup, shift, aligned, scales))) return rc;
  // the two slab buffers and their events live in the context and are shared by every matrix and every call:
  // the copy stream can therefore run ahead into the NEXT matrix while this one is still being transformed
  // sized by the WIDEST slab of the schedule, not by the nominal `slab`: equal-width cuts round down to a multiple of
  // 8 columns and hand the remainder to the last slab (300 columns -> 72, 72, 72, 84 with slab = 80), and a matrix
  // whose width rounds to one slab goes up whole (68 or 90 columns with slab = 64)
  // (cuts are multiples of 8 columns, so only the LAST slab of an odd-width matrix has an odd width; it gets one
  // padding column in the slab buffer -- zeroed, transformed along -- which lands in the padding column of the LDE)
  // An odd slab is uploaded DENSE into a staging region behind the slab (linear copy for whole rows, 2-D copy with a
  // dense destination otherwise -- the copy geometries whose PCIe rates are known, profiles/r1_h2d_probe.txt) and spread
  // to the even pitch by a kernel: a 2-D copy straight into the padded pitch (2^21 rows of 188 B into 192 B) is issued
  // row by row and took the execution-shard commit from 82 to 332 ms.
  uint32_t widest = 0, widest_odd = 0;
  for (uint32_t k = 0; k < nslab; k++) {
    const uint32_t nck = cuts[k + 1] - cuts[k];
    widest = std::max(widest, lde_pitch(nck));
    if (cuts[k] + lde_pitch(nck) <= out_pitch && lde_pitch(nck) > nck) widest_odd = std::max(widest_odd, nck);
  }
  const uint64_t stage_off = (h * (uint64_t)widest + 63) / 64 * 64;  // words; keeps the staging region 256-byte aligned
  if ((rc = ensure_slab_bufs(c, (stage_off + h * (uint64_t)widest_odd) * 4))) {
    free_scales(c, scales);
    return rc;
  }
  for (uint32_t k = 0; k < nslab && rc == ZK_OK; k++) {
    const uint32_t b = (uint32_t)(c->slab_seq++ % zk_ctx::NSLAB), c0 = cuts[k], nc = cuts[k + 1] - c0;
    const uint32_t ncp = (c0 + lde_pitch(nc) <= out_pitch) ? lde_pitch(nc) : nc;  // slab pitch = columns transformed
    uint32_t* buf = c->slab_buf[b];
    if (c->slab_used[b]) CK(cudaStreamWaitEvent(c->copy_stream, c->slab_free[b], 0));  // last reader of this buffer
    uint32_t* up = ncp > nc ? buf + stage_off : buf;  // dense upload target
    // With an upload helper the second half of the rows travels over the helper GPU's PCIe link into a staging buffer
    // there and is forwarded over NVLink; both halves are dense row ranges of the same slab buffer.
    uint64_t h_direct = h;
    const bool via_helper = c->helper_dev >= 0 && (uint64_t)h * nc * 4 >= c->helper_min_bytes;
    if (via_helper) {
      h_direct = h / 2;
      const uint64_t hr = h - h_direct, bytes = hr * nc * 4;
      if ((rc = helper_ensure(c, bytes))) break;
      cudaStream_t hs = c->helper_stream[b];
      uint32_t* stg = c->helper_stage[b];
      cudaError_t e = cudaSetDevice(c->helper_dev);
      if (e == cudaSuccess) {
        if (nc == w)
          e = cudaMemcpyAsync(stg, host + h_direct * w, bytes, cudaMemcpyHostToDevice, hs);
        else
          e = cudaMemcpy2DAsync(stg, (size_t)nc * 4, host + h_direct * w + c0, (size_t)w * 4, (size_t)nc * 4, hr,
                                cudaMemcpyHostToDevice, hs);
      }
      if (e == cudaSuccess && c->slab_used[b]) e = cudaStreamWaitEvent(hs, c->slab_free[b], 0);  // last reader of the slab buffer
      if (e == cudaSuccess) e = cudaMemcpyPeerAsync(up + h_direct * nc, c->device, stg, c->helper_dev, bytes, hs);
      if (e == cudaSuccess) e = cudaEventRecord(c->helper_done[b], hs);
      cudaSetDevice(c->device);
      CK(e);
    }
    if (nc == w)  // whole rows: one linear copy (a 2-D copy is issued row by row: 2^20 rows of 8 bytes take 2.7 ms)
      CK(cudaMemcpyAsync(up, host, (size_t)h_direct * w * 4, cudaMemcpyHostToDevice, c->copy_stream));
    else
      CK(cudaMemcpy2DAsync(up, (size_t)nc * 4, host + c0, (size_t)w * 4, (size_t)nc * 4, h_direct, cudaMemcpyHostToDevice,
                           c->copy_stream));
    CK(cudaEventRecord(c->slab_up[b], c->copy_stream));
    CK(cudaStreamWaitEvent(c->stream, c->slab_up[b], 0));
    if (via_helper) CK(cudaStreamWaitEvent(c->stream, c->helper_done[b], 0));
    if (ncp > nc) {  // dense rows -> even pitch, padding column zeroed
      const unsigned blocks = (unsigned)std::min<uint64_t>((h + 7) / 8, 148 * 16);
      ZK_LAUNCH(spread_rows_kernel, blocks, 256, 0, c->stream, up, buf, nc, ncp, h);
      CK(cudaGetLastError());
      c->launches++;
    }
    if (keep_trace) {  // retain the slab before the in-place inverse transform overwrites it
      if (nc == w && ncp == nc) {
        CK(cudaMemcpyAsync(keep_trace, buf, (size_t)h * w * 4, cudaMemcpyDeviceToDevice, c->stream));
      } else {
        // a kernel, not cudaMemcpy2DAsync: 2-D copies are issued row by row (2^20 rows of 8 bytes: 100 ms; the slabs
        // of a 3.2 GB execution shard: 22 ms), and uploading straight into the strided destination is slower still
        const bool v4 = ((nc | w | c0 | ncp) & 3u) == 0;
        const uint64_t total = (uint64_t)h * (v4 ? nc / 4 : nc);
        const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 32);
        if (v4)
          ZK_LAUNCH(scatter_cols_kernel<uint4>, blocks, 256, 0, c->stream, reinterpret_cast<const uint4*>(buf),
                    reinterpret_cast<uint4*>(keep_trace + c0), nc / 4, ncp / 4, w / 4, total);
        else
          ZK_LAUNCH(scatter_cols_kernel<uint32_t>, blocks, 256, 0, c->stream, buf, keep_trace + c0, nc, ncp, w, total);
        CK(cudaGetLastError());
        c->launches++;
      }
    }
    ntt::Cols sl{buf, ncp, 0};
    rc = lde_cols(c, sl, sl, ntt::Cols{out, out_pitch, c0}, ncp, h, log_blowup, scales);
    if (rc) break;
    CK(cudaEventRecord(c->slab_free[b], c->stream));
    c->slab_used[b] = true;
    if (cs && (rc = absorb_slab(c, out, out_pitch, c0, nc, H, *cs, last_member && k + 1 == nslab))) break;
  }
  free_scales(c, scales);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// MMCS
// ------------------------------------------------------------------------------------------------
static int32_t hash_group(zk_ctx* c, const std::vector<mk::MatDesc>& g, uint64_t h, uint32_t* out) {
  // fewer than ~3 resident CTAs of 256 threads per SM: use 128-thread CTAs so the rows spread evenly over the SMs
  const unsigned bs = h < (1ull << 19) ? 128 : 256;
  unsigned blocks = (unsigned)((h + bs - 1) / bs);
  if (g.size() == 1 && g[0].w % 8 == 0 && g[0].w > 0 && g[0].pitch == g[0].w && ((uintptr_t)g[0].ptr % 32) == 0) {
    ZK_LAUNCH(mk::hash_rows_w8, blocks, bs, 0, c->stream, g[0].ptr, g[0].w, h, out);
  } else {
    mk::MatDesc* d = nullptr;
    int32_t rc;
    if ((rc = dev_alloc(c, g.size() * sizeof(mk::MatDesc), (void**)&d))) return rc;
    CK(cudaMemcpyAsync(d, g.data(), g.size() * sizeof(mk::MatDesc), cudaMemcpyHostToDevice, c->stream));
    ZK_LAUNCH(mk::hash_rows_multi, blocks, bs, 0, c->stream, d, (uint32_t)g.size(), h, out);
    CK(cudaGetLastError());
    if ((rc = dev_free(c, d))) return rc;
  }
  CK(cudaGetLastError());
  c->launches++;
  return ZK_OK;
}

// Builds the digest layers over pd->mats (already on the device) and fills root.
// Layer geometry + digest buffer; separate from mmcs_build so that the streaming commit can hash leaves
// while the LDE is still being produced.
int32_t mmcs_alloc(zk_ctx* c, zk_pdata* pd) {
  uint32_t n = pd->n;
  pd->order.resize(n);
  for (uint32_t i = 0; i < n; i++) pd->order[i] = i;
  std::stable_sort(pd->order.begin(), pd->order.end(),
                   [&](uint32_t a, uint32_t b) { return pd->heights[a] > pd->heights[b]; });
  uint64_t hmax = pd->heights[pd->order[0]];
  pd->log_max = kbh::log2_exact(hmax);
  pd->layer_off.resize(pd->log_max + 1);
  uint64_t words = 0;
  for (uint32_t l = 0; l <= pd->log_max; l++) {
    pd->layer_off[l] = words;
    words += (hmax >> l) * 8;
  }
  return dev_alloc(c, words * 4, (void**)&pd->digests);
}

// Advances the digest layers as far as the available matrices allow.  `pending` maps a committed height to the
// number of its matrices whose LDE has not been enqueued yet (nullptr: everything is available).  The streaming
// commit calls this after every matrix, so the row digests of complete height classes and every tree layer that
// does not wait for a later (shorter) class run while the remaining traces are still being uploaded; what is
// left after the last upload is the last slab, the injection of its class and the layers above it.
int32_t mmcs_advance(zk_ctx* c, zk_pdata* pd, TreeProgress& tp, const std::map<uint64_t, int>* pending) {
  const uint32_t n = pd->n;
  int32_t rc;
  const uint64_t hmax = pd->heights[pd->order[0]], hmin = pd->heights[pd->order[n - 1]];
  auto avail = [&](uint64_t h) {
    if (!pending) return true;
    auto it = pending->find(h);
    return it == pending->end() || it->second == 0;
  };
  std::vector<mk::MatDesc> g;
  auto group = [&](uint64_t height) {
    g.clear();
    for (uint32_t k = 0; k < n; k++) {
      uint32_t m = pd->order[k];
      if (pd->heights[m] == height) g.push_back(mk::MatDesc{pd->mats[m], pd->widths[m], pd->pitches[m]});
    }
  };
  if (!tp.leaves) {
    if (!avail(hmax)) return ZK_OK;
    if (!tp.leaves_streamed) {
      ProfScope ps(c, "leaf_hash");
      group(hmax);
      if ((rc = hash_group(c, g, hmax, pd->digests))) return rc;
    }
    tp.leaves = true;
  }
  if (tp.next_l > pd->log_max) return ZK_OK;
  {
    group(hmax >> tp.next_l);
    if (!g.empty() && !avail(hmax >> tp.next_l)) return ZK_OK;  // nothing can be done yet: no empty profile record
  }
  ProfScope ps(c, "tree");
  while (tp.next_l <= pd->log_max) {
    const uint32_t l = tp.next_l;
    const uint64_t len = hmax >> l;
    if (hmin > len && 2 * len <= (1u << 16)) {
      // no matrix left to inject: ONE launch finishes the tree from the layer of 2*len digests (segments of 1024 per
      // CTA, the last CTA to finish does the top)
      ZK_LAUNCH_COOP(mk::compress_tail, (unsigned)std::max<uint64_t>(1, (2 * len) >> 10), 512, 0, c->stream, pd->digests,
                     pd->layer_off[l - 1], (uint32_t)(2 * len), c->tail_counter);
      CK(cudaGetLastError());
      c->launches++;
      tp.next_l = pd->log_max + 1;
      break;
    }
    group(len);
    const uint32_t* injp = nullptr;
    if (!g.empty()) {
      if (!avail(len)) break;
      auto it = pd->class_digests.find(len);
      if (it != pd->class_digests.end()) {
        injp = it->second;  // already hashed while the LDE streamed in
      } else {
        if (!tp.inj && (rc = dev_alloc(c, (hmax >> 1) * 32, (void**)&tp.inj))) return rc;
        if ((rc = hash_group(c, g, len, tp.inj))) return rc;
        injp = tp.inj;
      }
    }
    ZK_LAUNCH(mk::compress_layer, (unsigned)((len + 255) / 256), 256, 0, c->stream, pd->digests + pd->layer_off[l - 1],
              pd->digests + pd->layer_off[l], len, injp);
    CK(cudaGetLastError());
    c->launches++;
    tp.next_l++;
  }
  return ZK_OK;
}

int32_t mmcs_build(zk_ctx* c, zk_pdata* pd, bool fetch_root, bool leaves_done, bool with_open_desc, TreeProgress* resume) {
  uint32_t n = pd->n;
  int32_t rc;
  if (!pd->digests && (rc = mmcs_alloc(c, pd))) return rc;
  TreeProgress local;
  TreeProgress& tp = resume ? *resume : local;
  if (!resume) tp.leaves_streamed = leaves_done;
  rc = mmcs_advance(c, pd, tp, nullptr);
  if (tp.inj) {
    int32_t rc2 = dev_free(c, tp.inj);
    tp.inj = nullptr;
    if (rc == ZK_OK) rc = rc2;
  }
  if (rc) return rc;
  for (auto& kv : pd->class_digests) dev_free(c, kv.second);
  pd->class_digests.clear();
  if (with_open_desc) {
    // open_batch descriptors
    std::vector<zk_open_desc> od(n);
    uint32_t off = 0;
    for (uint32_t i = 0; i < n; i++) {
      od[i] = zk_open_desc{pd->mats[i], pd->widths[i], pd->pitches[i], kbh::log2_exact(pd->heights[i]), off};
      off += pd->widths[i];
    }
    pd->sum_w = off;
    if ((rc = dev_alloc(c, n * sizeof(zk_open_desc), (void**)&pd->d_desc))) return rc;
    CK(cudaMemcpyAsync(pd->d_desc, od.data(), n * sizeof(zk_open_desc), cudaMemcpyHostToDevice, c->stream));
  }
  if (fetch_root) {
    CK(cudaMemcpyAsync(pd->root, pd->digests + pd->layer_off[pd->log_max], 32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return ZK_OK;
}

int32_t mmcs_commit_one_dev(zk_ctx* c, uint32_t* mat, uint64_t h, uint32_t w, bool take_ownership, bool fetch_root,
                            zk_pdata** out, bool with_open_desc) {
  zk_pdata* pd = new zk_pdata();
  pd->ctx = c;
  pd->n = 1;
  pd->heights.assign(1, h);
  pd->widths.assign(1, w);
  pd->pitches.assign(1, w);
  pd->mats.assign(1, mat);
  pd->owned.assign(1, take_ownership);
  int32_t rc = mmcs_build(c, pd, fetch_root, false, with_open_desc);
  if (rc != ZK_OK) {
    pd->owned[0] = false;
    pdata_release(pd);
    return rc;
  }
  *out = pd;
  return ZK_OK;
}

static int32_t check_shapes(uint32_t n_mats, const void* ptrs, const uint64_t* heights, const uint32_t* widths) {
  if (n_mats == 0 || !ptrs || !heights || !widths) return zk_fail(ZK_ERR_ARG, "empty or null matrix list");
  for (uint32_t i = 0; i < n_mats; i++)
    if (heights[i] == 0 || (heights[i] & (heights[i] - 1)))
      return zk_fail(ZK_ERR_ARG, "matrix height must be a non-zero power of two");
  return ZK_OK;
}

void pdata_release(zk_pdata* pd) {
  zk_ctx* c = pd->ctx;
  cudaSetDevice(c->device);
  if (pd->counted) c->live_pdata--;
  for (uint32_t i = 0; i < pd->mats.size(); i++)
    if (pd->owned[i] && pd->mats[i]) cudaFreeAsync(pd->mats[i], c->stream);
  for (uint32_t i = 0; i < pd->traces.size(); i++)
    if (pd->trace_owned[i] && pd->traces[i]) cudaFreeAsync(pd->traces[i], c->stream);
  for (auto& kv : pd->class_digests) cudaFreeAsync(kv.second, c->stream);
  if (pd->digests) cudaFreeAsync(pd->digests, c->stream);
  if (pd->d_desc) cudaFreeAsync(pd->d_desc, c->stream);
  delete pd;
}

// common tail of the four commit entry points.  `src[i]` are DEVICE pointers to the input matrices, or HOST
// pointers when src_is_host (LDE case only: the streaming commit uploads them slab by slab).
static int32_t commit_common(zk_ctx* c, uint32_t n_mats, const uint32_t* const* src, bool src_is_host,
                             const uint64_t* heights, const uint32_t* widths, const uint32_t* domain_shifts, bool do_lde,
                             uint32_t log_blowup, uint32_t root[8], zk_pdata** out) {
  zk_pdata* pd = new zk_pdata();
  pd->ctx = c;
  pd->n = n_mats;
  pd->heights.resize(n_mats);
  pd->widths.assign(widths, widths + n_mats);
  pd->pitches.assign(widths, widths + n_mats);
  pd->mats.assign(n_mats, nullptr);
  pd->owned.assign(n_mats, false);
  pd->traces.assign(n_mats, nullptr);
  pd->trace_owned.assign(n_mats, false);
  int32_t rc = ZK_OK;
  bool leaves_done = false;
  TreeProgress tp;
  if (do_lde) {
    uint64_t hmax = 0;
    for (uint32_t i = 0; i < n_mats && rc == ZK_OK; i++) {
      pd->heights[i] = heights[i] << log_blowup;
      hmax = std::max(hmax, pd->heights[i]);
      pd->owned[i] = true;
      pd->pitches[i] = c->even_pitch ? lde_pitch(widths[i]) : widths[i];
      rc = dev_alloc(c, pd->heights[i] * pd->pitches[i] * 4ull, (void**)&pd->mats[i]);
      if (rc == ZK_OK && domain_shifts[i] == 0) rc = zk_fail(ZK_ERR_ARG, "domain shift must be non-zero");
    }
    if (rc == ZK_OK) rc = mmcs_alloc(c, pd);
    // Host traces: the row sponge of every height class is streamed with the LDE (leaf layer for the tallest class,
    // injected digests for the others); a class of several matrices is absorbed matrix after matrix, in input order.
    std::map<uint64_t, int> members;
    std::map<uint64_t, uint64_t> class_w;
    for (uint32_t i = 0; i < n_mats; i++) {
      members[pd->heights[i]]++;
      class_w[pd->heights[i]] += widths[i];
    }
    std::map<uint64_t, ClassStream> cls;
    if (src_is_host) {
      for (auto& kv : class_w) {
        if (rc != ZK_OK || kv.second == 0) continue;
        ClassStream st;
        st.remaining = members[kv.first];
        if (kv.first == hmax) {
          st.digests = pd->digests;
          leaves_done = true;
        } else {
          rc = dev_alloc(c, kv.first * 32, (void**)&st.digests);
          if (rc == ZK_OK) pd->class_digests[kv.first] = st.digests;
        }
        if (rc == ZK_OK) cls[kv.first] = st;
      }
    }
    std::map<uint64_t, int> pending = members;
    tp.leaves_streamed = leaves_done;
    // Device-resident sources (permutation traces, quotient chunks, generated traces): the transforms of DIFFERENT
    // matrices are independent, and most of these matrices are far too narrow to fill the GPU (ten quotient chunks of
    // 4 columns: ~8 launches of 10-15 us each), so matrix i runs on side stream i mod NSIDE and the context's stream
    // joins them before the tree is built.  lde_dev and everything below it enqueue on c->stream: the context mutex is
    // held, so the stream is swapped for the duration of the call.
    cudaStream_t main_stream = c->stream;
    const bool fan_out = !src_is_host && n_mats > 1 && c->open_streams > 1 && ensure_side_streams(c) == ZK_OK;
    bool side_used[zk_ctx::NSIDE] = {false, false, false, false};
    if (fan_out && cudaEventRecord(c->side_fork, main_stream) != cudaSuccess) rc = zk_fail(ZK_ERR_CUDA, "event record failed");
    for (uint32_t i = 0; i < n_mats && rc == ZK_OK; i++) {
      if (fan_out) {
        const int sl = (int)(i % (uint32_t)c->open_streams);
        if (!side_used[sl]) {
          cudaStreamWaitEvent(c->side[sl], c->side_fork, 0);
          side_used[sl] = true;
        }
        c->stream = c->side[sl];
      }
      uint32_t shift = kbh::mul(kbh::GEN, kbh::inv(domain_shifts[i]));  // GENERATOR / domain.shift
      uint32_t* keep = nullptr;
      if (c->keep_traces) {
        if (src_is_host) {
          rc = dev_alloc(c, heights[i] * widths[i] * 4ull, (void**)&keep);
          if (rc) break;
          pd->traces[i] = keep;
          pd->trace_owned[i] = true;
        } else {
          pd->traces[i] = const_cast<uint32_t*>(src[i]);  // borrowed: the caller keeps it alive
        }
      }
      if (src_is_host) {
        auto it = cls.find(pd->heights[i]);
        ClassStream* st = it == cls.end() ? nullptr : &it->second;
        rc = lde_stream_host(c, src[i], heights[i], widths[i], log_blowup, shift, pd->mats[i], pd->pitches[i], st,
                             st && st->remaining == 1, keep);
        if (st) st->remaining--;
      } else
        rc = lde_dev(c, src[i], heights[i], widths[i], log_blowup, shift, pd->mats[i], pd->pitches[i]);
      c->stream = main_stream;
      pending[pd->heights[i]]--;
      // host traces: hash complete height classes and build every tree layer that is already determined while
      // the copy stream is still uploading the remaining matrices
      if (rc == ZK_OK && src_is_host && i + 1 < n_mats) rc = mmcs_advance(c, pd, tp, &pending);
    }
    if (fan_out)
      for (int sl = 0; sl < zk_ctx::NSIDE; sl++)
        if (side_used[sl]) {
          cudaEventRecord(c->side_prod[sl], c->side[sl]);
          cudaStreamWaitEvent(main_stream, c->side_prod[sl], 0);
        }
    for (auto& kv : cls)
      if (kv.second.state) dev_free(c, kv.second.state);
  } else {
    for (uint32_t i = 0; i < n_mats; i++) {
      pd->heights[i] = heights[i];
      pd->mats[i] = const_cast<uint32_t*>(src[i]);
    }
  }
  if (rc == ZK_OK) rc = mmcs_build(c, pd, true, leaves_done, true, &tp);
  if (tp.inj) dev_free(c, tp.inj);
  if (rc != ZK_OK) {
    pdata_release(pd);
    return rc;
  }
  if (root) memcpy(root, pd->root, 32);
  pd->counted = true;
  c->live_pdata++;
  *out = pd;
  return ZK_OK;
}

// uploads host matrices; when keep != nullptr the uploads are handed to the pdata (no LDE case)
static int32_t upload_all(zk_ctx* c, uint32_t n_mats, const uint32_t* const* host, const uint64_t* heights,
                          const uint32_t* widths, std::vector<uint32_t*>& dev) {
  dev.assign(n_mats, nullptr);
  ProfScope ps(c, "h2d");
  for (uint32_t i = 0; i < n_mats; i++) {
    uint64_t bytes = heights[i] * widths[i] * 4ull;
    int32_t rc = dev_alloc(c, bytes, (void**)&dev[i]);
    if (rc) return rc;
    if (bytes) {
      if (!host[i]) return zk_fail(ZK_ERR_ARG, "null matrix pointer");
      CK(cudaMemcpyAsync(dev[i], host[i], bytes, cudaMemcpyHostToDevice, c->stream));
    }
  }
  return ZK_OK;
}

extern "C" int32_t zk_commit(zk_ctx* c, uint32_t n_mats, const uint32_t* const* mats_host, const uint64_t* heights,
                             const uint32_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup,
                             uint32_t root[8], zk_pdata** out) {
  if (!c || !out || !domain_shifts) return zk_fail(ZK_ERR_ARG, "null argument");
  int32_t rc = check_shapes(n_mats, mats_host, heights, widths);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  for (uint32_t i = 0; i < n_mats; i++)
    if (!mats_host[i] && heights[i] * widths[i]) return zk_fail(ZK_ERR_ARG, "null matrix pointer");
  return commit_common(c, n_mats, mats_host, true, heights, widths, domain_shifts, true, log_blowup, root, out);
}
extern "C" int32_t zk_commit_dev(zk_ctx* c, uint32_t n_mats, const zk_dptr* mats_dev, const uint64_t* heights,
                                 const uint32_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup,
                                 uint32_t root[8], zk_pdata** out) {
  if (!c || !out || !domain_shifts) return zk_fail(ZK_ERR_ARG, "null argument");
  int32_t rc = check_shapes(n_mats, mats_dev, heights, widths);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  std::vector<const uint32_t*> src(n_mats);
  for (uint32_t i = 0; i < n_mats; i++) src[i] = (const uint32_t*)mats_dev[i];
  return commit_common(c, n_mats, src.data(), false, heights, widths, domain_shifts, true, log_blowup, root, out);
}
extern "C" int32_t zk_mmcs_commit(zk_ctx* c, uint32_t n_mats, const uint32_t* const* mats_host, const uint64_t* heights,
                                  const uint32_t* widths, uint32_t root[8], zk_pdata** out) {
  if (!c || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  int32_t rc = check_shapes(n_mats, mats_host, heights, widths);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  std::vector<uint32_t*> dev;
  rc = upload_all(c, n_mats, mats_host, heights, widths, dev);
  if (rc == ZK_OK) rc = commit_common(c, n_mats, dev.data(), false, heights, widths, nullptr, false, 0, root, out);
  if (rc == ZK_OK) {
    for (uint32_t i = 0; i < n_mats; i++) (*out)->owned[i] = true;  // the uploads now belong to the pdata
  } else {
    for (auto p : dev) dev_free(c, p);
  }
  return rc;
}
extern "C" int32_t zk_mmcs_commit_dev(zk_ctx* c, uint32_t n_mats, const zk_dptr* mats_dev, const uint64_t* heights,
                                      const uint32_t* widths, uint32_t root[8], zk_pdata** out) {
  if (!c || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  int32_t rc = check_shapes(n_mats, mats_dev, heights, widths);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  std::vector<const uint32_t*> src(n_mats);
  for (uint32_t i = 0; i < n_mats; i++) src[i] = (const uint32_t*)mats_dev[i];
  return commit_common(c, n_mats, src.data(), false, heights, widths, nullptr, false, 0, root, out);
}

// ------------------------------------------------------------------------------------------------
// pdata accessors
// ------------------------------------------------------------------------------------------------
extern "C" void zk_pdata_free(zk_pdata* pd) {
  if (!pd) return;
  zk_ctx* c = pd->ctx;
  bool teardown = false;
  {
    std::lock_guard<std::mutex> g(c->mu);
    pdata_release(pd);
    teardown = c->destroy_requested && c->live_pdata == 0;
  }
  if (teardown) ctx_teardown(c);
}
extern "C" uint32_t zk_pdata_num_matrices(const zk_pdata* pd) { return pd ? pd->n : 0; }
extern "C" uint64_t zk_pdata_height(const zk_pdata* pd, uint32_t i) { return pd && i < pd->n ? pd->heights[i] : 0; }
extern "C" uint32_t zk_pdata_width(const zk_pdata* pd, uint32_t i) { return pd && i < pd->n ? pd->widths[i] : 0; }
extern "C" uint32_t zk_pdata_log_max_height(const zk_pdata* pd) { return pd ? pd->log_max : 0; }
extern "C" int32_t zk_pdata_root(const zk_pdata* pd, uint32_t root[8]) {
  if (!pd || !root) return zk_fail(ZK_ERR_ARG, "null argument");
  memcpy(root, pd->root, 32);
  return ZK_OK;
}
extern "C" zk_dptr zk_pdata_lde(const zk_pdata* pd, uint32_t i) { return pd && i < pd->n ? (zk_dptr)pd->mats[i] : 0; }
extern "C" zk_dptr zk_pdata_trace(const zk_pdata* pd, uint32_t i) {
  return pd && i < pd->traces.size() ? (zk_dptr)pd->traces[i] : 0;
}
extern "C" int32_t zk_ctx_keep_traces(zk_ctx* c, int32_t enable) {
  if (!c) return zk_fail(ZK_ERR_ARG, "ctx is null");
  std::lock_guard<std::mutex> g(c->mu);
  c->keep_traces = enable != 0;
  return ZK_OK;
}
extern "C" int32_t zk_pdata_copy_lde(const zk_pdata* pd, uint32_t i, uint32_t* out) {
  if (!pd || i >= pd->n || !out) return zk_fail(ZK_ERR_ARG, "bad argument");
  zk_ctx* c = pd->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  if (pd->pitches[i] == pd->widths[i])
    CK(cudaMemcpyAsync(out, pd->mats[i], pd->heights[i] * pd->widths[i] * 4ull, cudaMemcpyDeviceToHost, c->stream));
  else if (pd->widths[i])  // padded rows: the host image is dense
    CK(cudaMemcpy2DAsync(out, (size_t)pd->widths[i] * 4, pd->mats[i], (size_t)pd->pitches[i] * 4, (size_t)pd->widths[i] * 4,
                         pd->heights[i], cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}
extern "C" uint32_t zk_pdata_pitch(const zk_pdata* pd, uint32_t i) { return pd && i < pd->n ? pd->pitches[i] : 0; }
extern "C" int32_t zk_pdata_copy_layer(const zk_pdata* pd, uint32_t layer, uint32_t* out) {
  if (!pd || layer > pd->log_max || !out) return zk_fail(ZK_ERR_ARG, "bad argument");
  zk_ctx* c = pd->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint64_t len = (pd->heights[pd->order[0]] >> layer) * 32;
  CK(cudaMemcpyAsync(out, pd->digests + pd->layer_off[layer], len, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

// Mmcs::ProverData: Serialize + DeserializeOwned (crates/stark/src/prover.rs:221, machine.rs:56-57): the inverse of
// zk_pdata_copy_lde / zk_pdata_copy_layer.  Nothing is hashed -- like serde, it restores what was exported.
extern "C" int32_t zk_pdata_import(zk_ctx* c, uint32_t n_mats, const uint32_t* const* ldes_host, const uint64_t* heights,
                                   const uint32_t* widths, const uint32_t* const* layers_host, uint32_t n_layers,
                                   const uint32_t* const* traces_host, uint32_t log_blowup, zk_pdata** out) {
  if (!c || !out || !layers_host) return zk_fail(ZK_ERR_ARG, "null argument");
  int32_t rc = check_shapes(n_mats, ldes_host, heights, widths);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  zk_pdata* pd = new zk_pdata();
  pd->ctx = c;
  pd->n = n_mats;
  pd->heights.assign(heights, heights + n_mats);
  pd->widths.assign(widths, widths + n_mats);
  pd->pitches.assign(widths, widths + n_mats);
  pd->mats.assign(n_mats, nullptr);
  pd->owned.assign(n_mats, true);
  pd->traces.assign(n_mats, nullptr);
  pd->trace_owned.assign(n_mats, false);
  auto fail = [&](int32_t code) {
    pdata_release(pd);
    return code;
  };
  for (uint32_t i = 0; i < n_mats; i++) {
    const uint64_t bytes = heights[i] * widths[i] * 4ull;
    if (bytes && !ldes_host[i]) return fail(zk_fail(ZK_ERR_ARG, "null matrix pointer"));
    // same layout as a committed LDE: odd widths get the even row pitch (the generated quotient kernels and the opening
    // reduction read rows with vector loads), padding column zero
    pd->pitches[i] = c->even_pitch ? lde_pitch(widths[i]) : widths[i];
    if ((rc = dev_alloc(c, heights[i] * pd->pitches[i] * 4ull, (void**)&pd->mats[i]))) return fail(rc);
    if (pd->pitches[i] == widths[i]) {
      if (bytes && cudaMemcpyAsync(pd->mats[i], ldes_host[i], bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return fail(zk_fail(ZK_ERR_CUDA, "upload of an LDE matrix failed"));
    } else {
      uint32_t* dense = nullptr;
      if ((rc = dev_alloc(c, bytes, (void**)&dense))) return fail(rc);
      bool ok = cudaMemcpyAsync(dense, ldes_host[i], bytes, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
      if (ok) {
        const unsigned blocks = (unsigned)std::min<uint64_t>((heights[i] + 7) / 8, 148 * 16);
        ZK_LAUNCH(spread_rows_kernel, blocks, 256, 0, c->stream, dense, pd->mats[i], widths[i], pd->pitches[i], heights[i]);
        ok = cudaGetLastError() == cudaSuccess;
        c->launches++;
      }
      dev_free(c, dense);
      if (!ok) return fail(zk_fail(ZK_ERR_CUDA, "upload of an LDE matrix failed"));
    }
    if (traces_host && traces_host[i]) {
      if (heights[i] >> log_blowup == 0) return fail(zk_fail(ZK_ERR_ARG, "committed height below the blowup"));
      const uint64_t tb = (heights[i] >> log_blowup) * widths[i] * 4ull;
      if ((rc = dev_alloc(c, tb, (void**)&pd->traces[i]))) return fail(rc);
      pd->trace_owned[i] = true;
      if (tb && cudaMemcpyAsync(pd->traces[i], traces_host[i], tb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
        return fail(zk_fail(ZK_ERR_CUDA, "upload of a retained trace failed"));
    }
  }
  if ((rc = mmcs_alloc(c, pd))) return fail(rc);
  if (n_layers != pd->log_max + 1) return fail(zk_fail(ZK_ERR_ARG, "expected log_max_height + 1 digest layers"));
  const uint64_t hmax = pd->heights[pd->order[0]];
  for (uint32_t l = 0; l <= pd->log_max; l++) {
    if (!layers_host[l]) return fail(zk_fail(ZK_ERR_ARG, "null digest layer"));
    if (cudaMemcpyAsync(pd->digests + pd->layer_off[l], layers_host[l], (hmax >> l) * 32, cudaMemcpyHostToDevice,
                        c->stream) != cudaSuccess)
      return fail(zk_fail(ZK_ERR_CUDA, "upload of a digest layer failed"));
  }
  // open_batch descriptors and root, as mmcs_build's tail
  std::vector<zk_open_desc> od(n_mats);
  uint32_t off = 0;
  for (uint32_t i = 0; i < n_mats; i++) {
    od[i] = zk_open_desc{pd->mats[i], pd->widths[i], pd->pitches[i], kbh::log2_exact(pd->heights[i]), off};
    off += pd->widths[i];
  }
  pd->sum_w = off;
  if ((rc = dev_alloc(c, n_mats * sizeof(zk_open_desc), (void**)&pd->d_desc))) return fail(rc);
  if (cudaMemcpyAsync(pd->d_desc, od.data(), n_mats * sizeof(zk_open_desc), cudaMemcpyHostToDevice, c->stream) !=
          cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess)
    return fail(zk_fail(ZK_ERR_CUDA, "upload of the open_batch descriptors failed"));
  memcpy(pd->root, layers_host[pd->log_max], 32);
  pd->counted = true;
  c->live_pdata++;
  *out = pd;
  return ZK_OK;
}

int32_t pdata_open_dev(zk_ctx* c, const zk_pdata* pd, uint32_t n_idx, const uint64_t* d_idx, uint32_t shift,
                       uint32_t* d_opened, uint64_t opened_stride, uint32_t* d_proofs, uint64_t proofs_stride) {
  ZK_LAUNCH(open_gather_kernel, n_idx, 128, 0, c->stream, pd->d_desc, pd->n, pd->sum_w, pd->digests, pd->log_max, d_idx, shift, d_opened, opened_stride, d_proofs, proofs_stride);
  CK(cudaGetLastError());
  c->launches++;
  return ZK_OK;
}

extern "C" int32_t zk_pdata_open_batch(const zk_pdata* pd, uint32_t n_idx, const uint64_t* indices, uint32_t* opened,
                                       uint32_t* proofs) {
  if (!pd || !indices || !opened || (!proofs && pd->log_max)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n_idx == 0) return ZK_OK;
  uint64_t hmax = 1ull << pd->log_max;
  for (uint32_t i = 0; i < n_idx; i++)
    if (indices[i] >= hmax) return zk_fail(ZK_ERR_ARG, "open index out of range");
  zk_ctx* c = pd->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint64_t* d_idx = nullptr;
  uint32_t *d_op = nullptr, *d_pr = nullptr;
  int32_t rc;
  DevScope ds(c);
  uint64_t ob = (uint64_t)n_idx * pd->sum_w * 4, pb = (uint64_t)n_idx * pd->log_max * 32;
  if ((rc = ds.alloc(&d_idx, n_idx * 8ull))) return rc;
  if ((rc = ds.alloc(&d_op, ob))) return rc;
  if ((rc = ds.alloc(&d_pr, pb))) return rc;
  CK(cudaMemcpyAsync(d_idx, indices, n_idx * 8ull, cudaMemcpyHostToDevice, c->stream));
  if ((rc = pdata_open_dev(c, pd, n_idx, d_idx, 0, d_op, pd->sum_w, d_pr, (uint64_t)pd->log_max * 8))) return rc;
  if (ob) CK(cudaMemcpyAsync(opened, d_op, ob, cudaMemcpyDeviceToHost, c->stream));
  if (pb) CK(cudaMemcpyAsync(proofs, d_pr, pb, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

// ------------------------------------------------------------------------------------------------
// unit-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" int32_t zk_poseidon2_permute(zk_ctx* c, uint32_t* states, uint64_t n) {
  if (!c || (!states && n)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint32_t* d = nullptr;
  int32_t rc;
  DevScope ds(c);
  if ((rc = ds.alloc(&d, n * 64))) return rc;
  CK(cudaMemcpyAsync(d, states, n * 64, cudaMemcpyHostToDevice, c->stream));
  ZK_LAUNCH(mk::permute_states, (unsigned)((n + 255) / 256), 256, 0, c->stream, d, n);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(states, d, n * 64, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

extern "C" int32_t zk_hash_rows(zk_ctx* c, const uint32_t* mat, uint64_t h, uint32_t w, uint32_t* digests) {
  if (!c || !digests || (!mat && h * w)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (h == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint32_t *d = nullptr, *o = nullptr;
  int32_t rc;
  DevScope ds(c);
  if ((rc = ds.alloc(&d, h * w * 4ull))) return rc;
  if ((rc = ds.alloc(&o, h * 32))) return rc;
  if (h * w) CK(cudaMemcpyAsync(d, mat, h * w * 4ull, cudaMemcpyHostToDevice, c->stream));
  std::vector<mk::MatDesc> grp{mk::MatDesc{d, w, w}};
  if ((rc = hash_group(c, grp, h, o))) return rc;
  CK(cudaMemcpyAsync(digests, o, h * 32, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

extern "C" int32_t zk_compress_layer(zk_ctx* c, const uint32_t* prev, uint64_t n_out, uint32_t* out) {
  if (!c || !prev || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n_out == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint32_t *d = nullptr, *o = nullptr;
  int32_t rc;
  DevScope ds(c);
  if ((rc = ds.alloc(&d, n_out * 64))) return rc;
  if ((rc = ds.alloc(&o, n_out * 32))) return rc;
  CK(cudaMemcpyAsync(d, prev, n_out * 64, cudaMemcpyHostToDevice, c->stream));
  ZK_LAUNCH(mk::compress_layer, (unsigned)((n_out + 255) / 256), 256, 0, c->stream, d, o, n_out, nullptr);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(out, o, n_out * 32, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

extern "C" int32_t zk_coset_lde_dev(zk_ctx* c, zk_dptr in, uint64_t h, uint32_t w, uint32_t log_blowup, uint32_t shift,
                                    zk_dptr out) {
  if (!c || !in || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  return lde_dev(c, (const uint32_t*)in, h, w, log_blowup, shift, (uint32_t*)out, w);
}

extern "C" int32_t zk_coset_lde(zk_ctx* c, const uint32_t* in, uint64_t h, uint32_t w, uint32_t log_blowup,
                                uint32_t shift, uint32_t* out) {
  if (!c || !in || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint32_t *d = nullptr, *o = nullptr;
  int32_t rc;
  uint64_t ib = h * w * 4ull, obytes = ib << log_blowup;
  DevScope ds(c);
  if ((rc = ds.alloc(&d, ib))) return rc;
  if ((rc = ds.alloc(&o, obytes))) return rc;
  CK(cudaMemcpyAsync(d, in, ib, cudaMemcpyHostToDevice, c->stream));
  if ((rc = lde_dev(c, d, h, w, log_blowup, shift, o, w))) return rc;
  CK(cudaMemcpyAsync(out, o, obytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}

extern "C" int32_t zk_dft_batch(zk_ctx* c, const uint32_t* in, uint64_t h, uint32_t w, uint32_t* out) {
  if (!c || !in || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  if (h == 0 || (h & (h - 1))) return zk_fail(ZK_ERR_ARG, "height must be a power of two");
  uint32_t n = kbh::log2_exact(h);
  if (n > c->log_L) return zk_fail(ZK_ERR_ARG, "height above 2^22 is not supported");
  if (w == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  uint32_t *d = nullptr, *o = nullptr;
  int32_t rc;
  uint64_t bytes = h * w * 4ull;
  DevScope ds(c);
  if ((rc = ds.alloc(&d, bytes))) return rc;
  if ((rc = ds.alloc(&o, bytes))) return rc;
  CK(cudaMemcpyAsync(d, in, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(ntt::transform(ntt::Cols{d, w, 0}, ntt::Cols{d, w, 0}, w, n, ntt::DIR_FWD, c->tw[0], c->log_L, nullptr, false, c->stream));
  uint64_t total = h * w;
  ZK_LAUNCH(bitrev_rows_kernel, (unsigned)((total + 255) / 256), 256, 0, c->stream, d, o, n, w, total);
  CK(cudaGetLastError());
  c->launches += num_passes(n) + 1;
  CK(cudaMemcpyAsync(out, o, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ZK_OK;
}
