// quotient.cu -- C ABI of the quotient stage: registry of generated AIR kernels and zk_quotient.
#include "zkgpu_internal.cuh"
#include "gen/airs_gen.cuh"

namespace quot {
// alpha_pows[k] = alpha^(n-1-k)  (powers_of_alpha reversed, crates/stark/src/prover.rs:453-456)
__global__ void alpha_pows_rev_kernel(const uint32_t* alpha, uint32_t n, uint32_t* out) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  kb::Ext a{{alpha[0], alpha[1], alpha[2], alpha[3]}};
  kb::Ext r = kb::ext_pow(a, n - 1 - k);
  out[4 * k + 0] = r.c[0];
  out[4 * k + 1] = r.c[1];
  out[4 * k + 2] = r.c[2];
  out[4 * k + 3] = r.c[3];
}

}  // namespace quot

extern "C" int32_t zk_air_count(void) { return quotgen::NUM_AIRS; }
extern "C" const char* zk_air_name(int32_t id) {
  return (id >= 0 && id < quotgen::NUM_AIRS) ? quotgen::AIRS[id].name : nullptr;
}
extern "C" int32_t zk_air_find(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < quotgen::NUM_AIRS; i++)
    if (!strcmp(name, quotgen::AIRS[i].name)) return i;
  return -1;
}
extern "C" int32_t zk_air_info(int32_t id, zk_air_desc* out) {
  if (id < 0 || id >= quotgen::NUM_AIRS || !out) return zk_fail(ZK_ERR_ARG, "unknown air id");
  const quotgen::Entry& e = quotgen::AIRS[id];
  out->main_width = e.main_w;
  out->prep_width = e.prep_w;
  out->perm_width = e.perm_w;
  out->num_public_values = e.n_pv;
  out->num_challenges = e.n_chal;
  out->num_constraints = e.n_constraints;
  out->max_degree = e.max_degree;
  out->num_kernels = e.n_parts;
  out->num_lookups = e.n_lookups;
  return ZK_OK;
}

static int32_t lde_of(const zk_pdata* pd, uint32_t idx, uint32_t want_w, uint64_t want_h, const char* what,
                      const uint32_t** ptr, uint32_t* pitch) {
  *ptr = nullptr;
  *pitch = want_w;
  if (want_w == 0) return ZK_OK;
  if (!pd || idx >= pd->n) return zk_fail(ZK_ERR_ARG, std::string(what) + " trace is required by this AIR");
  if (pd->widths[idx] != want_w) return zk_fail(ZK_ERR_ARG, std::string(what) + " trace has the wrong width");
  if (pd->heights[idx] < want_h) return zk_fail(ZK_ERR_ARG, std::string(what) + " LDE is shorter than the quotient domain");
  // the generated kernels fetch groups of adjacent columns with 64- / 128-bit loads (codegen.py `_vec_width`), which
  // the even row pitch of a committed LDE guarantees to be aligned
  if (pd->pitches[idx] != lde_pitch(want_w))
    return zk_fail(ZK_ERR_ARG, std::string(what) + " LDE has a dense odd row pitch (committed with ZK_EVEN_PITCH=0): the "
                                                   "quotient kernels need the even pitch");
  *ptr = pd->mats[idx];
  *pitch = pd->pitches[idx];
  return ZK_OK;
}

extern "C" int32_t zk_quotient(zk_ctx* c, int32_t air_id, const zk_pdata* prep, uint32_t prep_idx, const zk_pdata* main_pd,
                               uint32_t main_idx, const zk_pdata* perm, uint32_t perm_idx, uint32_t log_degree,
                               uint32_t log_quotient_degree, const uint32_t alpha[4], const uint32_t* perm_challenges,
                               const uint32_t* public_values, uint32_t n_public_values, const uint32_t local_cumsum[4],
                               const uint32_t global_cumsum[14], zk_dptr* out_chunks) {
  if (!c || !main_pd || !alpha || !out_chunks) return zk_fail(ZK_ERR_ARG, "null argument");
  if (air_id < 0 || air_id >= quotgen::NUM_AIRS) return zk_fail(ZK_ERR_ARG, "unknown air id");
  const quotgen::Entry& e = quotgen::AIRS[air_id];
  if (n_public_values < e.n_pv || (e.n_pv && !public_values)) return zk_fail(ZK_ERR_ARG, "too few public values");
  if (e.perm_w && !perm_challenges) return zk_fail(ZK_ERR_ARG, "permutation challenges are required");
  if (log_degree + log_quotient_degree > 23) return zk_fail(ZK_ERR_ARG, "quotient domain too large");
  uint64_t qsize = 1ull << (log_degree + log_quotient_degree);
  quot::Args A;
  memset(&A, 0, sizeof A);
  int32_t rc;
  if ((rc = lde_of(prep, prep_idx, e.prep_w, qsize, "preprocessed", &A.prep, &A.wp))) return rc;
  if ((rc = lde_of(main_pd, main_idx, e.main_w, qsize, "main", &A.main, &A.wm))) return rc;
  if ((rc = lde_of(perm, perm_idx, 4 * e.perm_w, qsize, "permutation", &A.perm, &A.wq))) return rc;
  A.log_n = log_degree;
  A.lqd = log_quotient_degree;
  A.g_q = kbh::two_adic_generator(log_degree + log_quotient_degree);
  A.g_n_inv = kbh::inv(kbh::two_adic_generator(log_degree));
  A.shift_pow_n = kbh::pow(kbh::GEN, 1ull << log_degree);
  A.w_lqd = kbh::two_adic_generator(log_quotient_degree);

  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "quotient");
  // small host inputs -> one device block: [alpha | chal | lcs | gcs | pvs]
  uint32_t n_chal = e.n_chal;
  std::vector<uint32_t> host(4 + 4 * n_chal + 4 + 14 + std::max(n_public_values, 1u), 0);
  memcpy(host.data(), alpha, 16);
  if (perm_challenges) memcpy(host.data() + 4, perm_challenges, 16 * n_chal);
  if (local_cumsum) memcpy(host.data() + 4 + 4 * n_chal, local_cumsum, 16);
  if (global_cumsum) memcpy(host.data() + 8 + 4 * n_chal, global_cumsum, 56);
  if (n_public_values) memcpy(host.data() + 22 + 4 * n_chal, public_values, 4ull * n_public_values);
  uint32_t *d_in = nullptr, *d_ap = nullptr, *d_out = nullptr;
  DevScope ds(c);  // frees on every return path; the output is released to the caller at the end
  if ((rc = ds.alloc(&d_in, host.size() * 4))) return rc;
  if ((rc = ds.alloc(&d_ap, std::max(e.n_constraints, 1u) * 16ull))) return rc;
  if ((rc = ds.alloc(&d_out, qsize * 16))) return rc;
  CK(cudaMemcpyAsync(d_in, host.data(), host.size() * 4, cudaMemcpyHostToDevice, c->stream));
  if (e.n_constraints) {
    ZK_LAUNCH(quot::alpha_pows_rev_kernel, (e.n_constraints + 127) / 128, 128, 0, c->stream, d_in, e.n_constraints, d_ap);
    c->launches++;
  }
  A.alpha_pows = d_ap;
  A.chal = d_in + 4;
  A.lcs = d_in + 4 + 4 * n_chal;
  A.gcs = d_in + 8 + 4 * n_chal;
  A.pvs = d_in + 22 + 4 * n_chal;
  A.out = d_out;
  if (e.n_parts == 0) CK(cudaMemsetAsync(d_out, 0, qsize * 16, c->stream));
  for (uint32_t p = 0; p < e.n_parts; p++) {
    auto kfn = e.parts[p];
    ZK_LAUNCH(kfn, (unsigned)((qsize + 127) / 128), 128, 0, c->stream, A);
    CK(cudaGetLastError());
    c->launches++;
  }
  ds.release(d_out);
  *out_chunks = (zk_dptr)d_out;
  return ZK_OK;
}

// generate_permutation_trace on the device (crates/stark/src/permutation.rs:102-196; call site
// crates/stark/src/prover.rs:341-364).  Traces are natural-order device matrices (zk_pdata_trace).
extern "C" int32_t zk_permutation_trace(zk_ctx* c, int32_t air_id, zk_dptr prep_trace, zk_dptr main_trace, uint64_t height,
                                        const uint32_t perm_challenges[8], zk_dptr* out_trace, uint32_t local_cumsum[4]) {
  if (!c || !main_trace || !perm_challenges || !out_trace || !local_cumsum) return zk_fail(ZK_ERR_ARG, "null argument");
  if (air_id < 0 || air_id >= quotgen::NUM_AIRS) return zk_fail(ZK_ERR_ARG, "unknown air id");
  const quotgen::Entry& e = quotgen::AIRS[air_id];
  if (!e.logup) return zk_fail(ZK_ERR_ARG, "this AIR has no lookups");
  if (e.prep_w && !prep_trace) return zk_fail(ZK_ERR_ARG, "preprocessed trace is required by this AIR");
  if (height == 0 || (height & (height - 1))) return zk_fail(ZK_ERR_ARG, "height must be a power of two");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "permutation_trace");
  logup::Args A;
  A.prep = (const uint32_t*)prep_trace;
  A.main = (const uint32_t*)main_trace;
  A.wp = e.prep_w;
  A.wm = e.main_w;
  A.wq = 4 * e.perm_w;
  A.h = height;
  uint32_t nblocks = (uint32_t)((height + logup::SCAN_T - 1) / logup::SCAN_T);
  uint32_t *d_chal = nullptr, *d_perm = nullptr, *d_rowsum = nullptr, *d_bsum = nullptr;
  int32_t rc;
  DevScope ds(c);
  if ((rc = ds.alloc(&d_chal, 48))) return rc;  // alpha, beta, [last]
  if ((rc = ds.alloc(&d_perm, height * A.wq * 4ull))) return rc;
  if ((rc = ds.alloc(&d_rowsum, height * 16))) return rc;
  if ((rc = ds.alloc(&d_bsum, nblocks * 16ull))) return rc;
  CK(cudaMemcpyAsync(d_chal, perm_challenges, 32, cudaMemcpyHostToDevice, c->stream));
  A.chal = d_chal;
  A.perm = d_perm;
  A.rowsum = d_rowsum;
  auto kfn = e.logup;
  ZK_LAUNCH(kfn, (unsigned)((height + 127) / 128), 128, 0, c->stream, A);
  const uint32_t col4 = A.wq - 4;
  ZK_LAUNCH_COOP(logup::scan_block_kernel, nblocks, logup::SCAN_T, 0, c->stream, d_rowsum, height, d_perm, A.wq, col4, d_bsum);
  ZK_LAUNCH_COOP(logup::scan_totals_kernel, 1, logup::SCAN_T, 0, c->stream, d_bsum, nblocks);
  ZK_LAUNCH(logup::scan_add_kernel, nblocks, logup::SCAN_T, 0, c->stream, d_perm, height, A.wq, col4, d_bsum, d_chal + 8);
  CK(cudaGetLastError());
  c->launches += 4;
  CK(cudaMemcpyAsync(local_cumsum, d_chal + 8, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  ds.release(d_perm);
  *out_trace = (zk_dptr)d_perm;
  return ZK_OK;
}
