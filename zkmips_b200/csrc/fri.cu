// fri.cu -- host orchestration of TwoAdicFriPcs::open on the device (kernels in fri.cuh) and the
// device-challenger entry points of include/zkgpu.h.
#include "zkgpu_internal.cuh"
#include "fri.cuh"

static_assert(sizeof(zk_challenger) == sizeof(fri::Chal), "challenger images must match");

namespace {

struct Shape {
  uint32_t n_rounds = 0, log_max = 0, n_layers = 0, max_w = 0;
  uint64_t opened_words = 0, query_words = 0, total_words = 0;
  std::vector<uint64_t> round_off;  // offset of each round's block inside one query record
  std::vector<uint64_t> layer_q_off;  // offset of each commit-phase step inside one query record
};

// proof layout: see include/zkgpu.h ("flat proof layout")
Shape shape_of(uint32_t n_rounds, const zk_pdata* const* rounds, const uint32_t* n_points, uint32_t log_blowup,
               uint32_t num_queries) {
  Shape s;
  s.n_rounds = n_rounds;
  uint32_t k = 0;
  uint64_t per_query = 0;
  for (uint32_t r = 0; r < n_rounds; r++) {
    const zk_pdata* pd = rounds[r];
    s.round_off.push_back(per_query);
    for (uint32_t m = 0; m < pd->n; m++, k++) {
      s.opened_words += (uint64_t)n_points[k] * pd->widths[m] * 4;
      s.max_w = std::max(s.max_w, pd->widths[m]);
    }
    s.log_max = std::max(s.log_max, pd->log_max);
    per_query += pd->sum_w + (uint64_t)pd->log_max * 8;
  }
  s.n_layers = s.log_max >= log_blowup ? s.log_max - log_blowup : 0;
  for (uint32_t i = 0; i < s.n_layers; i++) {
    s.layer_q_off.push_back(per_query);
    per_query += 4 + (uint64_t)(s.log_max - i - 1) * 8;
  }
  s.query_words = per_query;
  s.total_words = s.opened_words + (uint64_t)s.n_layers * 8 + 4 + 1 + (uint64_t)num_queries * per_query;
  return s;
}

struct Scratch {  // frees everything it allocated when it goes out of scope
  zk_ctx* c;
  std::vector<void*> ptrs;
  std::vector<zk_pdata*> pds;
  explicit Scratch(zk_ctx* c) : c(c) {}
  template <class T>
  int32_t alloc(T** p, uint64_t bytes) {
    void* q = nullptr;
    int32_t rc = dev_alloc(c, bytes, &q);
    if (rc == ZK_OK) ptrs.push_back(q);
    *p = (T*)q;
    return rc;
  }
  ~Scratch() {
    for (auto pd : pds) pdata_release(pd);
    for (auto p : ptrs) cudaFreeAsync(p, c->stream);
  }
};

}  // namespace

extern "C" uint64_t zk_pcs_proof_words(uint32_t n_rounds, const zk_pdata* const* rounds, const uint32_t* n_points,
                                       uint32_t log_blowup, uint32_t num_queries) {
  if (!rounds || !n_points) return 0;
  return shape_of(n_rounds, rounds, n_points, log_blowup, num_queries).total_words;
}

#define RC(x)                    \
  do {                           \
    int32_t rc__ = (x);          \
    if (rc__ != ZK_OK) return rc__; \
  } while (0)

extern "C" int32_t zk_pcs_open(zk_ctx* c, uint32_t n_rounds, const zk_pdata* const* rounds, const uint32_t* n_points,
                               const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                               zk_challenger* ch, int64_t inject_witness, uint32_t* proof_host, uint64_t proof_cap) {
  if (!c || !rounds || !n_points || !ch || !proof_host || n_rounds == 0) return zk_fail(ZK_ERR_ARG, "null argument");
  if (pow_bits > 30) return zk_fail(ZK_ERR_ARG, "proof_of_work_bits must be <= 30");
  uint32_t total_pts = 0, kk = 0;
  for (uint32_t r = 0; r < n_rounds; r++) {
    if (!rounds[r] || rounds[r]->ctx != c) return zk_fail(ZK_ERR_ARG, "round data belongs to another context");
    for (uint32_t m = 0; m < rounds[r]->n; m++, kk++) {
      if (rounds[r]->heights[m] < (1ull << log_blowup)) return zk_fail(ZK_ERR_ARG, "committed height below the blowup");
      total_pts += n_points[kk];
    }
  }
  if (total_pts && !points) return zk_fail(ZK_ERR_ARG, "points is null");
  Shape S = shape_of(n_rounds, rounds, n_points, log_blowup, num_queries);
  if (proof_cap < S.total_words) return zk_fail(ZK_ERR_ARG, "proof buffer too small (see zk_pcs_proof_words)");

  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  Scratch sc(c);
  ProfScope ps_all(c, "pcs_open");

  uint32_t *d_proof, *d_pts, *d_alpha, *d_apow, *d_apow_split, *d_status, *d_found, *d_beta, *d_red, *d_aoff, *d_rowred, *d_partial;
  uint64_t* d_idx;
  fri::Chal* d_ch;
  uint64_t Hmax = 1ull << S.log_max;
  uint32_t nchunks_max = (uint32_t)(((Hmax >> log_blowup) + fri::BARY_ROWS - 1) / fri::BARY_ROWS);
  RC(sc.alloc(&d_proof, S.total_words * 4));
  RC(sc.alloc(&d_pts, (uint64_t)std::max(total_pts, 1u) * 16));
  RC(sc.alloc(&d_alpha, 16));
  RC(sc.alloc(&d_apow, (uint64_t)std::max(S.max_w, 1u) * 16));
  RC(sc.alloc(&d_apow_split, (uint64_t)std::max(S.max_w, 1u) * 32));
  if (S.max_w > 65536) return zk_fail(ZK_ERR_ARG, "matrix width above 65536 is not supported by the opening reduction");
  RC(sc.alloc(&d_status, 4));
  RC(sc.alloc(&d_found, 4));
  RC(sc.alloc(&d_beta, 16));
  // the opening reduction runs the per-matrix kernel chains on up to NSIDE streams: one set of scratch per slot
  const int nslots = c->open_streams;
  uint32_t *s_red[zk_ctx::NSIDE], *s_aoff[zk_ctx::NSIDE], *s_rowred[zk_ctx::NSIDE], *s_partial[zk_ctx::NSIDE];
  for (int i = 0; i < nslots; i++) {
    RC(sc.alloc(&s_red[i], 32));
    RC(sc.alloc(&s_aoff[i], 32));
    RC(sc.alloc(&s_rowred[i], Hmax * 16));
    RC(sc.alloc(&s_partial[i], (uint64_t)nchunks_max * 2 * std::max(S.max_w, 1u) * 16));
  }
  d_red = s_red[0]; d_aoff = s_aoff[0]; d_rowred = s_rowred[0]; d_partial = s_partial[0];
  (void)d_red; (void)d_aoff; (void)d_rowred; (void)d_partial;
  if (nslots > 1) RC(ensure_side_streams(c));
  RC(sc.alloc(&d_idx, (uint64_t)std::max(num_queries, 1u) * 8));
  RC(sc.alloc(&d_ch, sizeof(fri::Chal)));
  CK(cudaMemcpyAsync(d_ch, ch, sizeof(fri::Chal), cudaMemcpyHostToDevice, st));
  if (total_pts) CK(cudaMemcpyAsync(d_pts, points, (uint64_t)total_pts * 16, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_status, 0, 4, st));
  CK(cudaMemsetAsync(d_found, 0xff, 4, st));

  // batch-combination challenge alpha and its powers
  ZK_LAUNCH(fri::ch_sample_ext_kernel, 1, 1, 0, st, d_ch, d_alpha, 1u);
  if (S.max_w) ZK_LAUNCH(fri::ext_powers_kernel, (S.max_w + 127) / 128, 128, 0, st, d_alpha, d_apow, d_apow_split, S.max_w);
  CK(cudaGetLastError());
  c->launches += 2;

  // ---- opening reduction -----------------------------------------------------------------------
  uint32_t* ro[33];
  uint64_t num_reduced[33];
  for (int i = 0; i < 33; i++) {
    ro[i] = nullptr;
    num_reduced[i] = 0;
  }
  struct DenKey {
    uint32_t L;
    uint32_t z[4];
    bool operator<(const DenKey& o) const { return L != o.L ? L < o.L : memcmp(z, o.z, 16) < 0; }
  };
  struct DenBuf {
    uint32_t *inv, *wts;
  };
  std::map<DenKey, DenBuf> dens;  // per (LDE height, opening point): fri::inv_den_kernel
  {
    ProfScope ps(c, "open_reduce");
    // pass 1 (context stream): output buffers per height, and 1 / (z - x_r) over the LDE domain + the barycentric
    // weights over its low coset once per (height, point)
    {
      const uint32_t* pt = d_pts;
      uint32_t k = 0;
      for (uint32_t r = 0; r < n_rounds; r++) {
        const zk_pdata* pd = rounds[r];
        for (uint32_t m = 0; m < pd->n; m++, k++) {
          uint64_t H = pd->heights[m];
          uint32_t L = kbh::log2_exact(H), n = L - log_blowup;
          if (!ro[L]) {
            RC(sc.alloc(&ro[L], H * 16));
            CK(cudaMemsetAsync(ro[L], 0, H * 16, st));
          }
          if (pd->widths[m] != 0) {
            uint32_t gL = kbh::two_adic_generator(L);
            for (uint32_t q = 0; q < n_points[k]; q++) {
              const uint32_t* hz = points + 4 * (size_t)((pt - d_pts) / 4 + q);
              DenKey key{L, {hz[0], hz[1], hz[2], hz[3]}};
              if (dens.find(key) == dens.end()) {
                DenBuf b;
                RC(sc.alloc(&b.inv, H * 16));
                RC(sc.alloc(&b.wts, (1ull << n) * 32));
                ZK_LAUNCH(fri::inv_den_kernel, (unsigned)((H / fri::INV_BATCH + 255) / 256 + 1), 256, 0, st, pt + 4 * q, L, n, gL,
                          b.inv, b.wts);
                c->launches++;
                dens.emplace(key, b);
              }
            }
          }
          pt += 4 * n_points[k];
        }
      }
      CK(cudaGetLastError());
    }
    // pass 2: the chain of one matrix -- row reduction, barycentric partial sums and their finish, reduced openings --
    // runs on side stream (matrix index mod nslots); only the accumulation into ro[L] stays on the context's stream, in
    // matrix order.  Most of these kernels are far too small to fill the GPU (a 2^19 x 4 quotient chunk: 128 CTAs, 29 us),
    // so the chains of different matrices overlap instead of queueing behind each other.
    const bool fork = nslots > 1;
    bool slot_used[zk_ctx::NSIDE] = {false, false, false, false};
    if (fork) {
      CK(cudaEventRecord(c->side_fork, st));
      for (int i = 0; i < nslots; i++) CK(cudaStreamWaitEvent(c->side[i], c->side_fork, 0));
    }
    uint32_t* out = d_proof;
    const uint32_t* pt = d_pts;
    uint32_t k = 0, live = 0;
    for (uint32_t r = 0; r < n_rounds; r++) {
      const zk_pdata* pd = rounds[r];
      for (uint32_t m = 0; m < pd->n; m++, k++) {
        uint64_t H = pd->heights[m];
        uint32_t w = pd->widths[m], pitch = pd->pitches[m];
        uint32_t L = kbh::log2_exact(H), n = L - log_blowup;
        if (w == 0 || n_points[k] == 0) {
          pt += 4 * n_points[k];
          continue;
        }
        const int slot = fork ? (int)(live++ % (uint32_t)nslots) : 0;
        cudaStream_t ss = fork ? c->side[slot] : st;
        uint32_t *rowred = s_rowred[slot], *partial = s_partial[slot], *red = s_red[slot], *aoff = s_aoff[slot];
        // the slot's buffers are free once the context stream has consumed their previous contents
        if (fork && slot_used[slot]) CK(cudaStreamWaitEvent(ss, c->side_cons[slot], 0));
        if (w >= 64) {
          if (H % 4 == 0) {  // one warp per four rows
            auto kfn = fri::row_reduce_warp_kernel<4>;
            ZK_LAUNCH_COOP(kfn, (unsigned)((H / 4 * 32 + 255) / 256), 256, 0, ss, pd->mats[m], H, w, pitch, d_apow_split, rowred);
          } else {
            auto kfn = fri::row_reduce_warp_kernel<1>;
            ZK_LAUNCH_COOP(kfn, (unsigned)((H * 32 + 255) / 256), 256, 0, ss, pd->mats[m], H, w, pitch, d_apow_split, rowred);
          }
        } else
          ZK_LAUNCH_COOP(fri::row_reduce_kernel, (unsigned)((H + 255) / 256), 256, 0, ss, pd->mats[m], H, w, pitch, d_apow_split,
                         rowred);
        c->launches++;
        for (uint32_t p0 = 0; p0 < n_points[k]; p0 += 2) {
          uint32_t np = std::min(2u, n_points[k] - p0);
          uint64_t N = 1ull << n;
          uint32_t nchunks = (uint32_t)((N + fri::BARY_ROWS - 1) / fri::BARY_ROWS);
          const uint32_t *invp[2] = {nullptr, nullptr}, *wtsp[2] = {nullptr, nullptr};
          for (uint32_t q = 0; q < np; q++) {
            const uint32_t* hz = points + 4 * (size_t)((pt - d_pts) / 4 + q);
            DenKey key{L, {hz[0], hz[1], hz[2], hz[3]}};
            auto it = dens.find(key);
            invp[q] = it->second.inv;
            wtsp[q] = it->second.wts;
          }
          // a second pair of points of the same matrix reuses the slot's red / aoff / partial
          if (fork && p0 > 0) CK(cudaStreamWaitEvent(ss, c->side_cons[slot], 0));
          {
            // two columns per lane need 8-byte aligned rows: an even PITCH (odd widths are padded), the lane past the
            // last column accumulates the padding column and drops it
            const uint32_t cpl = ((pitch & 1u) == 0 && ((uintptr_t)pd->mats[m] % 8) == 0) ? 2 : 1;
            uint32_t log_cw = 0;
            while (log_cw < 5 && (cpl << log_cw) < w) log_cw++;
            const uint32_t ntile = (w + (cpl << log_cw) - 1) / (cpl << log_cw);
            if (cpl == 2) {
              auto kfn = fri::bary_partial_kernel<2>;
              ZK_LAUNCH_COOP(kfn, nchunks * ntile, 256, 0, ss, pd->mats[m], n, w, pitch, wtsp[0], wtsp[1], np, log_cw, partial);
            } else {
              auto kfn = fri::bary_partial_kernel<1>;
              ZK_LAUNCH_COOP(kfn, nchunks * ntile, 256, 0, ss, pd->mats[m], n, w, pitch, wtsp[0], wtsp[1], np, log_cw, partial);
            }
          }
          ZK_LAUNCH_COOP(fri::bary_final_kernel, w, 128, 0, ss, partial, nchunks, w, n, pt, np, out);
          ZK_LAUNCH_COOP(fri::reduce_ys_kernel, 1, 256, 0, ss, out, d_apow, d_alpha, w, np, num_reduced[L], red, aoff);
          if (fork) {
            CK(cudaEventRecord(c->side_prod[slot], ss));
            CK(cudaStreamWaitEvent(st, c->side_prod[slot], 0));
          }
          ZK_LAUNCH(fri::ro_accumulate_kernel, (unsigned)((H + 255) / 256), 256, 0, st, ro[L], rowred, L, invp[0], invp[1], np,
                    red, aoff);
          if (fork) {
            CK(cudaEventRecord(c->side_cons[slot], st));
            slot_used[slot] = true;
          }
          CK(cudaGetLastError());
          c->launches += 4;
          num_reduced[L] += (uint64_t)np * w;
          out += (uint64_t)np * w * 4;
          pt += 4 * np;
        }
      }
    }
    // every side chain ends in an ro_accumulate on the context's stream, which waited for it: the side streams are idle
    // once the context stream reaches this point
  }

  // ---- FRI commit phase ------------------------------------------------------------------------
  uint32_t* commits = d_proof + S.opened_words;
  uint32_t* final_slot = commits + (uint64_t)S.n_layers * 8;
  uint32_t* witness_slot = final_slot + 4;
  uint32_t* queries = witness_slot + 1;
  uint32_t* cur = ro[S.log_max];
  struct LayerRec {  // what the query phase needs of a committed layer
    const uint32_t* leaves;
    const uint32_t* digests;
    uint32_t log_h;
  };
  std::vector<LayerRec> layer_rec;
  {
    ProfScope ps(c, "fri_commit_phase");
    // layers whose input still has more than 2^FRI_TAIL_MAX_LOG elements: one MMCS commit + transcript + fold each
    uint32_t i = 0;
    for (; i < S.n_layers && S.log_max - i > (uint32_t)fri::FRI_TAIL_MAX_LOG; i++) {
      uint32_t Li = S.log_max - i;
      uint64_t hh = 1ull << (Li - 1);
      zk_pdata* lp = nullptr;
      RC(mmcs_commit_one_dev(c, cur, hh, 8, false, false, &lp, false));
      sc.pds.push_back(lp);
      layer_rec.push_back(LayerRec{lp->mats[0], lp->digests, lp->log_max});
      CK(cudaMemcpyAsync(commits + 8 * i, pdata_root_dev(lp), 32, cudaMemcpyDeviceToDevice, st));
      ZK_LAUNCH(fri::ch_observe_kernel, 1, 1, 0, st, d_ch, pdata_root_dev(lp), 8u);
      ZK_LAUNCH(fri::ch_sample_ext_kernel, 1, 1, 0, st, d_ch, d_beta, 1u);
      uint32_t* nxt;
      RC(sc.alloc(&nxt, hh * 16));
      ZK_LAUNCH(fri::fold_kernel, (unsigned)((hh + 255) / 256), 256, 0, st, cur, nxt, Li, kbh::inv(kbh::two_adic_generator(Li)),
                d_beta, ro[Li - 1]);
      CK(cudaGetLastError());
      c->launches += 3;
      cur = nxt;
    }
    if (i < S.n_layers) {
      // the remaining layers, the final-polynomial check and its observation: ONE launch (fri::fri_tail_kernel)
      std::vector<fri::TailLayer> tl;
      for (; i < S.n_layers; i++) {
        uint32_t Li = S.log_max - i;
        uint64_t hh = 1ull << (Li - 1);
        fri::TailLayer t;
        t.cur = cur;
        RC(sc.alloc(&t.nxt, hh * 16));
        RC(sc.alloc(&t.digests, (2 * hh - 1) * 32));
        t.ro_next = ro[Li - 1];
        t.commit_slot = commits + 8 * i;
        t.L = Li;
        t.gL_inv = kbh::inv(kbh::two_adic_generator(Li));
        tl.push_back(t);
        layer_rec.push_back(LayerRec{t.cur, t.digests, Li - 1});
        cur = t.nxt;
      }
      fri::TailLayer* d_tl;
      RC(sc.alloc(&d_tl, tl.size() * sizeof(fri::TailLayer)));
      CK(cudaMemcpyAsync(d_tl, tl.data(), tl.size() * sizeof(fri::TailLayer), cudaMemcpyHostToDevice, st));
      ZK_LAUNCH_COOP(fri::fri_tail_kernel, 1, 1024, 0, st, d_tl, (uint32_t)tl.size(), d_ch, 1u << (S.log_max - S.n_layers),
                     final_slot, d_status);
      CK(cudaGetLastError());
      c->launches++;
    } else {
      ZK_LAUNCH(fri::final_poly_kernel, 1, 1, 0, st, cur, 1u << (S.log_max - S.n_layers), final_slot, d_status);
      ZK_LAUNCH(fri::ch_observe_kernel, 1, 1, 0, st, d_ch, final_slot, 4u);
      CK(cudaGetLastError());
      c->launches += 2;
    }
  }

  // ---- proof of work ---------------------------------------------------------------------------
  {
    ProfScope ps(c, "grind");
    if (inject_witness < 0) {
      const uint32_t batch = 1u << 19;  // 8x the expected 2^16 tries: one batch suffices with probability 1 - e^-8
      for (uint64_t base = 0; base < kbh::P; base += batch) {
        ZK_LAUNCH(fri::grind_kernel, batch / 256, 256, 0, st, d_ch, pow_bits, (uint32_t)base, batch, d_found);
        c->launches++;
        uint32_t found;
        CK(cudaMemcpyAsync(&found, d_found, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (found != 0xffffffffu) break;
      }
    }
    ZK_LAUNCH(fri::ch_witness_kernel, 1, 1, 0, st, d_ch, pow_bits, d_found, inject_witness, witness_slot, d_status);
    c->launches++;
  }

  // ---- query phase -----------------------------------------------------------------------------
  {
    ProfScope ps(c, "fri_queries");
    if (num_queries) {
      ZK_LAUNCH(fri::ch_sample_bits_kernel, 1, 1, 0, st, d_ch, S.log_max, num_queries, d_idx);
      c->launches++;
      for (uint32_t r = 0; r < n_rounds; r++) {
        const zk_pdata* pd = rounds[r];
        uint32_t* base = queries + S.round_off[r];
        RC(pdata_open_dev(c, pd, num_queries, d_idx, S.log_max - pd->log_max, base, S.query_words, base + pd->sum_w,
                          S.query_words));
      }
      for (uint32_t i = 0; i < S.n_layers; i++) {
        const LayerRec& lp = layer_rec[i];
        ZK_LAUNCH(fri::fri_layer_query_kernel, num_queries, 64, 0, st, lp.leaves, lp.digests, lp.log_h, i,
                  d_idx, queries, S.query_words, S.layer_q_off[i]);
        c->launches++;
      }
      CK(cudaGetLastError());
    }
  }

  uint32_t status = 0;
  CK(cudaMemcpyAsync(proof_host, d_proof, S.total_words * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ch, d_ch, sizeof(fri::Chal), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (status & 4u) return zk_fail(ZK_ERR_STATE, "FRI: final folded vector is not constant (inputs are not low degree)");
  if (status & 2u) return zk_fail(ZK_ERR_STATE, "grind: no proof-of-work witness found");
  if (status & 1u) return zk_fail(ZK_ERR_VERIFY, "injected pow_witness does not satisfy the proof-of-work condition");
  return ZK_OK;
}

// ------------------------------------------------------------------------------------------------
// device challenger entry points (DuplexChallenger<Val, Perm, 16, 8>)
// ------------------------------------------------------------------------------------------------
namespace {
struct ChalOnDev {
  zk_ctx* c;
  fri::Chal* d = nullptr;
  int32_t up(const zk_challenger* ch) {
    RC(dev_alloc(c, sizeof(fri::Chal), (void**)&d));
    CK(cudaMemcpyAsync(d, ch, sizeof(fri::Chal), cudaMemcpyHostToDevice, c->stream));
    return ZK_OK;
  }
  int32_t down(zk_challenger* ch) {
    CK(cudaMemcpyAsync(ch, d, sizeof(fri::Chal), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return dev_free(c, d);
  }
};
}  // namespace

extern "C" int32_t zk_challenger_init(zk_challenger* ch) {
  if (!ch) return zk_fail(ZK_ERR_ARG, "null argument");
  memset(ch, 0, sizeof *ch);
  return ZK_OK;
}

extern "C" int32_t zk_challenger_observe(zk_ctx* c, zk_challenger* ch, const uint32_t* vals, uint32_t n) {
  if (!c || !ch || (!vals && n)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ChalOnDev D{c};
  RC(D.up(ch));
  uint32_t* dv;
  RC(dev_alloc(c, n * 4ull, (void**)&dv));
  CK(cudaMemcpyAsync(dv, vals, n * 4ull, cudaMemcpyHostToDevice, c->stream));
  ZK_LAUNCH(fri::ch_observe_kernel, 1, 1, 0, c->stream, D.d, dv, n);
  CK(cudaGetLastError());
  c->launches++;
  RC(D.down(ch));
  return dev_free(c, dv);
}

extern "C" int32_t zk_challenger_sample_ext(zk_ctx* c, zk_challenger* ch, uint32_t n_ext, uint32_t* out) {
  if (!c || !ch || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n_ext == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ChalOnDev D{c};
  RC(D.up(ch));
  uint32_t* dv;
  RC(dev_alloc(c, n_ext * 16ull, (void**)&dv));
  ZK_LAUNCH(fri::ch_sample_ext_kernel, 1, 1, 0, c->stream, D.d, dv, n_ext);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(out, dv, n_ext * 16ull, cudaMemcpyDeviceToHost, c->stream));
  RC(D.down(ch));
  return dev_free(c, dv);
}

extern "C" int32_t zk_challenger_sample_bits(zk_ctx* c, zk_challenger* ch, uint32_t bits, uint32_t n, uint64_t* out) {
  if (!c || !ch || !out) return zk_fail(ZK_ERR_ARG, "null argument");
  if (n == 0) return ZK_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ChalOnDev D{c};
  RC(D.up(ch));
  uint64_t* dv;
  RC(dev_alloc(c, n * 8ull, (void**)&dv));
  ZK_LAUNCH(fri::ch_sample_bits_kernel, 1, 1, 0, c->stream, D.d, bits, n, dv);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(out, dv, n * 8ull, cudaMemcpyDeviceToHost, c->stream));
  RC(D.down(ch));
  return dev_free(c, dv);
}

extern "C" int32_t zk_challenger_grind(zk_ctx* c, zk_challenger* ch, uint32_t bits, uint32_t* witness) {
  if (!c || !ch || !witness) return zk_fail(ZK_ERR_ARG, "null argument");
  if (bits > 30) return zk_fail(ZK_ERR_ARG, "bits must be <= 30");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ChalOnDev D{c};
  RC(D.up(ch));
  uint32_t* d_aux;  // [found, witness slot, status]
  RC(dev_alloc(c, 12, (void**)&d_aux));
  CK(cudaMemsetAsync(d_aux, 0xff, 4, c->stream));
  CK(cudaMemsetAsync(d_aux + 1, 0, 8, c->stream));
  const uint32_t batch = 1u << 22;
  uint32_t found = 0xffffffffu;
  for (uint64_t base = 0; base < kbh::P && found == 0xffffffffu; base += batch) {
    ZK_LAUNCH(fri::grind_kernel, batch / 256, 256, 0, c->stream, D.d, bits, (uint32_t)base, batch, d_aux);
    c->launches++;
    CK(cudaMemcpyAsync(&found, d_aux, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  ZK_LAUNCH(fri::ch_witness_kernel, 1, 1, 0, c->stream, D.d, bits, d_aux, (int64_t)-1, d_aux + 1, d_aux + 2);
  CK(cudaGetLastError());
  c->launches++;
  CK(cudaMemcpyAsync(witness, d_aux + 1, 4, cudaMemcpyDeviceToHost, c->stream));
  RC(D.down(ch));
  RC(dev_free(c, d_aux));
  if (found == 0xffffffffu) return zk_fail(ZK_ERR_STATE, "grind: no witness found");
  return ZK_OK;
}
