// zkgpu.hpp -- C++17 host-side mirror of the Plonky3 / Ziren interfaces this library stands behind,
// written over the C ABI of zkgpu.h (header only; no CUDA or torch types).
//
// The reference host is Rust; this image has no Rust toolchain, so the mirror is C++ with the reference's
// names, argument meaning and error behaviour, to be read next to:
//   TwoAdicMultiplicativeCoset / natural_domain_for_degree .... crates/stark/src/prover.rs:271,319,509
//   Pcs::commit ............................................... crates/stark/src/prover.rs:277,403,497
//   Pcs::get_evaluations_on_domain ............................ crates/stark/src/prover.rs:437-445
//   Pcs::open -> (OpenedValues, FriProof) ..................... crates/stark/src/prover.rs:546-556
//   FriProof / QueryProof / BatchOpening / CommitPhaseProofStep crates/recursion/circuit/src/witness/stark.rs:60-142
//   DuplexChallenger<Val, Perm, 16, 8> ........................ crates/stark/src/kb31_poseidon2.rs:180
//   quotient_values ........................................... crates/stark/src/quotient.rs:19-31
// Errors: every failing call throws zkgpu::Error carrying zk_last_error() (the Rust caller `unwrap`s,
// crates/core/machine/src/utils/prove.rs:497); nothing here computes field arithmetic on the CPU except
// the O(1) scalar helpers of Domain.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkgpu.h"

namespace zkgpu {

using Val = uint32_t;                  // KoalaBear, Montgomery form
using Challenge = std::array<Val, 4>;  // BinomialExtensionField<KoalaBear, 4>
using Digest = std::array<Val, 8>;     // Hash<Val, Val, 8>

struct Error : std::runtime_error {
  int32_t status;
  Error(int32_t st, const std::string& what) : std::runtime_error(what), status(st) {}
};
inline void check(int32_t rc) {
  if (rc != ZK_OK) throw Error(rc, zk_last_error());
}

// ---- scalar field helpers (host; launch parameters only) ---------------------------------------------
namespace field {
constexpr uint32_t P = 0x7f000001u, MU = 0x81000001u, ONE = 0x01fffffeu, RR = 0x17f7efe4u, GENERATOR = 0x05fffffau;
inline uint32_t reduce(uint64_t t) {
  uint32_t m = (uint32_t)t * MU, u = (uint32_t)(((uint64_t)m * P) >> 32), hi = (uint32_t)(t >> 32);
  return hi < u ? hi - u + P : hi - u;
}
inline uint32_t mul(uint32_t a, uint32_t b) { return reduce((uint64_t)a * b); }
inline uint32_t pow(uint32_t a, uint64_t e) {
  uint32_t r = ONE;
  for (; e; e >>= 1, a = mul(a, a))
    if (e & 1) r = mul(r, a);
  return r;
}
inline uint32_t from_canonical(uint32_t c) { return mul(c % P, RR); }
inline uint32_t two_adic_generator(unsigned bits) {
  uint32_t g = pow(GENERATOR, 127);
  for (unsigned i = bits; i < 24; i++) g = mul(g, g);
  return g;
}
}  // namespace field

// TwoAdicMultiplicativeCoset { log_n, shift }
struct Domain {
  uint32_t log_n = 0;
  Val shift = field::ONE;
  uint64_t size() const { return 1ull << log_n; }
  Val first_point() const { return shift; }
  // next_point(x) = x * g  (crates/recursion/circuit/src/domain.rs:32-44)
  Challenge next_point(const Challenge& x) const {
    Val g = field::two_adic_generator(log_n);
    return {field::mul(x[0], g), field::mul(x[1], g), field::mul(x[2], g), field::mul(x[3], g)};
  }
  // create_disjoint_domain(min_size): GENERATOR-shifted coset of at least that size
  Domain create_disjoint_domain(uint64_t min_size) const {
    uint32_t l = 0;
    while ((1ull << l) < min_size) l++;
    return Domain{l, field::mul(shift, field::GENERATOR)};
  }
  // split_domains(k): k cosets of size / k, shifts shift * g_{log_n}^i
  std::vector<Domain> split_domains(uint32_t num_chunks) const {
    uint32_t lc = 0;
    while ((1u << lc) < num_chunks) lc++;
    std::vector<Domain> out;
    Val g = field::two_adic_generator(log_n);
    for (uint32_t i = 0; i < num_chunks; i++) out.push_back(Domain{log_n - lc, field::mul(shift, field::pow(g, i))});
    return out;
  }
};

struct RowMajorMatrixView {
  const Val* values;
  uint64_t height;
  uint32_t width;
};

class Context {
 public:
  explicit Context(int device = 0) { check(zk_ctx_create(device, &ctx_)); }
  Context(int device, void* cuda_stream) { check(zk_ctx_create_on_stream(device, cuda_stream, &ctx_)); }
  ~Context() { zk_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  zk_ctx* raw() const { return ctx_; }
  void keep_traces(bool on) { check(zk_ctx_keep_traces(ctx_, on ? 1 : 0)); }

 private:
  zk_ctx* ctx_ = nullptr;
};

// Mmcs::ProverData: committed LDE matrices + digest layers, resident on the device
class ProverData {
 public:
  ProverData() = default;
  explicit ProverData(zk_pdata* p) : p_(p, zk_pdata_free) {}
  zk_pdata* raw() const { return p_.get(); }
  uint32_t num_matrices() const { return zk_pdata_num_matrices(raw()); }
  uint64_t height(uint32_t i) const { return zk_pdata_height(raw(), i); }
  uint32_t width(uint32_t i) const { return zk_pdata_width(raw(), i); }
  uint32_t log_max_height() const { return zk_pdata_log_max_height(raw()); }
  Digest root() const {
    Digest d;
    check(zk_pdata_root(raw(), d.data()));
    return d;
  }
  // Mmcs::get_matrices()[i].to_row_major_matrix() (host copy; only for callers that need the LDE on the CPU)
  std::vector<Val> get_matrix(uint32_t i) const {
    std::vector<Val> m(height(i) * width(i));
    check(zk_pdata_copy_lde(raw(), i, m.data()));
    return m;
  }
  // Pcs::get_evaluations_on_domain: device view of the LDE, no copy
  zk_dptr evaluations_on_domain(uint32_t i) const { return zk_pdata_lde(raw(), i); }
  // row stride of that view in words (odd widths are padded to even)
  uint32_t pitch(uint32_t i) const { return zk_pdata_pitch(raw(), i); }

 private:
  std::shared_ptr<zk_pdata> p_;
};

// TwoAdicSubgroupDft (the trait `Radix2DitParallel` implements; type alias crates/stark/src/kb31_poseidon2.rs:179)
class GpuDft {
 public:
  explicit GpuDft(Context& c) : c_(c) {}
  // dft_batch: natural-order DFT of every column
  std::vector<Val> dft_batch(const RowMajorMatrixView& m) const {
    std::vector<Val> out(m.height * m.width);
    check(zk_dft_batch(c_.raw(), m.values, m.height, m.width, out.data()));
    return out;
  }
  // coset_lde_batch(mat, added_bits, shift).bit_reverse_rows(): (height << added_bits) rows
  std::vector<Val> coset_lde_batch(const RowMajorMatrixView& m, uint32_t added_bits, Val shift) const {
    std::vector<Val> out((m.height << added_bits) * m.width);
    check(zk_coset_lde(c_.raw(), m.values, m.height, m.width, added_bits, shift, out.data()));
    return out;
  }

 private:
  Context& c_;
};

// MerkleTreeMmcs<_, _, MyHash, MyCompress, 8> (crates/stark/src/kb31_poseidon2.rs:176-177): commit / open_batch.
// (verify_batch stays on the CPU side of the caller, as in the reference's verifier.)
class MerkleTreeMmcs {
 public:
  explicit MerkleTreeMmcs(Context& c) : c_(c) {}
  std::pair<Digest, ProverData> commit(const std::vector<RowMajorMatrixView>& mats) const {
    std::vector<const Val*> ptrs;
    std::vector<uint64_t> hs;
    std::vector<uint32_t> ws;
    for (auto& m : mats) {
      ptrs.push_back(m.values);
      hs.push_back(m.height);
      ws.push_back(m.width);
    }
    Digest root;
    zk_pdata* pd = nullptr;
    check(zk_mmcs_commit(c_.raw(), (uint32_t)mats.size(), ptrs.data(), hs.data(), ws.data(), root.data(), &pd));
    return {root, ProverData(pd)};
  }
  // open_batch(index, data) -> (opened rows per matrix, sibling digests bottom-up)
  std::pair<std::vector<std::vector<Val>>, std::vector<Digest>> open_batch(uint64_t index, const ProverData& pd) const {
    uint32_t sum_w = 0;
    for (uint32_t i = 0; i < pd.num_matrices(); i++) sum_w += pd.width(i);
    std::vector<Val> flat(sum_w ? sum_w : 1), path((size_t)pd.log_max_height() * 8 + 8);
    check(zk_pdata_open_batch(pd.raw(), 1, &index, flat.data(), path.data()));
    std::vector<std::vector<Val>> rows;
    const Val* p = flat.data();
    for (uint32_t i = 0; i < pd.num_matrices(); i++) {
      rows.emplace_back(p, p + pd.width(i));
      p += pd.width(i);
    }
    std::vector<Digest> proof(pd.log_max_height());
    for (uint32_t l = 0; l < pd.log_max_height(); l++)
      for (int k = 0; k < 8; k++) proof[l][k] = path[8 * l + k];
    return {rows, proof};
  }

 private:
  Context& c_;
};

// DuplexChallenger whose permutations run on the device
class DuplexChallenger {
 public:
  explicit DuplexChallenger(Context& c) : c_(c) { zk_challenger_init(&st_); }
  void observe(Val v) { check(zk_challenger_observe(c_.raw(), &st_, &v, 1)); }
  void observe_slice(const Val* v, uint32_t n) { check(zk_challenger_observe(c_.raw(), &st_, v, n)); }
  void observe(const Digest& d) { observe_slice(d.data(), 8); }
  void observe_ext_element(const Challenge& e) { observe_slice(e.data(), 4); }
  Challenge sample_ext_element() {
    Challenge e;
    check(zk_challenger_sample_ext(c_.raw(), &st_, 1, e.data()));
    return e;
  }
  uint64_t sample_bits(uint32_t bits) {
    uint64_t v;
    check(zk_challenger_sample_bits(c_.raw(), &st_, bits, 1, &v));
    return v;
  }
  Val grind(uint32_t bits) {
    Val w;
    check(zk_challenger_grind(c_.raw(), &st_, bits, &w));
    return w;
  }
  zk_challenger& state() { return st_; }

 private:
  Context& c_;
  zk_challenger st_;
};

// ---- proof structures (field order as in Plonky3 / crates/recursion/circuit/src/witness/stark.rs) -----
struct BatchOpening {
  std::vector<std::vector<Val>> opened_values;  // one row per matrix of the round
  std::vector<Digest> opening_proof;
};
struct CommitPhaseProofStep {
  Challenge sibling_value;
  std::vector<Digest> opening_proof;
};
struct QueryProof {
  std::vector<BatchOpening> input_proof;  // one per round
  std::vector<CommitPhaseProofStep> commit_phase_openings;
};
struct FriProof {
  std::vector<Digest> commit_phase_commits;
  std::vector<QueryProof> query_proofs;
  Challenge final_poly;
  Val pow_witness;
};
// OpenedValues[round][matrix][point][column]
using OpenedValues = std::vector<std::vector<std::vector<std::vector<Challenge>>>>;

struct FriConfig {
  uint32_t log_blowup = 1, num_queries = 84, proof_of_work_bits = 16;  // crates/stark/src/kb31_poseidon2.rs:203-213
};

// TwoAdicFriPcs<Val, Dft, ValMmcs, ChallengeMmcs>
class TwoAdicFriPcs {
 public:
  TwoAdicFriPcs(Context& c, FriConfig cfg = {}) : c_(c), cfg_(cfg) {}
  const FriConfig& fri_config() const { return cfg_; }
  static Domain natural_domain_for_degree(uint64_t degree) {
    uint32_t l = 0;
    while ((1ull << l) < degree) l++;
    return Domain{l, field::ONE};
  }
  // Pcs::commit: (Com, ProverData)
  std::pair<Digest, ProverData> commit(const std::vector<std::pair<Domain, RowMajorMatrixView>>& evaluations) {
    std::vector<const Val*> ptrs;
    std::vector<uint64_t> hs;
    std::vector<uint32_t> ws, shifts;
    for (auto& [d, m] : evaluations) {
      if (m.height != d.size()) throw Error(ZK_ERR_ARG, "matrix height does not match its domain");
      ptrs.push_back(m.values);
      hs.push_back(m.height);
      ws.push_back(m.width);
      shifts.push_back(d.shift);
    }
    Digest root;
    zk_pdata* pd = nullptr;
    check(zk_commit(c_.raw(), (uint32_t)ptrs.size(), ptrs.data(), hs.data(), ws.data(), shifts.data(), cfg_.log_blowup,
                    root.data(), &pd));
    return {root, ProverData(pd)};
  }
  // same with device-resident matrices (e.g. quotient chunks, permutation traces produced on the device)
  std::pair<Digest, ProverData> commit_device(const std::vector<Domain>& domains, const std::vector<zk_dptr>& mats,
                                              const std::vector<uint32_t>& widths) {
    std::vector<uint64_t> hs;
    std::vector<uint32_t> shifts;
    for (auto& d : domains) {
      hs.push_back(d.size());
      shifts.push_back(d.shift);
    }
    Digest root;
    zk_pdata* pd = nullptr;
    check(zk_commit_dev(c_.raw(), (uint32_t)mats.size(), mats.data(), hs.data(), widths.data(), shifts.data(),
                        cfg_.log_blowup, root.data(), &pd));
    return {root, ProverData(pd)};
  }
  // Pcs::open(rounds = [(prover data, points per matrix)], challenger) -> (OpenedValues, FriProof)
  std::pair<OpenedValues, FriProof> open(const std::vector<std::pair<const ProverData*, std::vector<std::vector<Challenge>>>>& rounds,
                                         DuplexChallenger& challenger, int64_t inject_pow_witness = -1) {
    std::vector<const zk_pdata*> pds;
    std::vector<uint32_t> n_points;
    std::vector<Val> points;
    for (auto& [pd, pts] : rounds) {
      if (pts.size() != pd->num_matrices()) throw Error(ZK_ERR_ARG, "one point list per matrix is required");
      pds.push_back(pd->raw());
      for (auto& mp : pts) {
        n_points.push_back((uint32_t)mp.size());
        for (auto& z : mp) points.insert(points.end(), z.begin(), z.end());
      }
    }
    uint64_t words = zk_pcs_proof_words((uint32_t)pds.size(), pds.data(), n_points.data(), cfg_.log_blowup, cfg_.num_queries);
    std::vector<Val> flat(words);
    check(zk_pcs_open(c_.raw(), (uint32_t)pds.size(), pds.data(), n_points.data(), points.data(), cfg_.log_blowup,
                      cfg_.num_queries, cfg_.proof_of_work_bits, &challenger.state(), inject_pow_witness, flat.data(), words));
    return unflatten(rounds, flat);
  }

 private:
  // slices the flat proof (layout documented in zkgpu.h) into Plonky3's structures
  std::pair<OpenedValues, FriProof> unflatten(
      const std::vector<std::pair<const ProverData*, std::vector<std::vector<Challenge>>>>& rounds, const std::vector<Val>& f) {
    const Val* p = f.data();
    auto ext = [&]() {
      Challenge e{p[0], p[1], p[2], p[3]};
      p += 4;
      return e;
    };
    auto digest = [&]() {
      Digest d;
      for (int i = 0; i < 8; i++) d[i] = p[i];
      p += 8;
      return d;
    };
    OpenedValues ov;
    uint32_t log_max = 0;
    for (auto& [pd, pts] : rounds) {
      log_max = std::max(log_max, pd->log_max_height());
      ov.emplace_back();
      for (uint32_t m = 0; m < pd->num_matrices(); m++) {
        ov.back().emplace_back();
        for (size_t k = 0; k < pts[m].size(); k++) {
          ov.back().back().emplace_back();
          for (uint32_t c = 0; c < pd->width(m); c++) ov.back().back().back().push_back(ext());
        }
      }
    }
    FriProof fp;
    uint32_t n_layers = log_max - cfg_.log_blowup;
    for (uint32_t i = 0; i < n_layers; i++) fp.commit_phase_commits.push_back(digest());
    fp.final_poly = ext();
    fp.pow_witness = *p++;
    for (uint32_t q = 0; q < cfg_.num_queries; q++) {
      QueryProof qp;
      for (auto& [pd, pts] : rounds) {
        BatchOpening bo;
        for (uint32_t m = 0; m < pd->num_matrices(); m++) {
          bo.opened_values.emplace_back(p, p + pd->width(m));
          p += pd->width(m);
        }
        for (uint32_t l = 0; l < pd->log_max_height(); l++) bo.opening_proof.push_back(digest());
        qp.input_proof.push_back(std::move(bo));
      }
      for (uint32_t i = 0; i < n_layers; i++) {
        CommitPhaseProofStep st;
        st.sibling_value = ext();
        for (uint32_t l = 0; l + i + 1 < log_max; l++) st.opening_proof.push_back(digest());
        qp.commit_phase_openings.push_back(std::move(st));
      }
      fp.query_proofs.push_back(std::move(qp));
    }
    if ((size_t)(p - f.data()) != f.size()) throw Error(ZK_ERR_STATE, "flat proof has an unexpected length");
    return {std::move(ov), std::move(fp)};
  }
  Context& c_;
  FriConfig cfg_;
};

// quotient_values for one chip (crates/stark/src/quotient.rs:19-31); returns the device buffer holding the
// 2^log_quotient_degree chunk matrices (N x 4 each, back to back); free it with zk_dev_free after committing.
struct QuotientInputs {
  const ProverData* preprocessed = nullptr;
  uint32_t preprocessed_idx = 0;
  const ProverData* main = nullptr;
  uint32_t main_idx = 0;
  const ProverData* permutation = nullptr;
  uint32_t permutation_idx = 0;
  std::vector<Challenge> perm_challenges;
  std::vector<Val> public_values;
  Challenge local_cumulative_sum{};
  std::array<Val, 14> global_cumulative_sum{};
};
inline zk_dptr quotient_values(Context& c, const std::string& chip_name, const Domain& trace_domain, uint32_t log_quotient_degree,
                               const Challenge& alpha, const QuotientInputs& in) {
  int32_t id = zk_air_find(chip_name.c_str());
  if (id < 0) throw Error(ZK_ERR_ARG, "chip '" + chip_name + "' is not compiled into libzkgpu");
  std::vector<Val> chal;
  for (auto& e : in.perm_challenges) chal.insert(chal.end(), e.begin(), e.end());
  zk_dptr out = 0;
  check(zk_quotient(c.raw(), id, in.preprocessed ? in.preprocessed->raw() : nullptr, in.preprocessed_idx,
                    in.main ? in.main->raw() : nullptr, in.main_idx, in.permutation ? in.permutation->raw() : nullptr,
                    in.permutation_idx, trace_domain.log_n, log_quotient_degree, alpha.data(),
                    chal.empty() ? nullptr : chal.data(), in.public_values.empty() ? nullptr : in.public_values.data(),
                    (uint32_t)in.public_values.size(), in.local_cumulative_sum.data(), in.global_cumulative_sum.data(), &out));
  return out;
}

}  // namespace zkgpu
