// tests/cpp/host_mirror.cpp -- drives include/zkgpu.hpp the way Ziren's prover drives Plonky3
// (crates/stark/src/prover.rs:258-292, 298-653) for a one-chip Fibonacci shard, and prints the transcript
// artefacts as JSON so that the Python test can compare them with the oracle.  Linked against libzkgpu.so on
// a GPU box, or against the test-only emulator build in the GPU-less container.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "zkgpu.hpp"

using namespace zkgpu;

static void print_words(const char* key, const Val* v, size_t n, bool comma = true) {
  printf("\"%s\": [", key);
  for (size_t i = 0; i < n; i++) printf("%s%u", i ? ", " : "", v[i]);
  printf("]%s\n", comma ? "," : "");
}

int main(int argc, char** argv) {
  const uint32_t log_n = argc > 1 ? (uint32_t)atoi(argv[1]) : 6;
  try {
    Context ctx(0);
    TwoAdicFriPcs pcs(ctx, FriConfig{1, (uint32_t)(argc > 2 ? atoi(argv[2]) : 8), (uint32_t)(argc > 3 ? atoi(argv[3]) : 6)});
    // generate_trace_rows (crates/stark/src/stark_testing.rs:63-81), values converted to Montgomery form
    const uint64_t n = 1ull << log_n;
    std::vector<Val> trace(2 * n);
    uint64_t a = 1, b = 1;
    for (uint64_t i = 0; i < n; i++) {
      trace[2 * i] = field::from_canonical((uint32_t)a);
      trace[2 * i + 1] = field::from_canonical((uint32_t)b);
      uint64_t c = (a + b) % field::P;
      a = b;
      b = c;
    }
    std::vector<Val> pis = {field::from_canonical(1), field::from_canonical(1), trace[2 * (n - 1) + 1]};

    DuplexChallenger ch(ctx);
    Domain trace_domain = TwoAdicFriPcs::natural_domain_for_degree(n);
    auto [main_commit, main_data] = pcs.commit({{trace_domain, RowMajorMatrixView{trace.data(), n, 2}}});
    ch.observe_slice(pis.data(), 3);
    ch.observe(main_commit);
    Challenge p0 = ch.sample_ext_element(), p1 = ch.sample_ext_element();  // permutation challenges (unused by this chip)
    (void)p0;
    (void)p1;
    Challenge alpha = ch.sample_ext_element();

    const uint32_t lqd = 1;
    QuotientInputs qi;
    qi.main = &main_data;
    qi.public_values = pis;
    zk_dptr chunks = quotient_values(ctx, "fibonacci", trace_domain, lqd, alpha, qi);
    Domain quotient_domain = trace_domain.create_disjoint_domain(n << lqd);
    std::vector<Domain> qc_domains = quotient_domain.split_domains(1u << lqd);
    std::vector<zk_dptr> chunk_ptrs;
    for (uint32_t c = 0; c < (1u << lqd); c++) chunk_ptrs.push_back(chunks + (uint64_t)c * n * 16);
    auto [quotient_commit, quotient_data] = pcs.commit_device(qc_domains, chunk_ptrs, {4, 4});
    check(zk_dev_free(ctx.raw(), chunks));
    ch.observe(quotient_commit);
    Challenge zeta = ch.sample_ext_element();

    std::vector<std::pair<const ProverData*, std::vector<std::vector<Challenge>>>> rounds = {
        {&main_data, {{zeta, trace_domain.next_point(zeta)}}},
        {&quotient_data, {{zeta}, {zeta}}},
    };
    auto [opened, proof] = pcs.open(rounds, ch);

    // trait-level pieces: TwoAdicSubgroupDft::coset_lde_batch and Mmcs::{commit, open_batch}
    GpuDft dft(ctx);
    std::vector<Val> lde = dft.coset_lde_batch(RowMajorMatrixView{trace.data(), n, 2}, 1, field::GENERATOR);
    MerkleTreeMmcs mmcs(ctx);
    auto [mmcs_root, mmcs_data] = mmcs.commit({RowMajorMatrixView{lde.data(), 2 * n, 2}});
    auto [rows, path] = mmcs.open_batch(5, mmcs_data);

    printf("{\n");
    print_words("mmcs_root_of_lde", mmcs_root.data(), 8);
    print_words("opened_row_5", rows[0].data(), rows[0].size());
    printf("\"path_len\": %zu,\n", path.size());
    print_words("main_commit", main_commit.data(), 8);
    print_words("quotient_commit", quotient_commit.data(), 8);
    print_words("alpha", alpha.data(), 4);
    print_words("zeta", zeta.data(), 4);
    print_words("final_poly", proof.final_poly.data(), 4);
    printf("\"pow_witness\": %u,\n", proof.pow_witness);
    printf("\"n_layers\": %zu, \"n_queries\": %zu,\n", proof.commit_phase_commits.size(), proof.query_proofs.size());
    print_words("main_at_zeta", opened[0][0][0][0].data(), 4);
    print_words("first_sibling", proof.query_proofs[0].commit_phase_openings[0].sibling_value.data(), 4);
    print_words("challenger_state", ch.state().state, 16, false);
    printf("}\n");
  } catch (const Error& e) {
    fprintf(stderr, "zkgpu error %d: %s\n", e.status, e.what());
    return 1;
  }
  return 0;
}
