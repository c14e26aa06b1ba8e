/* oracle/zk_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of Ziren's STARK commit / quotient / FRI hot path.  The arithmetic
 * of that path lives in the un-vendored dependency ProjectZKM/Plonky3 @ faa24ca4597eebeecbf71b194b71c7d1a99b3f01
 * (Cargo.lock:3911-4218), which is NOT under /root/reference and cannot be compiled here (no Rust).  The
 * restatement therefore follows the reference's own in-repo statements of the same algorithms (each
 * function cites them) and is pinned by:
 *   (i)  the Poseidon2 known-answer test of examples/poseidon2/host/src/main.rs:33-37,
 *   (ii) oracle/_ref: the reference's own C++ field class and Poseidon2 headers compiled in place
 *        (oracle/Makefile) and compared word for word with this file,
 *   (iii) the transliterated in-repo verifier (ork_pcs_verify, from recursion/circuit/src/fri.rs:71-405)
 *        accepting the proofs this oracle and the CUDA library produce.
 * PARITY STATUS: Poseidon2 / sponge / compression / field arithmetic are pinned by (i)+(ii).  Merkle
 * roots, LDE layout, quotient commitments and FRI transcripts have NO stored golden vector in the
 * reference ("parity unpinned" against Plonky3 bit-for-bit); they are pinned structurally by (iii).
 *
 * All field elements cross this interface as uint32 Montgomery residues (R = 2^32), the in-memory
 * form of a Rust `KoalaBear` (crates/core/machine/cpp/extern.cpp:12).
 */
#ifndef ZK_ORACLE_H
#define ZK_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ork_tree ork_tree;

/* ---- field helpers exported for the Python tests ---- */
uint32_t ork_to_monty(uint32_t canonical);
uint32_t ork_from_monty(uint32_t m);
uint32_t ork_mul(uint32_t a, uint32_t b);
uint32_t ork_inv(uint32_t a);
uint32_t ork_two_adic_generator(uint32_t bits);
void ork_ext_mul(const uint32_t a[4], const uint32_t b[4], uint32_t out[4]);
void ork_ext_inv(const uint32_t a[4], uint32_t out[4]);
void ork_to_monty_vec(const uint32_t* canonical, uint32_t* out, uint64_t n);
void ork_from_monty_vec(const uint32_t* m, uint32_t* out, uint64_t n);

/* ---- Poseidon2 ---- */
void ork_poseidon2_permute(uint32_t state[16]);                          /* Montgomery in/out */
void ork_poseidon2_permute_canonical(uint32_t state[16]);                /* canonical in/out (the guest syscall form) */
/* trace fillers: Poseidon2WideChip<3|9> main / preprocessed trace, AddSubChip main trace (see zk_oracle.c) */
void ork_poseidon2_wide_trace(const uint32_t* inputs, uint64_t n_events, uint64_t rows, int32_t sbox, uint32_t* out);
void ork_poseidon2_wide_prep(const uint32_t* instrs, uint64_t n, uint64_t rows, uint32_t* out);
void ork_add_sub_trace(const uint32_t* events, uint64_t n_events, uint64_t rows, uint32_t* out);
void ork_hash(const uint32_t* in, uint64_t n, uint32_t out[8]);          /* PaddingFreeSponge<16,8,8> */
void ork_compress(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]); /* TruncatedPermutation<2,8,16> */
void ork_hash_rows(const uint32_t* mat, uint64_t h, uint64_t w, uint32_t* digests /* h*8 */);

/* ---- DFT / LDE ---- */
/* Natural-order forward DFT of every column: out[k][c] = sum_j in[j][c] g_n^{jk}. */
void ork_dft_batch(const uint32_t* in, uint64_t h, uint64_t w, uint32_t* out);
/* coset_lde_batch(...).bit_reverse_rows(): out has (h << log_blowup) rows; row r holds
 * q(shift * g_{n+b}^{bitrev(r)}) where q interpolates `in` over the subgroup of order h. */
void ork_coset_lde(const uint32_t* in, uint64_t h, uint64_t w, uint32_t log_blowup, uint32_t shift,
                   uint32_t* out);

/* ---- MMCS ---- */
/* Mixed-height Merkle tree over row-major matrices (heights powers of two).  The tree COPIES the
 * matrices when copy != 0, otherwise it borrows the caller's pointers (they must outlive the tree). */
int32_t ork_mmcs_commit(uint32_t n_mats, const uint32_t* const* mats, const uint64_t* heights,
                        const uint64_t* widths, int32_t copy, uint32_t root[8], ork_tree** out);
void ork_tree_free(ork_tree* t);
uint32_t ork_tree_num_matrices(const ork_tree* t);
uint64_t ork_tree_height(const ork_tree* t, uint32_t i);
uint64_t ork_tree_width(const ork_tree* t, uint32_t i);
const uint32_t* ork_tree_matrix(const ork_tree* t, uint32_t i);
uint32_t ork_tree_log_max_height(const ork_tree* t);
const uint32_t* ork_tree_layer(const ork_tree* t, uint32_t layer); /* layer 0 = leaves */
/* open_batch: opened rows are written back to back in matrix order (sum of widths words),
 * proof = log_max_height siblings of 8 words, bottom-up. */
void ork_tree_open(const ork_tree* t, uint64_t index, uint32_t* opened, uint32_t* proof);
/* verify_batch transliterated from recursion/circuit/src/fri.rs:363-405.  Returns 1 when accepted. */
int32_t ork_mmcs_verify(const uint32_t root[8], uint32_t n_mats, const uint64_t* heights,
                        const uint64_t* widths, uint64_t index, const uint32_t* opened,
                        const uint32_t* proof, uint32_t proof_len);

/* ---- Pcs::commit: LDE every matrix with shift = 3 / domain_shift, then MMCS ---- */
int32_t ork_pcs_commit(uint32_t n_mats, const uint32_t* const* mats, const uint64_t* heights,
                       const uint64_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup,
                       uint32_t root[8], ork_tree** out);

int32_t ork_num_threads(void);
void ork_set_num_threads(int32_t n);

/* ---- DuplexChallenger<Val, Perm, 16, 8> ---- */
typedef struct {
  uint32_t state[16];
  uint32_t in[8];
  uint32_t n_in;
  uint32_t out[8];
  uint32_t n_out;
} ork_challenger;
void ork_ch_init(ork_challenger* c);
void ork_ch_observe(ork_challenger* c, const uint32_t* v, uint32_t n);
uint32_t ork_ch_sample(ork_challenger* c);
void ork_ch_sample_ext(ork_challenger* c, uint32_t out[4]);
uint32_t ork_ch_sample_bits(ork_challenger* c, uint32_t bits);
int32_t ork_ch_check_witness(ork_challenger* c, uint32_t bits, uint32_t witness);
uint32_t ork_ch_grind(ork_challenger* c, uint32_t bits); /* smallest canonical witness, Montgomery form */

/* ---- TwoAdicFriPcs::open / verify ----
 * Matrices are described round by round, in commit order; n_points[k] / points (4 words each) list the
 * opening points of matrix k.  The proof is one flat u32 buffer:
 *   opened values : for round, matrix, point: width ext elements
 *   fri           : n_layers * 8 (commit-phase roots), final_poly (4), pow_witness (1)
 *   queries       : num_queries x { per round: opened rows (sum of widths) + path (log_max_r * 8);
 *                                   per layer i: sibling (4) + path ((log_max - i - 1) * 8) }
 * inject_witness >= 0 uses that pow witness instead of grinding (parity with a given transcript). */
uint64_t ork_pcs_proof_words(uint32_t n_rounds, const uint32_t* n_mats, const uint64_t* lde_heights,
                             const uint32_t* widths, const uint32_t* n_points, uint32_t log_blowup,
                             uint32_t num_queries);
int32_t ork_pcs_open(uint32_t n_rounds, const ork_tree* const* trees, const uint32_t* n_points,
                     const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                     ork_challenger* ch, int64_t inject_witness, uint32_t* proof, uint64_t proof_cap);
/* returns 1 when the proof is accepted, a negative code naming the failed check otherwise */
int32_t ork_pcs_verify(uint32_t n_rounds, const uint32_t* roots, const uint32_t* n_mats,
                       const uint64_t* lde_heights, const uint32_t* widths, const uint32_t* n_points,
                       const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                       ork_challenger* ch, const uint32_t* proof, uint64_t proof_words);

#ifdef __cplusplus
}
#endif
#endif
