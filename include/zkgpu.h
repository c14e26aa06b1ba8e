/* zkgpu.h -- C ABI of libzkgpu.so, the B200 (sm_100a) implementation of Ziren's STARK trace-commitment,
 * quotient and FRI hot path.
 *
 * This is the boundary a thin Rust `-sys` crate binds, following the reference's own FFI idiom
 * (`extern "C-unwind"` over `u32` Montgomery words: crates/core/machine/src/sys.rs:14-42,
 * crates/core/machine/cpp/extern.cpp:12).  INTEGRATION.md shows the Rust side.
 *
 * Conventions
 *   - every field element is a uint32 canonical Montgomery residue of KoalaBear (R = 2^32), i.e. the
 *     in-memory form of a Rust `KoalaBear` (crates/core/machine/include/kb31_t.hpp:27-34);
 *     an extension element (F_p[X]/(X^4-3)) is 4 such words, coefficient 0 first
 *     (crates/stark/src/air/extension.rs:14-25);
 *   - matrices are row-major (`RowMajorMatrix<KoalaBear>`), heights are powers of two;
 *   - every function returns 0 on success and a negative zk_status otherwise; nothing throws or
 *     unwinds across the boundary; zk_last_error() gives the text of the calling thread's last error;
 *   - `*_host` pointers are caller-owned host memory, `zk_dptr` values are device addresses owned by the
 *     library (or handed in by the caller for the `_dev` variants);
 *   - a zk_ctx is bound to one GPU and one CUDA stream; calls on the same ctx are serialised by an
 *     internal mutex, so rayon workers may share it (crates/core/machine/src/utils/prove.rs:487-497).
 *   - zk_pdata handles point into their context: zk_ctx_destroy with handles still alive is DEFERRED until the
 *     last of them is freed (bindings drop handles from destructors in any order).
 *   - there is NO CPU fallback: without a CUDA device zk_ctx_create fails with ZK_ERR_CUDA.
 */
#ifndef ZKGPU_H
#define ZKGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ZK_OK = 0,
  ZK_ERR_ARG = -1,    /* bad argument (null pointer, height not a power of two, ...) */
  ZK_ERR_CUDA = -2,   /* CUDA runtime error; see zk_last_error() */
  ZK_ERR_STATE = -3,  /* object used in the wrong state */
  ZK_ERR_VERIFY = -4  /* a verification entry point rejected its input */
} zk_status;

typedef struct zk_ctx zk_ctx;
typedef struct zk_pdata zk_pdata; /* = Mmcs::ProverData: LDE matrices + digest layers, device resident */
typedef uint64_t zk_dptr;          /* device address */

/* ---- context ---------------------------------------------------------------------------------- */
/* One context per GPU (shard dispatch places one shard per GPU: crates/core/machine/src/utils/prove.rs:480-526). */
int32_t zk_ctx_create(int32_t device, zk_ctx** out);
/* Same, but all work is enqueued on the caller's cudaStream_t (e.g. the host framework's current stream). */
int32_t zk_ctx_create_on_stream(int32_t device, void* cuda_stream, zk_ctx** out);
void zk_ctx_destroy(zk_ctx* ctx);
int32_t zk_ctx_sync(zk_ctx* ctx);
const char* zk_last_error(void);
/* "sm_100a;<git-less build tag>" -- lets a binding check it loaded the right library. */
const char* zk_build_info(void);

/* Per-stage device timings (CUDA events on the ctx stream).  enable=1 starts recording. */
int32_t zk_prof_enable(zk_ctx* ctx, int32_t enable);
int32_t zk_prof_reset(zk_ctx* ctx);
int32_t zk_prof_count(zk_ctx* ctx);
int32_t zk_prof_get(zk_ctx* ctx, int32_t i, char* name, int32_t name_cap, float* ms, uint64_t* launches);
/* start of record i on the ctx stream, in ms after the start of record 0 (timeline of the stages) */
int32_t zk_prof_start(zk_ctx* ctx, int32_t i, float* ms_after_first);
/* number of kernels of this library launched on the ctx since creation */
uint64_t zk_launch_count(zk_ctx* ctx);
/* NTT passes this process has run on the persistent TMA-fed kernel (csrc/ntt_tma.cuh) -- tests assert the path was taken */
uint64_t zk_ntt_tma_passes(void);

/* ---- device memory plumbing --------------------------------------------------------------------- */
int32_t zk_dev_alloc(zk_ctx* ctx, uint64_t bytes, zk_dptr* out);
int32_t zk_dev_free(zk_ctx* ctx, zk_dptr p);
int32_t zk_h2d(zk_ctx* ctx, zk_dptr dst, const void* src_host, uint64_t bytes);
int32_t zk_d2h(zk_ctx* ctx, void* dst_host, zk_dptr src, uint64_t bytes);

/* ---- unit-level entry points (parity tests, micro-benchmarks) ------------------------------------ */
/* Poseidon2 width-16 permutation (crates/primitives/src/lib.rs:1107-1121 `poseidon2_init`) of n states. */
int32_t zk_poseidon2_permute(zk_ctx* ctx, uint32_t* states_host, uint64_t n);
/* PaddingFreeSponge<Perm,16,8,8> of every row (crates/stark/src/kb31_poseidon2.rs:173). digests: h*8 words. */
int32_t zk_hash_rows(zk_ctx* ctx, const uint32_t* mat_host, uint64_t h, uint32_t w, uint32_t* digests_host);
/* TruncatedPermutation<Perm,2,8,16> of n_out adjacent digest pairs (kb31_poseidon2.rs:175). */
int32_t zk_compress_layer(zk_ctx* ctx, const uint32_t* prev_host, uint64_t n_out, uint32_t* out_host);
/* TwoAdicSubgroupDft::dft_batch: natural-order DFT of every column (out[k] = sum_j in[j] g^(jk)). */
int32_t zk_dft_batch(zk_ctx* ctx, const uint32_t* in_host, uint64_t h, uint32_t w, uint32_t* out_host);
/* TwoAdicSubgroupDft::coset_lde_batch(in, log_blowup, shift).bit_reverse_rows(). */
int32_t zk_coset_lde(zk_ctx* ctx, const uint32_t* in_host, uint64_t h, uint32_t w, uint32_t log_blowup,
                     uint32_t shift, uint32_t* out_host);
int32_t zk_coset_lde_dev(zk_ctx* ctx, zk_dptr in, uint64_t h, uint32_t w, uint32_t log_blowup, uint32_t shift,
                         zk_dptr out);

/* ---- Pcs::commit (TwoAdicFriPcs::commit; call sites crates/stark/src/prover.rs:277,403,497 and
 *      crates/stark/src/machine.rs:397) -------------------------------------------------------------
 * For every (domain, evals): LDE with shift GENERATOR/domain_shift, rows bit-reversed; then one mixed-
 * height Poseidon2 Merkle tree over all LDEs.  root = Com (8 words).  The LDEs stay on the device. */
int32_t zk_commit(zk_ctx* ctx, uint32_t n_mats, const uint32_t* const* mats_host, const uint64_t* heights,
                  const uint32_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup, uint32_t root[8],
                  zk_pdata** out);
/* Same with the traces already resident in HBM (device trace generation, or uploaded by zk_h2d). */
int32_t zk_commit_dev(zk_ctx* ctx, uint32_t n_mats, const zk_dptr* mats_dev, const uint64_t* heights,
                      const uint32_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup,
                      uint32_t root[8], zk_pdata** out);
/* Mmcs::commit without LDE (MerkleTreeMmcs::commit; used for FRI layers through ExtensionMmcs). */
int32_t zk_mmcs_commit(zk_ctx* ctx, uint32_t n_mats, const uint32_t* const* mats_host, const uint64_t* heights,
                       const uint32_t* widths, uint32_t root[8], zk_pdata** out);
int32_t zk_mmcs_commit_dev(zk_ctx* ctx, uint32_t n_mats, const zk_dptr* mats_dev, const uint64_t* heights,
                           const uint32_t* widths, uint32_t root[8], zk_pdata** out);

/* When enabled, zk_commit keeps a device copy of every input trace in the prover data (zk_pdata_trace), so
 * that stages between commits which read the trace itself (LogUp) need no second upload. */
int32_t zk_ctx_keep_traces(zk_ctx* ctx, int32_t enable);
/* Upload helper: `device` is an IDLE peer GPU of the same node (NVLink peer access required).  From then on zk_commit
 * sends the second half of the rows of every trace slab over THAT GPU's PCIe link into a staging buffer there and
 * forwards it over NVLink, so a rank on a partly used node gets two links' worth of host bandwidth (measured 109 vs
 * 55.6 GB/s, profiles/r2_h2d_multi_probe.txt).  device < 0 turns it off.  Host buffers must be pinned. */
int32_t zk_ctx_set_upload_helper(zk_ctx* ctx, int32_t device);

/* ---- ProverData accessors -------------------------------------------------------------------------- */
void zk_pdata_free(zk_pdata* pd);
uint32_t zk_pdata_num_matrices(const zk_pdata* pd);
uint64_t zk_pdata_height(const zk_pdata* pd, uint32_t i); /* committed (LDE) height */
uint32_t zk_pdata_width(const zk_pdata* pd, uint32_t i);
uint32_t zk_pdata_log_max_height(const zk_pdata* pd);
int32_t zk_pdata_root(const zk_pdata* pd, uint32_t root[8]);
/* Pcs::get_evaluations_on_domain (crates/stark/src/prover.rs:437-445): the committed LDE, no copy.
 * Row r of the returned matrix is the evaluation at GENERATOR * g^bitrev(r). */
zk_dptr zk_pdata_lde(const zk_pdata* pd, uint32_t i);
/* row stride of that matrix in WORDS: the width rounded up to even (rows of a committed LDE start 8-byte aligned, so
 * that 47-, 115-, 119-column chips take the same vector paths as even widths); columns >= width are padding. */
uint32_t zk_pdata_pitch(const zk_pdata* pd, uint32_t i);
/* the retained input trace (natural order, height >> log_blowup rows), or 0 when traces were not kept */
zk_dptr zk_pdata_trace(const zk_pdata* pd, uint32_t i);
/* Mmcs::get_matrices / Serialize support: copy an LDE matrix or a digest layer (0 = leaves) to the host. */
int32_t zk_pdata_copy_lde(const zk_pdata* pd, uint32_t i, uint32_t* out_host);
int32_t zk_pdata_copy_layer(const zk_pdata* pd, uint32_t layer, uint32_t* out_host);
/* Deserialize support (`PcsProverData<SC>: Serialize + DeserializeOwned`, crates/stark/src/prover.rs:221,
 * crates/stark/src/machine.rs:56-57: the proving key holds the preprocessed round's ProverData and is serialised):
 * the inverse of zk_pdata_copy_lde / zk_pdata_copy_layer.  heights are COMMITTED heights; layers_host[l] is digest layer
 * l (0 = leaves), n_layers = log2(max height) + 1.  traces_host (nullable, entries nullable) restores the retained
 * input traces of zk_ctx_keep_traces (height >> log_blowup rows each; `StarkProvingKey::traces`).  Nothing is
 * re-hashed: like serde, it restores exactly what was exported. */
int32_t zk_pdata_import(zk_ctx* ctx, uint32_t n_mats, const uint32_t* const* ldes_host, const uint64_t* heights,
                        const uint32_t* widths, const uint32_t* const* layers_host, uint32_t n_layers,
                        const uint32_t* const* traces_host, uint32_t log_blowup, zk_pdata** out);
/* Mmcs::open_batch for n_idx indices at once (crates/recursion/circuit/src/fri.rs:383-387):
 * opened: for each index, the rows of all matrices back to back in matrix order (sum of widths words);
 * proofs: for each index, log_max_height siblings of 8 words, bottom-up. */
int32_t zk_pdata_open_batch(const zk_pdata* pd, uint32_t n_idx, const uint64_t* indices, uint32_t* opened_host,
                            uint32_t* proofs_host);

/* ---- quotient (quotient_values, crates/stark/src/quotient.rs:19-171) ---------------------------------
 * Chips are compiled into the library as generated kernels (zkmips_b200/air/codegen.py); a chip is
 * addressed by the index of its name.  zk_quotient evaluates, for every point of the quotient domain
 * (size 2^(log_degree + log_quotient_degree), shift GENERATOR), the folded constraints
 * sum_k alpha^(n-1-k) C_k times inv_zeroifier, reading local/next rows directly from the committed LDEs
 * (Pcs::get_evaluations_on_domain without the copy), and writes the 2^log_quotient_degree quotient chunks
 * (split_evals, crates/stark/src/prover.rs:477-488) as N x 4 base matrices, back to back, into a device
 * buffer the caller frees with zk_dev_free (after committing it with zk_commit_dev, domain shift of
 * chunk c = GENERATOR * g_{n+lqd}^c).  The 11 data arguments mirror quotient_values' parameters. */
typedef struct {
  uint32_t main_width, prep_width, perm_width /* extension columns */, num_public_values, num_challenges;
  uint32_t num_constraints, max_degree, num_kernels, num_lookups;
} zk_air_desc;
int32_t zk_air_count(void);
const char* zk_air_name(int32_t id);
int32_t zk_air_find(const char* name); /* -1 when unknown */
int32_t zk_air_info(int32_t id, zk_air_desc* out);
int32_t zk_quotient(zk_ctx* ctx, int32_t air_id, const zk_pdata* prep, uint32_t prep_idx, const zk_pdata* main_data,
                    uint32_t main_idx, const zk_pdata* perm, uint32_t perm_idx, uint32_t log_degree,
                    uint32_t log_quotient_degree, const uint32_t alpha[4], const uint32_t* perm_challenges,
                    const uint32_t* public_values, uint32_t n_public_values, const uint32_t local_cumsum[4],
                    const uint32_t global_cumsum[14], zk_dptr* out_chunks);

/* generate_permutation_trace (crates/stark/src/permutation.rs:102-196; call site prover.rs:341-364) on the
 * device: LogUp batch entries and the running sum for a chip with lookups.  Inputs are the natural-order
 * traces on the device (zk_pdata_trace after zk_ctx_keep_traces(ctx, 1), or buffers uploaded with zk_h2d).
 * out_trace: height x (4 * perm_width) base matrix (`flatten_to_base`, prover.rs:393), freed with zk_dev_free
 * after it has been committed with zk_commit_dev. */
int32_t zk_permutation_trace(zk_ctx* ctx, int32_t air_id, zk_dptr prep_trace, zk_dptr main_trace, uint64_t height,
                             const uint32_t perm_challenges[8], zk_dptr* out_trace, uint32_t local_cumsum[4]);

/* ---- device trace generation ---------------------------------------------------------------------------------------
 * Batched device twins of the per-row trace fillers the reference calls through its own C FFI:
 *   poseidon2_wide_event_to_row_koalabear / poseidon2_wide_instr_to_row_koalabear (crates/recursion/core/src/sys.rs:104-112,
 *     C++ crates/recursion/core/include/poseidon2_wide.hpp:148-208; callers chips/poseidon2_wide/trace.rs:76-108,183-262),
 *   add_sub_event_to_row_koalabear (crates/core/machine/src/sys.rs:23; C++ crates/core/machine/include/add_sub.hpp:24-39),
 * and of the Rust `event_to_row` of BitwiseChip and LtChip (crates/core/machine/src/alu/bitwise/mod.rs:141-170,
 * alu/lt/mod.rs:179-262).  One call fills the whole padded trace (rows = power of two >= n_events; padding exactly as
 * `generate_trace` pads: zero rows for the ALU chips, the row of the all-zero input for Poseidon2) in a device buffer
 * that goes straight into zk_commit_dev and is freed with zk_dev_free: only the events cross PCIe.
 * The `_dev` variants read events already on the device. */
typedef struct { /* AluEvent, #[repr(C)]: crates/core/executor/src/events/instr.rs:10-26 */
  uint32_t pc, next_pc;
  uint8_t opcode; /* crates/core/executor/src/opcode.rs:14-76 */
  uint8_t pad_[3];
  uint32_t hi, a, b, c;
} zk_alu_event;
enum { ZK_CHIP_ADD_SUB = 0, ZK_CHIP_BITWISE = 1, ZK_CHIP_LT = 2, ZK_CHIP_SHIFT_LEFT = 3, ZK_CHIP_SHIFT_RIGHT = 4,
       ZK_CHIP_CLO_CLZ = 5 }; /* alu/add_sub, bitwise, lt, sll, sr, clo_clz: every chip whose events are AluEvents */
uint32_t zk_tracegen_alu_width(int32_t chip); /* 19 / 18 / 36 / 44 / 71 / 22 main columns; 0 for an unknown chip */
int32_t zk_tracegen_alu(zk_ctx* ctx, int32_t chip, const zk_alu_event* events_host, uint64_t n_events, uint64_t rows,
                        zk_dptr* out_trace);
int32_t zk_tracegen_alu_dev(zk_ctx* ctx, int32_t chip, zk_dptr events_dev, uint64_t n_events, uint64_t rows,
                            zk_dptr* out_trace);
/* CpuChip (crates/core/machine/src/cpu/trace.rs:118-237, columns cpu/columns/mod.rs:16-62): one 67-column row per executed
 * instruction from a packed event: the CpuEvent's scalars, its instruction and the previous (shard, clk) of the three register
 * accesses (a_record / b_record / c_record; ignored for an immediate operand).  flags: bit 0 op_a_0, 1 imm_b, 2 imm_c,
 * 3 is_rw_a, 4 is_check_memory, 5 is_halt, 6 is_sequential, 7 op_a_immutable.  Padding rows as generate_trace leaves them
 * (imm_b = imm_c = is_rw_a = 1). */
typedef struct {
  uint32_t pc, next_pc, next_next_pc, clk, shard, opcode, op_a, op_b, op_c, flags, num_extra_cycles;
  uint32_t a, b, c, hi;
  uint32_t a_prev_value, a_prev_shard, a_prev_clk, b_prev_shard, b_prev_clk, c_prev_shard, c_prev_clk;
} zk_cpu_event;
uint32_t zk_tracegen_cpu_width(void); /* 67 */
int32_t zk_tracegen_cpu(zk_ctx* ctx, const zk_cpu_event* events_host, uint64_t n_events, uint64_t rows, zk_dptr* out_trace);
int32_t zk_tracegen_cpu_dev(zk_ctx* ctx, zk_dptr events_dev, uint64_t n_events, uint64_t rows, zk_dptr* out_trace);
/* Poseidon2WideChip<DEGREE>: inputs = n_events x 16 Montgomery words (Poseidon2Event::input); sbox_state = 1 for
 * DEGREE 3 (313 columns), 0 for DEGREE 9 (172 columns). */
uint32_t zk_tracegen_poseidon2_wide_width(int32_t sbox_state);
int32_t zk_tracegen_poseidon2_wide(zk_ctx* ctx, const uint32_t* inputs_host, uint64_t n_events, uint64_t rows,
                                   int32_t sbox_state, zk_dptr* out_trace);
int32_t zk_tracegen_poseidon2_wide_dev(zk_ctx* ctx, zk_dptr inputs_dev, uint64_t n_events, uint64_t rows,
                                       int32_t sbox_state, zk_dptr* out_trace);
/* preprocessed trace: instrs = n x 48 words (Poseidon2SkinnyInstr: input addrs[16], output addrs[16], mults[16]) ->
 * rows x 49 (Poseidon2PreprocessedColsWide); padding rows are zero */
int32_t zk_tracegen_poseidon2_wide_prep(zk_ctx* ctx, const uint32_t* instrs_host, uint64_t n, uint64_t rows,
                                        zk_dptr* out_trace);
/* Poseidon2SkinnyChip<DEGREE> (the wrap machine's Poseidon2 chip; crates/recursion/core/src/sys.rs binds
 * poseidon2_skinny_event_to_row_koalabear / poseidon2_skinny_instr_to_row_koalabear -> include/poseidon2_skinny.hpp:50-115):
 * ELEVEN rows of 28 (main) / 51 (preprocessed) words per permutation, rows = power of two >= 11 * n_events, zero padding. */
uint32_t zk_tracegen_poseidon2_skinny_width(void);
int32_t zk_tracegen_poseidon2_skinny(zk_ctx* ctx, const uint32_t* inputs_host, uint64_t n_events, uint64_t rows,
                                     zk_dptr* out_trace);
int32_t zk_tracegen_poseidon2_skinny_dev(zk_ctx* ctx, zk_dptr inputs_dev, uint64_t n_events, uint64_t rows,
                                         zk_dptr* out_trace);
int32_t zk_tracegen_poseidon2_skinny_prep(zk_ctx* ctx, const uint32_t* instrs_host, uint64_t n, uint64_t rows,
                                          zk_dptr* out_trace);

/* ---- DuplexChallenger<Val, Perm, 16, 8> (crates/stark/src/kb31_poseidon2.rs:180; semantics restated at
 *      crates/recursion/circuit/src/challenger.rs:90-233) ------------------------------------------------
 * The 34-word image of a Plonky3 DuplexChallenger: sponge_state, input_buffer (+ length), output_buffer
 * (+ length; samples pop from the END).  The Rust side converts its challenger to/from this struct around
 * zk_pcs_open; the zk_challenger_* calls run the same transcript operations on the device. */
typedef struct {
  uint32_t state[16];
  uint32_t in[8];
  uint32_t n_in;
  uint32_t out[8];
  uint32_t n_out;
} zk_challenger;
int32_t zk_challenger_init(zk_challenger* ch);
int32_t zk_challenger_observe(zk_ctx* ctx, zk_challenger* ch, const uint32_t* vals, uint32_t n);
int32_t zk_challenger_sample_ext(zk_ctx* ctx, zk_challenger* ch, uint32_t n_ext, uint32_t* out /* 4*n_ext */);
int32_t zk_challenger_sample_bits(zk_ctx* ctx, zk_challenger* ch, uint32_t bits, uint32_t n, uint64_t* out);
/* DuplexChallenger::grind: finds the SMALLEST canonical witness (Plonky3 takes any), applies it to ch. */
int32_t zk_challenger_grind(zk_ctx* ctx, zk_challenger* ch, uint32_t bits, uint32_t* witness);

/* ---- Pcs::open (TwoAdicFriPcs::open; call site crates/stark/src/prover.rs:546-556) --------------------
 * rounds[r] = prover data of round r (e.g. preprocessed, main, permutation, quotient); n_points[k] and
 * points (4 words each, flattened) give the opening points of matrix k, matrices counted round by round
 * in commit order.  Steps: alpha = sample_ext; opened values by barycentric evaluation;
 * reduced openings per height class; FRI commit phase (fold, ExtensionMmcs commit, observe, beta);
 * final_poly; proof-of-work; num_queries query openings.
 *
 * Flat proof layout (u32 words), shared with the CPU oracle:
 *   opened values : for round, matrix, point: width extension elements
 *   fri           : n_layers * 8 commit-phase roots, final_poly (4), pow_witness (1)
 *   queries       : num_queries x { per round: opened rows (sum of widths) + path (log_max_r * 8);
 *                                   per layer i: sibling value (4) + path ((log_max - i - 1) * 8) }
 * with log_max = log2 of the tallest committed matrix and n_layers = log_max - log_blowup.
 * inject_witness >= 0 uses that (Montgomery) pow witness instead of grinding, so a transcript produced by
 * the reference prover (whose grind is non-deterministic) can be reproduced bit for bit.
 * ch is advanced exactly as the reference's challenger would be. */
uint64_t zk_pcs_proof_words(uint32_t n_rounds, const zk_pdata* const* rounds, const uint32_t* n_points,
                            uint32_t log_blowup, uint32_t num_queries);
int32_t zk_pcs_open(zk_ctx* ctx, uint32_t n_rounds, const zk_pdata* const* rounds, const uint32_t* n_points,
                    const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                    zk_challenger* ch, int64_t inject_witness, uint32_t* proof_host, uint64_t proof_cap);

#ifdef __cplusplus
}
#endif
#endif
