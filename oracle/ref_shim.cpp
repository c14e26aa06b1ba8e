// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" exports over the REFERENCE's own C++
// sources, compiled where they lie under /root/reference by oracle/Makefile (target `ref`), output
// oracle/_ref/libzkref.so.  Nothing from the reference is copied into this repository: this file only
// includes
//   crates/core/machine/include/kb31_t.hpp            (field class, host branch :458-623)
//   crates/recursion/core/include/poseidon2_wide.hpp   (populate_perm :96-146)
//   crates/recursion/core/include/poseidon2_skinny.hpp (event_to_row :50-76, instr_to_row :78-115)
//   crates/recursion/core/include/poseidon2.hpp        (linear layers :21-71)
//   crates/recursion/core/include/poseidon2_constants.hpp
// through a shim for the cbindgen-generated header they expect (written by the Makefile; it carries
// the constants of crates/recursion/core/src/chips/poseidon2_wide/mod.rs:18-23 and
// chips/poseidon2_skinny/trace.rs:47).
#include "kb31_t.hpp"
#include "poseidon2_wide.hpp"
#include "poseidon2_skinny.hpp"
#include <cstdint>

using namespace zkm_recursion_core_sys;

extern "C" {
uint32_t ref_add(uint32_t a, uint32_t b) { return (kb31_t(a) + kb31_t(b)).val; }
uint32_t ref_sub(uint32_t a, uint32_t b) { return (kb31_t(a) - kb31_t(b)).val; }
uint32_t ref_mul(uint32_t a, uint32_t b) { return (kb31_t(a) * kb31_t(b)).val; }
uint32_t ref_inv(uint32_t a) { return kb31_t(a).reciprocal().val; }
uint32_t ref_to_monty(uint32_t c) { return kb31_t::to_monty(c); }
uint32_t ref_from_monty(uint32_t m) { return kb31_t::from_monty(m); }
// Poseidon2 permutation exactly as the reference's trace filler computes it (Montgomery in/out).
void ref_poseidon2_permute(uint32_t state[16]) {
  kb31_t in[WIDTH], ext_state[WIDTH * NUM_EXTERNAL_ROUNDS], int_state[WIDTH],
      s0[NUM_INTERNAL_ROUNDS - 1], ext_sbox[WIDTH * NUM_EXTERNAL_ROUNDS], int_sbox[NUM_INTERNAL_ROUNDS],
      out[WIDTH];
  for (size_t i = 0; i < WIDTH; i++) in[i] = kb31_t(state[i]);
  poseidon2_wide::populate_perm<kb31_t>(in, ext_state, int_state, s0, ext_sbox, int_sbox, out);
  for (size_t i = 0; i < WIDTH; i++) state[i] = out[i].val;
}
}

// ---- trace fillers (the functions crates/recursion/core/src/sys.rs:104-113 binds) ---------------------------------
extern "C" {
// poseidon2_wide::event_to_row (poseidon2_wide.hpp:148-197): one main-trace row of Poseidon2WideChip<DEGREE>
// (313 words with the S-box columns = DEGREE 3, 172 without = DEGREE 9) from the 16-word permutation input.
void ref_poseidon2_wide_event_to_row(const uint32_t input[16], uint32_t* row, int sbox_state) {
  kb31_t in[WIDTH];
  for (size_t i = 0; i < WIDTH; i++) in[i] = kb31_t(input[i]);
  poseidon2_wide::event_to_row<kb31_t>(in, reinterpret_cast<kb31_t*>(row), 0, 1, sbox_state != 0);
}
// poseidon2_wide::instr_to_row (poseidon2_wide.hpp:199-208): instr = input addrs[16], output addrs[16], mults[16]
// (Poseidon2SkinnyInstr); cols = input[16], output[16] x {addr, mult}, is_real_neg (49 words).
void ref_poseidon2_wide_instr_to_row(const uint32_t instr[48], uint32_t cols[49]) {
  static_assert(sizeof(Poseidon2SkinnyInstr<kb31_t>) == 48 * 4 && sizeof(Poseidon2PreprocessedColsWide<kb31_t>) == 49 * 4);
  poseidon2_wide::instr_to_row<kb31_t>(*reinterpret_cast<const Poseidon2SkinnyInstr<kb31_t>*>(instr),
                                       *reinterpret_cast<Poseidon2PreprocessedColsWide<kb31_t>*>(cols));
}
}

// ---- Poseidon2SkinnyChip fillers (crates/recursion/core/src/sys.rs binds poseidon2_skinny_event_to_row_koalabear /
// poseidon2_skinny_instr_to_row_koalabear) -----------------------------------------------------------------------------
extern "C" {
// poseidon2_skinny::event_to_row (poseidon2_skinny.hpp:50-76): the ELEVEN main-trace rows (28 words each) of one
// permutation from its 16-word input.
void ref_poseidon2_skinny_event_to_rows(const uint32_t input[16], uint32_t rows[11 * 28]) {
  static_assert(sizeof(Poseidon2<kb31_t>) == 28 * 4 && OUTPUT_ROUND_IDX + 1 == 11);
  Poseidon2Event<kb31_t> ev{};
  for (size_t i = 0; i < WIDTH; i++) ev.input[i] = kb31_t(input[i]);
  poseidon2_skinny::event_to_row<kb31_t>(ev, reinterpret_cast<Poseidon2<kb31_t>*>(rows));
}
// poseidon2_skinny::instr_to_row (poseidon2_skinny.hpp:78-115): preprocessed row i (51 words) of an instruction.
void ref_poseidon2_skinny_instr_to_row(const uint32_t instr[48], uint64_t i, uint32_t cols[51]) {
  static_assert(sizeof(Poseidon2PreprocessedColsSkinny<kb31_t>) == 51 * 4);
  poseidon2_skinny::instr_to_row<kb31_t>(*reinterpret_cast<const Poseidon2Instr<kb31_t>*>(instr), i,
                                         *reinterpret_cast<Poseidon2PreprocessedColsSkinny<kb31_t>*>(cols));
}
}
