// oracle/ref_shim_core.cpp -- TEST INFRASTRUCTURE.  extern "C" export over the REFERENCE's own core-machine row filler,
// compiled where it lies under /root/reference by oracle/Makefile (target `ref`) into oracle/_ref/libzkref.so:
//   crates/core/machine/include/add_sub.hpp   (event_to_row :24-39, AddOperation populate :8-22)
//   crates/core/machine/include/utils.hpp     (write_word_from_u32_v2 :63-69)
// The cbindgen-generated header those include is replaced by a stub the Makefile writes into oracle/_ref: the
// #[repr(C)] structs of crates/core/executor/src/events/instr.rs:10-26 (AluEvent),
// crates/core/machine/src/alu/add_sub/mod.rs:41-62 (AddSubCols), operations/add.rs:11-19 (AddOperation),
// crates/stark/src/word.rs (Word) and the Opcode enum extracted from crates/core/executor/src/opcode.rs by sed.
#include "kb31_t.hpp"
#include "add_sub.hpp"
#include <cstdint>

using namespace zkm_core_machine_sys;

extern "C" void ref_add_sub_event_to_row(const uint32_t event[7], uint32_t cols[19]) {
  static_assert(sizeof(AluEvent) == 28 && sizeof(AddSubCols<kb31_t>) == 19 * 4);
  add_sub::event_to_row<kb31_t>(*reinterpret_cast<const AluEvent*>(event), *reinterpret_cast<AddSubCols<kb31_t>*>(cols));
}
