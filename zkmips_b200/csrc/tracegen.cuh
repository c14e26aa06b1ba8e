// tracegen.cuh -- main / preprocessed trace rows filled on the device from event records: batched twins of the per-row
// fillers the reference calls through its own C FFI (crates/recursion/core/src/sys.rs:20-113 ->
// crates/recursion/core/include/*.hpp, crates/core/machine/src/sys.rs:14-42 -> crates/core/machine/include/*.hpp) and
// of the Rust `event_to_row` of three ALU chips.  The reference fills rows on the CPU and the prover then uploads the
// whole trace; here only the events cross PCIe (64 B per Poseidon2 permutation instead of a 1252-byte row, 28 B per ALU
// event instead of 76..144 B) and the rows are written straight into the buffer zk_commit_dev reads.
//
// All kernels: one thread per row; whole rows are staged in shared memory and leave the CTA as ONE contiguous run (a
// block of consecutive rows of a dense row-major matrix is one contiguous range), so HBM sees full-sector writes
// whatever the row width.  HBM-bound (a row of 313 words costs ONE permutation; hashing it costs 40).
#pragma once
#include "kb31.cuh"
#include "poseidon2.cuh"

namespace tg {

constexpr int ROWS = 128;  // rows (= threads) per CTA

// ---- Poseidon2WideChip<DEGREE>::generate_trace (crates/recursion/core/src/chips/poseidon2_wide/trace.rs:76-108,
//      populate_perm :271-330; C++ twin poseidon2_wide.hpp:96-197).  Row layout (columns/permutation.rs:20-35):
//      external_rounds_state[8][16] | internal_rounds_state[16] | internal_rounds_s0[12] | output_state[16]
//      | external_rounds_sbox[8][16] | internal_rounds_sbox[13]       -- 172 words (DEGREE 9), 313 with S-boxes (DEGREE 3)
//      Padding rows are the row of the all-zero input (trace.rs:97-103), not zeros.
constexpr uint32_t P2W_INT_STATE = 128, P2W_S0 = 144, P2W_OUT = 156, P2W_EXT_SBOX = 172, P2W_INT_SBOX = 300;
constexpr uint32_t P2W_WIDTH_NO_SBOX = 172, P2W_WIDTH_SBOX = 313;

// One WARP per CTA, one row per lane, and the CTA's 32 whole rows staged in shared memory (40 KB with the S-box columns:
// five CTAs per SM) before they leave as ONE contiguous, 16-byte aligned run of 128-bit stores -- every 32-byte sector
// of the trace is written exactly once, completely.  The first version flushed each 16-column group of 128 rows as it
// was produced: rows are 1252 bytes, so seven rows in eight start inside a sector and every group wrote two partial
// sectors per row; ncu: 44 % of warp samples in MIO throttle behind the store path, 116 MB of DRAM reads (sector fills)
// for a kernel that reads 15 MB, 0.24 ms for 2^18 rows (profiles/r2_ncu_tracegen.txt).
constexpr int P2W_ROWS = 32;

template <bool SBOX>
__global__ void __launch_bounds__(P2W_ROWS) poseidon2_wide_rows(const uint32_t* __restrict__ inputs, uint64_t n_events,
                                                                uint64_t rows, uint32_t* __restrict__ out) {
  constexpr uint32_t W = SBOX ? P2W_WIDTH_SBOX : P2W_WIDTH_NO_SBOX;
  // W = 313: lanes hit distinct banks per column; W = 172: 4-way conflicts on the 172 staging stores (noise beside
  // the permutation)
  __shared__ __align__(16) uint32_t tile[P2W_ROWS * W];
  const uint32_t lane = threadIdx.x;
  const uint64_t row0 = (uint64_t)blockIdx.x * P2W_ROWS;
  uint32_t* const t = tile + lane * W;
  uint32_t s[16];
  if (row0 + lane < n_events) {
    const uint4* in = reinterpret_cast<const uint4*>(inputs + (row0 + lane) * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      uint4 v = in[q];
      s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = 0u;  // padding rows: the row of the all-zero input
  }
  auto put = [&](uint32_t col) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[col + i] = s[i];
  };
  put(0);  // external_rounds_state[0] = input
  p2::external_layer(s);
#pragma unroll 1
  for (int r = 0; r < 8; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = p2::sbox(s[i], p2::EXT_RC[r][i]);
    if (SBOX) put(P2W_EXT_SBOX + 16 * r);
    p2::external_layer(s);
    if (r == 3) {
      put(P2W_INT_STATE);
#pragma unroll 1
      for (int k = 0; k < 13; k++) {
        s[0] = p2::sbox(s[0], p2::INT_RC[k]);
        if (SBOX) t[P2W_INT_SBOX + k] = s[0];
        p2::internal_layer(s);
        if (k < 12) t[P2W_S0 + k] = s[0];
      }
      put(16 * 4);  // external_rounds_state[4] = state after the internal rounds
    } else if (r == 7) {
      put(P2W_OUT);
    } else {
      put(16 * (r + 1));
    }
  }
  __syncthreads();
  if (row0 + P2W_ROWS <= rows) {
    // 32 rows x W words = a multiple of 4 words, starting 16-byte aligned (32 * W * 4 bytes per CTA)
    uint4* o = reinterpret_cast<uint4*>(out + row0 * W);
    const uint4* src = reinterpret_cast<const uint4*>(tile);
#pragma unroll 4
    for (uint32_t k = lane; k < P2W_ROWS * W / 4; k += P2W_ROWS) o[k] = src[k];
  } else {  // a trace shorter than one CTA
    for (uint64_t k = lane; k < (rows - row0) * W; k += P2W_ROWS) out[row0 * W + k] = tile[k];
  }
}

// ---- generate_preprocessed_trace (trace.rs:183-216; instr_to_row poseidon2_wide.hpp:199-208):
//      instr = input addrs[16], output addrs[16], mults[16]  ->  input[16], output[16] x {addr, mult}, is_real_neg = -1
__global__ void poseidon2_wide_prep_rows(const uint32_t* __restrict__ instrs, uint64_t n, uint64_t rows,
                                         uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 49) return;
  uint64_t r = i / 49;
  uint32_t c = (uint32_t)(i % 49);
  uint32_t v = 0;
  if (r < n) {
    const uint32_t* in = instrs + r * 48;
    if (c < 16) v = in[c];
    else if (c < 48) v = ((c - 16) & 1) ? in[32 + ((c - 16) >> 1)] : in[16 + ((c - 16) >> 1)];
    else v = kb::P - kb::ONE;
  }
  out[i] = v;
}

// ---- Poseidon2SkinnyChip<DEGREE>::generate_trace (crates/recursion/core/src/chips/poseidon2_skinny/trace.rs:77-130; C++
//      twin poseidon2_skinny.hpp:50-76): ELEVEN rows of 28 words per permutation -- state_var[16] | internal_rounds_s0[12]
//      (columns/mod.rs:20-26).  Row 0: the input; rows 1..4: the state entering external rounds 0..3; row 5: the state
//      entering the 13 internal rounds, with s0 of rounds 0..11; rows 6..9: external rounds 4..7; row 10: the output.
//      Padding rows (>= 11 * n_events) are ZERO (trace.rs:126 `rows.resize(.., [F::ZERO; ..])`), unlike the wide chip.
//      One warp per CTA, one permutation per lane: the CTA's 32 x 11 rows are one contiguous run of 9856 words, staged in
//      shared memory (38.5 KB) and written with 128-bit stores like the wide filler above.
constexpr uint32_t P2S_WIDTH = 28, P2S_ROWS_PER_EVENT = 11, P2S_EVENTS = 32;
constexpr uint32_t P2S_TILE = P2S_EVENTS * P2S_ROWS_PER_EVENT * P2S_WIDTH;  // words per CTA

__global__ void __launch_bounds__(P2S_EVENTS) poseidon2_skinny_rows(const uint32_t* __restrict__ inputs, uint64_t n_events,
                                                                    uint64_t rows, uint32_t* __restrict__ out) {
  __shared__ __align__(16) uint32_t tile[P2S_TILE];
  const uint32_t lane = threadIdx.x;
  const uint64_t ev = (uint64_t)blockIdx.x * P2S_EVENTS + lane;
  uint32_t* const t = tile + lane * (P2S_ROWS_PER_EVENT * P2S_WIDTH);
#pragma unroll 4
  for (uint32_t k = 0; k < P2S_ROWS_PER_EVENT * P2S_WIDTH; k++) t[k] = 0u;
  if (ev < n_events) {
    uint32_t s[16];
    const uint4* in = reinterpret_cast<const uint4*>(inputs + ev * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      uint4 v = in[q];
      s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
    }
    auto put = [&](uint32_t row) {
#pragma unroll
      for (int i = 0; i < 16; i++) t[row * P2S_WIDTH + i] = s[i];
    };
    put(0);
    p2::external_layer(s);
    put(1);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = p2::sbox(s[i], p2::EXT_RC[r][i]);
      p2::external_layer(s);
      if (r == 3) {
        put(5);
#pragma unroll 1
        for (int k = 0; k < 13; k++) {
          s[0] = p2::sbox(s[0], p2::INT_RC[k]);
          p2::internal_layer(s);
          if (k < 12) t[5 * P2S_WIDTH + 16 + k] = s[0];
        }
        put(6);
      } else {
        put(r < 3 ? r + 2 : r + 3);
      }
    }
  }
  __syncthreads();
  const uint64_t row0 = (uint64_t)blockIdx.x * (P2S_EVENTS * P2S_ROWS_PER_EVENT);
  if (row0 >= rows) return;
  if (row0 + P2S_EVENTS * P2S_ROWS_PER_EVENT <= rows) {
    uint4* o = reinterpret_cast<uint4*>(out + row0 * P2S_WIDTH);  // 9856 words per CTA: 16-byte aligned
    const uint4* src = reinterpret_cast<const uint4*>(tile);
#pragma unroll 4
    for (uint32_t k = lane; k < P2S_TILE / 4; k += P2S_EVENTS) o[k] = src[k];
  } else {
    for (uint64_t k = lane; k < (rows - row0) * P2S_WIDTH; k += P2S_EVENTS) out[row0 * P2S_WIDTH + k] = tile[k];
  }
}

// ---- Poseidon2SkinnyChip::generate_preprocessed_trace (trace.rs:184-251; instr_to_row poseidon2_skinny.hpp:78-115): eleven
//      51-word rows per instruction (input addrs[16], output addrs[16], mults[16]): memory_preprocessed[16] x {addr, mult}
//      (input row: the input addresses with multiplicity -1; output row: output addresses and their multiplicities), then
//      is_input_round, is_external_round, is_internal_round, round_constants[16] (external rows: RC[round][0..16]; the
//      internal row: RC[4 + j][0], j < 16 -- the 13 internal constants, then column 0 of external rounds 4, 5, 6).
__global__ void poseidon2_skinny_prep_rows(const uint32_t* __restrict__ instrs, uint64_t n, uint64_t rows,
                                           uint32_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 51) return;
  const uint64_t r = i / 51;
  const uint32_t c = (uint32_t)(i % 51);
  const uint64_t e = r / 11;
  const uint32_t k = (uint32_t)(r % 11);
  uint32_t v = 0;
  if (e < n) {
    const uint32_t* in = instrs + e * 48;
    const bool ext = k != 0 && k != 5 && k != 10;
    if (c < 32) {
      if (k == 0) v = (c & 1) ? kb::P - kb::ONE : in[c >> 1];
      else if (k == 10) v = (c & 1) ? in[32 + (c >> 1)] : in[16 + (c >> 1)];
    } else if (c == 32) v = k == 0 ? kb::ONE : 0u;
    else if (c == 33) v = ext ? kb::ONE : 0u;
    else if (c == 34) v = k == 5 ? kb::ONE : 0u;
    else {
      const uint32_t j = c - 35;
      if (ext) v = p2::EXT_RC[k < 5 ? k - 1 : k - 2][j];
      else if (k == 5) v = j < 13 ? p2::INT_RC[j] : p2::EXT_RC[j - 13 + 4][0];
    }
  }
  out[i] = v;
}

// ---- ALU chips from `AluEvent` records (#[repr(C)], crates/core/executor/src/events/instr.rs:10-26: 7 words, the
//      opcode in the low byte of word 2).  Each filler writes CANONICAL values into its zeroed staging row; the flush
//      converts to Montgomery form.  Padding rows (>= n_events) are what the reference's generate_trace leaves there: zero, or the chip's
//      padding row (CHIP::pad).
struct AluEv {
  uint32_t pc, next_pc, opcode, hi, a, b, c;
};

// AddSubChip::event_to_row + AddOperation::populate (crates/core/machine/src/alu/add_sub/mod.rs:150-172,
// operations/add.rs:26-60; C++ twin crates/core/machine/include/add_sub.hpp:8-39)
struct AddSub {
  static constexpr uint32_t W = 19;
  __device__ static void pad(uint32_t*) {}
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    bool is_add = e.opcode == 0;  // Opcode::ADD
    uint32_t op1 = is_add ? e.b : e.a, op2 = e.c, val = op1 + op2, carry = 0;
    t[0] = e.pc;
    t[1] = e.next_pc;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t x = (op1 >> (8 * k)) & 0xFF, y = (op2 >> (8 * k)) & 0xFF;
      t[2 + k] = (val >> (8 * k)) & 0xFF;
      t[9 + k] = x;
      t[13 + k] = y;
      carry = (x + y + carry) > 0xFF;
      if (k < 3) t[6 + k] = carry;
    }
    t[17] = is_add;
    t[18] = !is_add;
  }
};

// BitwiseChip::event_to_row (crates/core/machine/src/alu/bitwise/mod.rs:141-170)
struct Bitwise {
  static constexpr uint32_t W = 18;
  __device__ static void pad(uint32_t*) {}
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    t[0] = e.pc;
    t[1] = e.next_pc;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      t[2 + k] = (e.a >> (8 * k)) & 0xFF;
      t[6 + k] = (e.b >> (8 * k)) & 0xFF;
      t[10 + k] = (e.c >> (8 * k)) & 0xFF;
    }
    t[14] = e.opcode == 18;  // NOR
    t[15] = e.opcode == 17;  // XOR
    t[16] = e.opcode == 16;  // OR
    t[17] = e.opcode == 15;  // AND
  }
};

// LtChip::event_to_row (crates/core/machine/src/alu/lt/mod.rs:179-262)
struct Lt {
  static constexpr uint32_t W = 36;
  __device__ static void pad(uint32_t*) {}
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    const uint32_t is_slt = e.opcode == 13;  // Opcode::SLT (SLTU = 14)
    uint32_t bb[4], cb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      bb[k] = (e.b >> (8 * k)) & 0xFF;
      cb[k] = (e.c >> (8 * k)) & 0xFF;
      t[8 + k] = bb[k];
      t[12 + k] = cb[k];
    }
    uint32_t b_masked = bb[3] & 0x7F, c_masked = cb[3] & 0x7F;
    uint32_t bc3 = is_slt ? b_masked : bb[3], cc3 = is_slt ? c_masked : cb[3];
    uint32_t x = 0, y = 0, sltu = 0, found = 0;
#pragma unroll
    for (int k = 3; k >= 0; k--) {  // most significant differing byte
      uint32_t bk = k == 3 ? bc3 : bb[k], ck = k == 3 ? cc3 : cb[k];
      uint32_t hit = !found && bk != ck;
      t[16 + k] = hit;
      if (hit) { x = bk; y = ck; sltu = bk < ck; found = 1; }
    }
    uint32_t msb_b = bb[3] >> 7, msb_c = cb[3] >> 7;
    uint32_t is_sign_eq = is_slt ? (msb_b == msb_c) : 1u;
    uint32_t bit_b = msb_b & is_slt, bit_c = msb_c & is_slt;
    t[0] = e.pc;
    t[1] = e.next_pc;
    t[2] = is_slt;
    t[3] = !is_slt;
    t[4] = bit_b * (1 - bit_c) + is_sign_eq * sltu;  // a[0]; a[1..3] = 0
    t[20] = b_masked;
    t[21] = c_masked;
    // not_eq_inv = 1 / (comparison_bytes[0] - comparison_bytes[1]) when they differ
    t[22] = found ? kb::from_monty(kb::inv(kb::to_monty(x >= y ? x - y : x + kb::P - y))) : 0u;
    t[23] = msb_b;
    t[24] = msb_c;
    t[25] = bit_b;
    t[26] = bit_c;
    t[27] = sltu;
    t[28] = !found;  // is_comp_eq
    t[29] = is_sign_eq;
    t[30] = x;
    t[31] = y;
    // t[32..36] = byte_equality_check: allocated by the reference, never written
  }
};

// ShiftLeft::event_to_row (crates/core/machine/src/alu/sll/mod.rs:150-215); the padding row is not zero (:95-106): the
// one-hot sums of shift_by_n_bits / shift_by_n_bytes are constrained on every row
struct ShiftLeft {
  static constexpr uint32_t W = 44;
  __device__ static void pad(uint32_t* t) { t[22] = 1; t[30] = 1; t[39] = 1; }
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    t[0] = e.pc;
    t[1] = e.next_pc;
    const uint32_t nbits = e.c & 7, nbytes = (e.c & 31) >> 3, mult = 1u << nbits;
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      t[2 + k] = (e.a >> (8 * k)) & 0xFF;
      t[6 + k] = (e.b >> (8 * k)) & 0xFF;
      t[10 + k] = (e.c >> (8 * k)) & 0xFF;
      uint32_t v = ((e.b >> (8 * k)) & 0xFF) * mult + carry;
      carry = v >> 8;
      t[31 + k] = v & 0xFF;
      t[35 + k] = carry;
      t[39 + k] = nbytes == (uint32_t)k;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      t[14 + i] = (e.c >> i) & 1;
      t[22 + i] = nbits == (uint32_t)i;
    }
    t[30] = mult;
    t[43] = 1;
  }
};

// ShiftRightChip::event_to_row (crates/core/machine/src/alu/sr/mod.rs:153-247): SRL 10 / SRA 11 / ROR 12 over the 8-byte
// extension of b; padding rows have shift_by_n_bits[0] = shift_by_n_bytes[0] = 1 (:108-111)
struct ShiftRight {
  static constexpr uint32_t W = 71;
  __device__ static void pad(uint32_t* t) { t[14] = 1; t[22] = 1; }
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    t[0] = e.pc;
    t[1] = e.next_pc;
    const uint32_t nbits = e.c & 7, nbytes = (e.c & 31) >> 3, msb = e.b >> 31;
    const uint32_t hi = e.opcode == 11 ? (msb ? 0xFFFFFFFFu : 0u) : e.opcode == 12 ? e.b : 0u;
    const uint64_t ext = ((uint64_t)hi << 32) | e.b;
    const uint64_t shifted_bytes = ext >> (8 * nbytes);  // byte_shift_result: bytes above the top are zero
#pragma unroll
    for (int k = 0; k < 4; k++) {
      t[2 + k] = (e.a >> (8 * k)) & 0xFF;
      t[6 + k] = (e.b >> (8 * k)) & 0xFF;
      t[10 + k] = (e.c >> (8 * k)) & 0xFF;
      t[22 + k] = nbytes == (uint32_t)k;
    }
    uint32_t last = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
      const uint32_t byte = (uint32_t)(shifted_bytes >> (8 * i)) & 0xFF;
      const uint32_t sh = byte >> nbits, carry = nbits ? ((byte << (8 - nbits)) & 0xFF) >> (8 - nbits) : 0u;
      t[26 + i] = byte;
      t[34 + i] = (sh + (last << (8 - nbits))) & 0xFF;
      t[42 + i] = carry;
      t[50 + i] = sh;
      last = carry;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      t[14 + i] = nbits == (uint32_t)i;
      t[59 + i] = (e.c >> i) & 1;
    }
    t[58] = msb;
    t[67] = e.opcode == 10;
    t[68] = e.opcode == 12;
    t[69] = e.opcode == 11;
    t[70] = 1;
  }
};

// CloClzChip::generate_trace (crates/core/machine/src/alu/clo_clz/mod.rs:64-128): CLZ 19 / CLO 20; the padding row is CLZ
// of zero (a = 32, is_clz, is_bb_zero)
struct CloClz {
  static constexpr uint32_t W = 22;
  __device__ static void pad(uint32_t* t) { t[2] = 32; t[14] = 1; t[19] = 1; }
  __device__ static void fill(const AluEv& e, uint32_t* t) {
    const bool clo = e.opcode == 20;
    const uint32_t bb = clo ? 0xFFFFFFFFu - e.b : e.b;
    const uint32_t sr1 = bb ? bb >> (31 - min(e.a, 31u)) : 0u;
    t[0] = e.pc;
    t[1] = e.next_pc;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      t[2 + k] = (e.a >> (8 * k)) & 0xFF;
      t[6 + k] = (e.b >> (8 * k)) & 0xFF;
      t[10 + k] = (bb >> (8 * k)) & 0xFF;
      t[15 + k] = (sr1 >> (8 * k)) & 0xFF;
    }
    t[14] = bb == 0;
    t[19] = !clo;
    t[20] = clo;
    t[21] = 1;
  }
};

// ---- CpuChip::event_to_row (crates/core/machine/src/cpu/trace.rs:118-237, MemoryAccessCols::populate_access
//      memory/consistency/trace.rs:69-105) from packed CPU events of 22 words (zk_cpu_event, include/zkgpu.h): the 67-column
//      CPU row of one executed instruction -- shard / clk limbs, pcs, the instruction, the flags, the operand words and
//      the three register-access records (previous shard / clk, same-shard flag, 16 + 8 bit limbs of the timestamp
//      difference).  Padding rows carry imm_b = imm_c = is_rw_a = 1 (trace.rs:60-66).  Canonical values in the staging row;
//      the flush converts to Montgomery form.
constexpr uint32_t CPU_EV_WORDS = 22, CPU_W = 67;
struct CpuEv {
  uint32_t pc, next_pc, next_next_pc, clk, shard, opcode, op_a, op_b, op_c, flags, num_extra_cycles, a, b, c, hi, a_prev_value,
      a_prev_shard, a_prev_clk, b_prev_shard, b_prev_clk, c_prev_shard, c_prev_clk;
};
__device__ __forceinline__ void cpu_put_word(uint32_t* t, uint32_t v) {
#pragma unroll
  for (int k = 0; k < 4; k++) t[k] = (v >> (8 * k)) & 0xFF;
}
// access columns at t: value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb
__device__ __forceinline__ void cpu_put_access(uint32_t* t, uint32_t value, uint32_t shard, uint32_t clk, uint32_t prev_shard,
                                               uint32_t prev_clk) {
  cpu_put_word(t, value);
  const bool same = prev_shard == shard;
  const uint32_t diff = (same ? clk : shard) - (same ? prev_clk : prev_shard) - 1u;
  t[4] = prev_shard;
  t[5] = prev_clk;
  t[6] = same;
  t[7] = diff & 0xFFFF;
  t[8] = (diff >> 16) & 0xFF;
}
__global__ void __launch_bounds__(ROWS) cpu_rows(const uint32_t* __restrict__ events, uint64_t n_events, uint64_t rows,
                                                 uint32_t* __restrict__ out) {
  __shared__ uint32_t tile[ROWS][CPU_W + 1];
  const uint32_t tid = threadIdx.x;
  const uint64_t row0 = (uint64_t)blockIdx.x * ROWS;
  uint32_t* t = tile[tid];
  for (uint32_t c = 0; c < CPU_W; c++) t[c] = 0u;
  if (row0 + tid < n_events) {
    const uint32_t* p = events + (row0 + tid) * CPU_EV_WORDS;
    CpuEv e;
    uint32_t* w = reinterpret_cast<uint32_t*>(&e);
#pragma unroll
    for (uint32_t k = 0; k < CPU_EV_WORDS; k++) w[k] = __ldg(p + k);
    const uint32_t imm_b = (e.flags >> 1) & 1, imm_c = (e.flags >> 2) & 1, check = (e.flags >> 4) & 1;
    t[0] = e.shard;
    t[1] = e.clk & 0xFFFF;
    t[2] = (e.clk >> 16) & 0xFF;
    t[3] = check ? e.shard : 0u;
    t[4] = check ? e.clk : 0u;
    t[5] = e.pc;
    t[6] = e.next_pc;
    t[7] = e.next_next_pc;
    t[8] = e.opcode;
    t[9] = e.op_a;
    cpu_put_word(t + 10, e.op_b);
    cpu_put_word(t + 14, e.op_c);
    t[18] = e.flags & 1;
    t[19] = imm_b;
    t[20] = imm_c;
    t[21] = e.num_extra_cycles;
    t[22] = (e.flags >> 3) & 1;
    t[23] = check;
    t[24] = (e.flags >> 5) & 1;
    t[25] = (e.flags >> 6) & 1;
    cpu_put_word(t + 26, e.a);
    cpu_put_word(t + 30, e.hi);
    cpu_put_word(t + 34, e.a_prev_value);
    cpu_put_access(t + 38, e.a, e.shard, e.clk + 3, e.a_prev_shard, e.a_prev_clk);
    if (imm_b) cpu_put_word(t + 47, e.b);
    else cpu_put_access(t + 47, e.b, e.shard, e.clk + 2, e.b_prev_shard, e.b_prev_clk);
    if (imm_c) cpu_put_word(t + 56, e.c);
    else cpu_put_access(t + 56, e.c, e.shard, e.clk + 1, e.c_prev_shard, e.c_prev_clk);
    t[65] = 1;
    t[66] = (e.flags >> 7) & 1;
  } else {
    t[19] = 1;
    t[20] = 1;
    t[22] = 1;
  }
  __syncthreads();
  uint64_t base = row0 * CPU_W, end = rows * CPU_W;
  for (uint32_t idx = tid; idx < ROWS * CPU_W; idx += ROWS)
    if (base + idx < end) out[base + idx] = kb::to_monty(tile[idx / CPU_W][idx % CPU_W]);
}

template <class CHIP>
__global__ void __launch_bounds__(ROWS) alu_rows(const uint32_t* __restrict__ events, uint64_t n_events, uint64_t rows,
                                                 uint32_t* __restrict__ out) {
  constexpr uint32_t W = CHIP::W;
  __shared__ uint32_t ev[ROWS * 7];
  __shared__ uint32_t tile[ROWS][W + 1];
  const uint32_t tid = threadIdx.x;
  const uint64_t row0 = (uint64_t)blockIdx.x * ROWS;
  for (uint32_t idx = tid; idx < ROWS * 7; idx += ROWS) ev[idx] = row0 * 7 + idx < n_events * 7 ? events[row0 * 7 + idx] : 0u;
  for (uint32_t c = 0; c < W; c++) tile[tid][c] = 0u;
  __syncthreads();
  if (row0 + tid < n_events) {
    const uint32_t* e = ev + 7 * tid;
    AluEv a{e[0], e[1], e[2] & 0xFFu, e[3], e[4], e[5], e[6]};
    CHIP::fill(a, tile[tid]);
  } else {
    CHIP::pad(tile[tid]);
  }
  __syncthreads();
  uint64_t base = row0 * W, end = rows * W;
  for (uint32_t idx = tid; idx < ROWS * W; idx += ROWS)
    if (base + idx < end) out[base + idx] = kb::to_monty(tile[idx / W][idx % W]);
}

}  // namespace tg
