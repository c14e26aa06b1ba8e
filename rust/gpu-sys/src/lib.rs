//! FFI declarations of `libzkgpu.so` -- one `pub fn` per function of `include/zkgpu.h`, same order.
//!
//! Idiom of the reference's own native bindings: `extern "C-unwind"` over `u32` Montgomery words
//! (crates/core/machine/src/sys.rs:14-42; `KoalaBear` is `#[repr(transparent)]` over its Montgomery `u32`,
//! crates/core/machine/cpp/extern.cpp:12).  Every function returns 0 or a negative `zk_status`; nothing unwinds.
//! NOT COMPILED in the build container (no Rust toolchain there); tests/test_abi.py checks it against the header.
#![allow(non_camel_case_types)]

#[repr(C)]
pub struct ZkCtx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct ZkPdata {
    _p: [u8; 0],
}
/// device address
pub type ZkDptr = u64;

pub const ZK_OK: i32 = 0;
pub const ZK_ERR_ARG: i32 = -1;
pub const ZK_ERR_CUDA: i32 = -2;
pub const ZK_ERR_STATE: i32 = -3;
pub const ZK_ERR_VERIFY: i32 = -4;

/// `zk_alu_event`: the layout of `zkm_core_executor::events::AluEvent` (#[repr(C)], events/instr.rs:10-26), so a
/// `&[AluEvent]` can be handed over as `*const ZkAluEvent` without a copy.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ZkAluEvent {
    pub pc: u32,
    pub next_pc: u32,
    pub opcode: u8,
    pub pad_: [u8; 3],
    pub hi: u32,
    pub a: u32,
    pub b: u32,
    pub c: u32,
}
/// `zk_cpu_event`: the CpuEvent's scalars, its instruction and the previous (shard, clk) of its three register accesses, packed
/// by the host from `CpuEvent` + `Instruction` (crates/core/executor/src/events/cpu.rs) -- 22 words per executed instruction.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ZkCpuEvent {
    pub pc: u32, pub next_pc: u32, pub next_next_pc: u32, pub clk: u32, pub shard: u32, pub opcode: u32, pub op_a: u32,
    pub op_b: u32, pub op_c: u32, pub flags: u32, pub num_extra_cycles: u32,
    pub a: u32, pub b: u32, pub c: u32, pub hi: u32,
    pub a_prev_value: u32, pub a_prev_shard: u32, pub a_prev_clk: u32, pub b_prev_shard: u32, pub b_prev_clk: u32,
    pub c_prev_shard: u32, pub c_prev_clk: u32,
}
pub const ZK_CHIP_ADD_SUB: i32 = 0;
pub const ZK_CHIP_BITWISE: i32 = 1;
pub const ZK_CHIP_LT: i32 = 2;
pub const ZK_CHIP_SHIFT_LEFT: i32 = 3;
pub const ZK_CHIP_SHIFT_RIGHT: i32 = 4;
pub const ZK_CHIP_CLO_CLZ: i32 = 5;

/// 34-word image of a Plonky3 `DuplexChallenger<KoalaBear, Perm, 16, 8>`: `sponge_state`, `input_buffer` (+ length),
/// `output_buffer` (+ length; samples pop from the END).
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct ZkChallenger {
    pub state: [u32; 16],
    pub inp: [u32; 8],
    pub n_in: u32,
    pub out: [u32; 8],
    pub n_out: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct ZkAirDesc {
    pub main_width: u32,
    pub prep_width: u32,
    pub perm_width: u32,
    pub num_public_values: u32,
    pub num_challenges: u32,
    pub num_constraints: u32,
    pub max_degree: u32,
    pub num_kernels: u32,
    pub num_lookups: u32,
}

#[link(name = "zkgpu")]
extern "C-unwind" {
    pub fn zk_ctx_create(device: i32, out: *mut *mut ZkCtx) -> i32;
    pub fn zk_ctx_create_on_stream(device: i32, cuda_stream: *mut core::ffi::c_void, out: *mut *mut ZkCtx) -> i32;
    pub fn zk_ctx_destroy(ctx: *mut ZkCtx);
    pub fn zk_ctx_sync(ctx: *mut ZkCtx) -> i32;
    pub fn zk_last_error() -> *const core::ffi::c_char;
    pub fn zk_build_info() -> *const core::ffi::c_char;
    pub fn zk_prof_enable(ctx: *mut ZkCtx, enable: i32) -> i32;
    pub fn zk_prof_reset(ctx: *mut ZkCtx) -> i32;
    pub fn zk_prof_count(ctx: *mut ZkCtx) -> i32;
    pub fn zk_prof_get(ctx: *mut ZkCtx, i: i32, name: *mut core::ffi::c_char, name_cap: i32, ms: *mut f32, launches: *mut u64) -> i32;
    pub fn zk_prof_start(ctx: *mut ZkCtx, i: i32, ms_after_first: *mut f32) -> i32;
    pub fn zk_launch_count(ctx: *mut ZkCtx) -> u64;
    pub fn zk_ntt_tma_passes() -> u64;
    pub fn zk_dev_alloc(ctx: *mut ZkCtx, bytes: u64, out: *mut ZkDptr) -> i32;
    pub fn zk_dev_free(ctx: *mut ZkCtx, p: ZkDptr) -> i32;
    pub fn zk_h2d(ctx: *mut ZkCtx, dst: ZkDptr, src_host: *const core::ffi::c_void, bytes: u64) -> i32;
    pub fn zk_d2h(ctx: *mut ZkCtx, dst_host: *mut core::ffi::c_void, src: ZkDptr, bytes: u64) -> i32;
    pub fn zk_poseidon2_permute(ctx: *mut ZkCtx, states_host: *mut u32, n: u64) -> i32;
    pub fn zk_hash_rows(ctx: *mut ZkCtx, mat_host: *const u32, h: u64, w: u32, digests_host: *mut u32) -> i32;
    pub fn zk_compress_layer(ctx: *mut ZkCtx, prev_host: *const u32, n_out: u64, out_host: *mut u32) -> i32;
    pub fn zk_dft_batch(ctx: *mut ZkCtx, in_host: *const u32, h: u64, w: u32, out_host: *mut u32) -> i32;
    pub fn zk_coset_lde(ctx: *mut ZkCtx, in_host: *const u32, h: u64, w: u32, log_blowup: u32, shift: u32, out_host: *mut u32) -> i32;
    pub fn zk_coset_lde_dev(ctx: *mut ZkCtx, in_: ZkDptr, h: u64, w: u32, log_blowup: u32, shift: u32, out: ZkDptr) -> i32;
    pub fn zk_commit(ctx: *mut ZkCtx, n_mats: u32, mats_host: *const *const u32, heights: *const u64, widths: *const u32, domain_shifts: *const u32, log_blowup: u32, root: *mut u32, out: *mut *mut ZkPdata) -> i32;
    pub fn zk_commit_dev(ctx: *mut ZkCtx, n_mats: u32, mats_dev: *const ZkDptr, heights: *const u64, widths: *const u32, domain_shifts: *const u32, log_blowup: u32, root: *mut u32, out: *mut *mut ZkPdata) -> i32;
    pub fn zk_mmcs_commit(ctx: *mut ZkCtx, n_mats: u32, mats_host: *const *const u32, heights: *const u64, widths: *const u32, root: *mut u32, out: *mut *mut ZkPdata) -> i32;
    pub fn zk_mmcs_commit_dev(ctx: *mut ZkCtx, n_mats: u32, mats_dev: *const ZkDptr, heights: *const u64, widths: *const u32, root: *mut u32, out: *mut *mut ZkPdata) -> i32;
    pub fn zk_ctx_keep_traces(ctx: *mut ZkCtx, enable: i32) -> i32;
    pub fn zk_ctx_set_upload_helper(ctx: *mut ZkCtx, device: i32) -> i32;
    pub fn zk_pdata_free(pd: *mut ZkPdata);
    pub fn zk_pdata_num_matrices(pd: *const ZkPdata) -> u32;
    pub fn zk_pdata_height(pd: *const ZkPdata, i: u32) -> u64;
    pub fn zk_pdata_width(pd: *const ZkPdata, i: u32) -> u32;
    pub fn zk_pdata_log_max_height(pd: *const ZkPdata) -> u32;
    pub fn zk_pdata_root(pd: *const ZkPdata, root: *mut u32) -> i32;
    pub fn zk_pdata_lde(pd: *const ZkPdata, i: u32) -> ZkDptr;
    pub fn zk_pdata_pitch(pd: *const ZkPdata, i: u32) -> u32;
    pub fn zk_pdata_trace(pd: *const ZkPdata, i: u32) -> ZkDptr;
    pub fn zk_pdata_copy_lde(pd: *const ZkPdata, i: u32, out_host: *mut u32) -> i32;
    pub fn zk_pdata_copy_layer(pd: *const ZkPdata, layer: u32, out_host: *mut u32) -> i32;
    pub fn zk_pdata_import(ctx: *mut ZkCtx, n_mats: u32, ldes_host: *const *const u32, heights: *const u64, widths: *const u32, layers_host: *const *const u32, n_layers: u32, traces_host: *const *const u32, log_blowup: u32, out: *mut *mut ZkPdata) -> i32;
    pub fn zk_pdata_open_batch(pd: *const ZkPdata, n_idx: u32, indices: *const u64, opened_host: *mut u32, proofs_host: *mut u32) -> i32;
    pub fn zk_air_count() -> i32;
    pub fn zk_air_name(id: i32) -> *const core::ffi::c_char;
    pub fn zk_air_find(name: *const core::ffi::c_char) -> i32;
    pub fn zk_air_info(id: i32, out: *mut ZkAirDesc) -> i32;
    pub fn zk_quotient(ctx: *mut ZkCtx, air_id: i32, prep: *const ZkPdata, prep_idx: u32, main_data: *const ZkPdata, main_idx: u32, perm: *const ZkPdata, perm_idx: u32, log_degree: u32, log_quotient_degree: u32, alpha: *const u32, perm_challenges: *const u32, public_values: *const u32, n_public_values: u32, local_cumsum: *const u32, global_cumsum: *const u32, out_chunks: *mut ZkDptr) -> i32;
    pub fn zk_permutation_trace(ctx: *mut ZkCtx, air_id: i32, prep_trace: ZkDptr, main_trace: ZkDptr, height: u64, perm_challenges: *const u32, out_trace: *mut ZkDptr, local_cumsum: *mut u32) -> i32;
    pub fn zk_tracegen_alu_width(chip: i32) -> u32;
    pub fn zk_tracegen_alu(ctx: *mut ZkCtx, chip: i32, events_host: *const ZkAluEvent, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_alu_dev(ctx: *mut ZkCtx, chip: i32, events_dev: ZkDptr, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_cpu_width() -> u32;
    pub fn zk_tracegen_cpu(ctx: *mut ZkCtx, events_host: *const ZkCpuEvent, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_cpu_dev(ctx: *mut ZkCtx, events_dev: ZkDptr, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_wide_width(sbox_state: i32) -> u32;
    pub fn zk_tracegen_poseidon2_wide(ctx: *mut ZkCtx, inputs_host: *const u32, n_events: u64, rows: u64, sbox_state: i32, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_wide_dev(ctx: *mut ZkCtx, inputs_dev: ZkDptr, n_events: u64, rows: u64, sbox_state: i32, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_wide_prep(ctx: *mut ZkCtx, instrs_host: *const u32, n: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_skinny_width() -> u32;
    pub fn zk_tracegen_poseidon2_skinny(ctx: *mut ZkCtx, inputs_host: *const u32, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_skinny_dev(ctx: *mut ZkCtx, inputs_dev: ZkDptr, n_events: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_tracegen_poseidon2_skinny_prep(ctx: *mut ZkCtx, instrs_host: *const u32, n: u64, rows: u64, out_trace: *mut ZkDptr) -> i32;
    pub fn zk_challenger_init(ch: *mut ZkChallenger) -> i32;
    pub fn zk_challenger_observe(ctx: *mut ZkCtx, ch: *mut ZkChallenger, vals: *const u32, n: u32) -> i32;
    pub fn zk_challenger_sample_ext(ctx: *mut ZkCtx, ch: *mut ZkChallenger, n_ext: u32, out: *mut u32) -> i32;
    pub fn zk_challenger_sample_bits(ctx: *mut ZkCtx, ch: *mut ZkChallenger, bits: u32, n: u32, out: *mut u64) -> i32;
    pub fn zk_challenger_grind(ctx: *mut ZkCtx, ch: *mut ZkChallenger, bits: u32, witness: *mut u32) -> i32;
    pub fn zk_pcs_proof_words(n_rounds: u32, rounds: *const *const ZkPdata, n_points: *const u32, log_blowup: u32, num_queries: u32) -> u64;
    pub fn zk_pcs_open(ctx: *mut ZkCtx, n_rounds: u32, rounds: *const *const ZkPdata, n_points: *const u32, points: *const u32, log_blowup: u32, num_queries: u32, pow_bits: u32, ch: *mut ZkChallenger, inject_witness: i64, proof_host: *mut u32, proof_cap: u64) -> i32;
}

/// text of the calling thread's last error
pub fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(zk_last_error()).to_string_lossy().into_owned() }
}
