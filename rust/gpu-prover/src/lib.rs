//! `GpuProver<A>`: the reference's `MachineProver` (crates/stark/src/prover.rs:30-184) over libzkgpu's C ABI.
//!
//! NOT COMPILED in the build container (no Rust toolchain, Plonky3 not vendored).  The call sequence below is the one
//! `zkmips_b200/prover.py::GpuShardProver` executes on the GPU through the same ABI, where it is tested: proofs are
//! byte-identical (bincode image) to the CPU restatement of `CpuProver` and accepted by the verifier restated from
//! crates/stark/src/verifier.rs (tests/test_shard_prove.py).  Line references are to crates/stark/src/prover.rs.
//!
//! Concrete over `KoalaBearPoseidon2`: the FFI speaks Montgomery `u32` words of KoalaBear and Poseidon2-16 digests.
//! Field elements cross the boundary by pointer cast: `KoalaBear` is `#[repr(transparent)]` over its Montgomery
//! `u32` (the cast the reference itself makes, crates/core/machine/cpp/extern.cpp:12), an extension element is its 4
//! base coefficients in order (crates/stark/src/air/extension.rs:14-25).
use std::{
    cmp::Reverse,
    ffi::CString,
    ptr,
    sync::atomic::{AtomicUsize, Ordering},
};

use hashbrown::HashMap;
use p3_air::Air;
use p3_challenger::{CanObserve, DuplexChallenger, FieldChallenger};
use p3_commit::{BatchOpening, Pcs, PolynomialSpace};
use p3_field::{FieldAlgebra, FieldExtensionAlgebra, TwoAdicField};
use p3_fri::{CommitPhaseProofStep, FriProof, QueryProof};
use p3_koala_bear::KoalaBear;
use p3_matrix::{dense::RowMajorMatrix, Matrix};
use p3_maybe_rayon::prelude::*;
use p3_symmetric::Hash;
use p3_util::log2_strict_usize;
use zkm_gpu_sys as sys;
use zkm_stark::{
    air::{LookupScope, MachineAir},
    koala_bear_poseidon2::KoalaBearPoseidon2 as SC,
    septic_curve::SepticCurve,
    septic_digest::SepticDigest,
    septic_extension::SepticExtension,
    AirOpenedValues, Challenge, Challenger, ChipOpenedValues, Com, DebugConstraintBuilder, MachineProof, MachineProver,
    MachineProvingKey, MachineRecord, ShardCommitment, ShardMainData, ShardOpenedValues, ShardProof, StarkGenericConfig,
    StarkMachine, StarkProvingKey, StarkVerifyingKey, Val, ZKMCoreOpts,
};

type F = KoalaBear;
type EF = Challenge<SC>;
const D: usize = 4;

#[derive(Debug)]
pub struct GpuProverError(pub i32, pub String);
impl std::fmt::Display for GpuProverError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "libzkgpu status {}: {}", self.0, self.1)
    }
}
impl std::error::Error for GpuProverError {}

fn ck(rc: i32) -> Result<(), GpuProverError> {
    if rc == sys::ZK_OK {
        Ok(())
    } else {
        Err(GpuProverError(rc, sys::last_error()))
    }
}

#[inline]
fn words(v: &[F]) -> *const u32 {
    v.as_ptr().cast() // #[repr(transparent)] Montgomery u32
}
#[inline]
fn ext_words(e: &EF) -> [u32; D] {
    let s: &[F] = e.as_base_slice();
    unsafe { [*words(&s[0..1]), *words(&s[1..2]), *words(&s[2..3]), *words(&s[3..4])] }
}
#[inline]
fn felt(w: u32) -> F {
    unsafe { std::mem::transmute::<u32, F>(w) } // the library returns canonical Montgomery residues
}
#[inline]
fn ext(w: &[u32]) -> EF {
    EF::from_base_fn(|i| felt(w[i]))
}
fn digest(w: &[u32]) -> Com<SC> {
    Hash::from(core::array::from_fn::<F, 8, _>(|i| felt(w[i])))
}

/// One `zk_ctx` = one GPU + one stream.  Entry points on a context are serialised inside the library, so rayon workers
/// may share it; two contexts per GPU keep two shards in flight (`shard_batch_size`).
pub struct GpuCtx(*mut sys::ZkCtx);
unsafe impl Send for GpuCtx {}
unsafe impl Sync for GpuCtx {}
impl GpuCtx {
    pub fn new(device: i32) -> Result<Self, GpuProverError> {
        let mut p = ptr::null_mut();
        ck(unsafe { sys::zk_ctx_create(device, &mut p) })?;
        ck(unsafe { sys::zk_ctx_keep_traces(p, 1) })?; // LogUp reads the traces between the commits
        Ok(Self(p))
    }
}
impl Drop for GpuCtx {
    fn drop(&mut self) {
        unsafe { sys::zk_ctx_destroy(self.0) } // deferred by the library while prover data handles are alive
    }
}

/// `Mmcs::ProverData` on the device: LDE matrices + digest layers (+ retained traces).
pub struct GpuPdata(*mut sys::ZkPdata);
unsafe impl Send for GpuPdata {}
unsafe impl Sync for GpuPdata {}
impl Drop for GpuPdata {
    fn drop(&mut self) {
        unsafe { sys::zk_pdata_free(self.0) }
    }
}

/// `type DeviceProvingKey`: the host key plus the preprocessed round resident on every context (`pk_to_device`, :63).
pub struct GpuProvingKey {
    pub host: StarkProvingKey<SC>,
    pub data: Vec<GpuPdata>, // one per context
}
impl MachineProvingKey<SC> for GpuProvingKey {
    fn preprocessed_commit(&self) -> Com<SC> {
        self.host.commit.clone()
    }
    fn pc_start(&self) -> Val<SC> {
        self.host.pc_start
    }
    fn initial_global_cumulative_sum(&self) -> SepticDigest<Val<SC>> {
        self.host.initial_global_cumulative_sum
    }
    fn observe_into(&self, challenger: &mut Challenger<SC>) {
        self.host.observe_into(challenger)
    }
}

/// `type DeviceProverData`: the main round's handle and the context it lives on.
pub struct GpuMainData {
    pub pdata: GpuPdata,
    pub ctx: usize,
}

pub struct GpuProver<A> {
    machine: StarkMachine<SC, A>,
    ctxs: Vec<GpuCtx>,
    next: AtomicUsize,
}

impl<A> GpuProver<A> {
    fn log_blowup(&self) -> u32 {
        self.machine.config().pcs().fri_config().log_blowup as u32 // fork accessor, recursion/circuit/src/lib.rs:566
    }

    /// `Pcs::commit` of host matrices on natural domains (shift 1): :277, machine.rs:397.
    fn commit_host(&self, ctx: &GpuCtx, mats: &[&RowMajorMatrix<F>]) -> Result<(Com<SC>, GpuPdata), GpuProverError> {
        let ptrs: Vec<*const u32> = mats.iter().map(|m| words(&m.values)).collect();
        let heights: Vec<u64> = mats.iter().map(|m| m.height() as u64).collect();
        let widths: Vec<u32> = mats.iter().map(|m| m.width() as u32).collect();
        let one = [F::ONE];
        let shifts: Vec<u32> = mats.iter().map(|_| unsafe { *words(&one) }).collect();
        let mut root = [0u32; 8];
        let mut pd = ptr::null_mut();
        ck(unsafe {
            sys::zk_commit(ctx.0, mats.len() as u32, ptrs.as_ptr(), heights.as_ptr(), widths.as_ptr(), shifts.as_ptr(),
                           self.log_blowup(), root.as_mut_ptr(), &mut pd)
        })?;
        Ok((digest(&root), GpuPdata(pd)))
    }

    /// `Pcs::commit` of device-resident matrices: :403 (permutation traces), :497 (quotient chunks on shifted domains).
    fn commit_dev(&self, ctx: &GpuCtx, ptrs: &[u64], heights: &[u64], widths: &[u32], shifts: &[F])
                  -> Result<(Com<SC>, GpuPdata), GpuProverError> {
        let mut root = [0u32; 8];
        let mut pd = ptr::null_mut();
        ck(unsafe {
            sys::zk_commit_dev(ctx.0, ptrs.len() as u32, ptrs.as_ptr(), heights.as_ptr(), widths.as_ptr(), words(shifts),
                               self.log_blowup(), root.as_mut_ptr(), &mut pd)
        })?;
        Ok((digest(&root), GpuPdata(pd)))
    }
}

/// `DuplexChallenger{sponge_state, input_buffer, output_buffer}` <-> the library's 34-word image.
fn challenger_to_ffi(ch: &Challenger<SC>) -> sys::ZkChallenger {
    let mut z = sys::ZkChallenger::default();
    for (i, v) in ch.sponge_state.iter().enumerate() {
        z.state[i] = unsafe { *words(core::slice::from_ref(v)) };
    }
    for (i, v) in ch.input_buffer.iter().enumerate() {
        z.inp[i] = unsafe { *words(core::slice::from_ref(v)) };
    }
    z.n_in = ch.input_buffer.len() as u32;
    for (i, v) in ch.output_buffer.iter().enumerate() {
        z.out[i] = unsafe { *words(core::slice::from_ref(v)) };
    }
    z.n_out = ch.output_buffer.len() as u32;
    z
}
fn challenger_from_ffi(z: &sys::ZkChallenger, ch: &mut Challenger<SC>) {
    for i in 0..16 {
        ch.sponge_state[i] = felt(z.state[i]);
    }
    ch.input_buffer = z.inp[..z.n_in as usize].iter().map(|&w| felt(w)).collect();
    ch.output_buffer = z.out[..z.n_out as usize].iter().map(|&w| felt(w)).collect();
}

impl<A> MachineProver<SC, A> for GpuProver<A>
where
    A: MachineAir<F> + for<'a> Air<DebugConstraintBuilder<'a, F, EF>> + 'static + Send + Sync,
    A::Record: MachineRecord<Config = ZKMCoreOpts>,
{
    type DeviceMatrix = RowMajorMatrix<F>; // host copy (last row feeds the global sum); the device copy is retained in the pdata
    type DeviceProverData = GpuMainData;
    type DeviceProvingKey = GpuProvingKey;
    type Error = GpuProverError;

    fn new(machine: StarkMachine<SC, A>) -> Self {
        // one context per (GPU, in-flight slot): ZKGPU_DEVICES="0,1,.." (default "0"), two slots per GPU
        let devs = std::env::var("ZKGPU_DEVICES").unwrap_or_else(|_| "0".into());
        let mut ctxs = Vec::new();
        for d in devs.split(',').filter_map(|s| s.trim().parse::<i32>().ok()) {
            for _slot in 0..2 {
                ctxs.push(GpuCtx::new(d).expect("zk_ctx_create failed: libzkgpu has no CPU fallback"));
            }
        }
        Self { machine, ctxs, next: AtomicUsize::new(0) }
    }

    fn machine(&self) -> &StarkMachine<SC, A> {
        &self.machine
    }

    fn setup(&self, program: &A::Program) -> (Self::DeviceProvingKey, StarkVerifyingKey<SC>) {
        // machine.rs:330-440 generates and orders the preprocessed traces and commits them on the CPU; the device copy
        // is made by pk_to_device.  (A GPU-side setup would replace the `pcs.commit` at machine.rs:396-397 by
        // commit_host and export the prover data with zk_pdata_copy_lde / zk_pdata_copy_layer for `Serialize`.)
        let (pk, vk) = self.machine.setup(program);
        (self.pk_to_device(&pk), vk)
    }

    fn pk_from_vk(&self, program: &A::Program, vk: &StarkVerifyingKey<SC>) -> Self::DeviceProvingKey {
        self.pk_to_device(&self.machine.setup_core(program, vk.initial_global_cumulative_sum).0)
    }

    fn pk_to_device(&self, pk: &StarkProvingKey<SC>) -> Self::DeviceProvingKey {
        // same trace order as machine.rs:383-384, so matrix i of the device round is pk.traces[i]; the root must be pk.commit
        let mats: Vec<&RowMajorMatrix<F>> = pk.traces.iter().collect();
        let data = self
            .ctxs
            .iter()
            .map(|ctx| {
                let (root, pd) = self.commit_host(ctx, &mats).expect("preprocessed commit failed");
                assert!(root == pk.commit, "device commitment of the preprocessed traces differs from pk.commit");
                pd
            })
            .collect();
        GpuProvingKey { host: pk.clone(), data }
    }

    fn pk_to_host(&self, pk: &Self::DeviceProvingKey) -> StarkProvingKey<SC> {
        pk.host.clone()
    }

    /// :258-292
    fn commit(&self, record: &A::Record, mut named_traces: Vec<(String, RowMajorMatrix<F>)>)
              -> ShardMainData<SC, Self::DeviceMatrix, Self::DeviceProverData> {
        named_traces.sort_by_key(|(name, trace)| (Reverse(trace.height()), name.clone())); // :264
        let slot = self.next.fetch_add(1, Ordering::Relaxed) % self.ctxs.len(); // shard i -> context i mod n (prove.rs:480-526)
        let mats: Vec<&RowMajorMatrix<F>> = named_traces.iter().map(|(_, t)| t).collect();
        let (main_commit, pdata) = self.commit_host(&self.ctxs[slot], &mats).expect("main commit failed"); // :277
        let chip_ordering = named_traces.iter().enumerate().map(|(i, (name, _))| (name.to_owned(), i)).collect();
        let traces = named_traces.into_iter().map(|(_, t)| t).collect();
        ShardMainData {
            traces,
            main_commit,
            main_data: GpuMainData { pdata, ctx: slot },
            chip_ordering,
            public_values: record.public_values(),
        }
    }

    /// :298-653
    fn open(&self, pk: &Self::DeviceProvingKey, data: ShardMainData<SC, Self::DeviceMatrix, Self::DeviceProverData>,
            challenger: &mut Challenger<SC>) -> Result<ShardProof<SC>, Self::Error> {
        let chips = self.machine.shard_chips_ordered(&data.chip_ordering).collect::<Vec<_>>();
        let traces = &data.traces;
        let slot = data.main_data.ctx;
        let ctx = &self.ctxs[slot];
        let main_pd = &data.main_data.pdata;
        let prep_pd = &pk.data[slot];
        let lb = self.log_blowup();
        let log_degrees: Vec<usize> = traces.iter().map(|t| log2_strict_usize(t.height())).collect();
        let lqds: Vec<usize> = chips.iter().map(|c| c.log_quotient_degree()).collect();
        let pcs = self.machine.config().pcs();
        let trace_domains: Vec<_> = traces.iter().map(|t| pcs.natural_domain_for_degree(t.height())).collect();
        let air_ids: Vec<i32> = chips
            .iter()
            .map(|c| unsafe { sys::zk_air_find(CString::new(c.name()).unwrap().as_ptr()) })
            .collect();
        assert!(air_ids.iter().all(|&id| id >= 0), "a chip of this shard is not compiled into libzkgpu");

        challenger.observe_slice(&data.public_values[0..self.num_pv_elts()]); // :322
        challenger.observe(data.main_commit.clone()); // :323
        let perm_challenges: Vec<EF> = (0..2).map(|_| challenger.sample_ext_element()).collect(); // :326-329
        let chal_words: Vec<u32> = perm_challenges.iter().flat_map(ext_words).collect();

        // permutation traces, on the device from the retained traces (:341-364); width 0 without local lookups
        let mut perm_ptrs = Vec::new();
        let mut perm_widths = Vec::new();
        let mut local_sums: Vec<EF> = Vec::new();
        let mut global_sums: Vec<SepticDigest<F>> = Vec::new();
        for (i, chip) in chips.iter().enumerate() {
            let mut desc = sys::ZkAirDesc::default();
            ck(unsafe { sys::zk_air_info(air_ids[i], &mut desc) })?;
            let mut dptr: u64 = 0;
            let mut lcs = [0u32; 4];
            if desc.num_lookups > 0 {
                let prep_trace = pk.host.chip_ordering.get(&chip.name())
                    .map(|&k| unsafe { sys::zk_pdata_trace(prep_pd.0, k as u32) }).unwrap_or(0);
                ck(unsafe {
                    sys::zk_permutation_trace(ctx.0, air_ids[i], prep_trace, sys::zk_pdata_trace(main_pd.0, i as u32),
                                              traces[i].height() as u64, chal_words.as_ptr(), &mut dptr, lcs.as_mut_ptr())
                })?;
            }
            perm_ptrs.push(dptr);
            perm_widths.push((D as u32) * desc.perm_width); // flatten_to_base, :393
            local_sums.push(ext(&lcs));
            global_sums.push(if chip.commit_scope() == LookupScope::Local {
                SepticDigest::<F>::zero()
            } else {
                let t = &traces[i]; // :353-361: last 14 words of the main trace
                let last = &t.values[t.values.len() - 14..];
                SepticDigest(SepticCurve {
                    x: SepticExtension::<F>::from_base_fn(|k| last[k]),
                    y: SepticExtension::<F>::from_base_fn(|k| last[k + 7]),
                })
            });
        }
        let heights: Vec<u64> = traces.iter().map(|t| t.height() as u64).collect();
        let ones = vec![F::ONE; chips.len()];
        let (permutation_commit, perm_pd) = self.commit_dev(ctx, &perm_ptrs, &heights, &perm_widths, &ones)?; // :401-403
        for &p in perm_ptrs.iter().filter(|&&p| p != 0) {
            ck(unsafe { sys::zk_dev_free(ctx.0, p) })?;
        }
        challenger.observe(permutation_commit.clone()); // :406
        for (l, g) in local_sums.iter().zip(global_sums.iter()) {
            challenger.observe_slice(l.as_base_slice()); // :407-413
            challenger.observe_slice(&g.0.x.0);
            challenger.observe_slice(&g.0.y.0);
        }
        let alpha: EF = challenger.sample_ext_element(); // :426

        // quotient values (:429-475) written as split chunk matrices (:477-488), committed on the shifted domains
        let mut chunk_ptrs = Vec::new();
        let mut chunk_heights = Vec::new();
        let mut chunk_shifts: Vec<F> = Vec::new();
        let mut chunk_bufs = Vec::new();
        for (i, chip) in chips.iter().enumerate() {
            let (n, lqd) = (log_degrees[i], lqds[i]);
            let g = global_sums[i];
            let gcs: Vec<u32> = g.0.x.0.iter().chain(g.0.y.0.iter()).map(|v| unsafe { *words(core::slice::from_ref(v)) }).collect();
            let (prep, prep_idx) = match pk.host.chip_ordering.get(&chip.name()) {
                Some(&k) => (prep_pd.0 as *const sys::ZkPdata, k as u32),
                None => (ptr::null(), 0),
            };
            let mut out: u64 = 0;
            ck(unsafe {
                sys::zk_quotient(ctx.0, air_ids[i], prep, prep_idx, main_pd.0, i as u32, perm_pd.0, i as u32, n as u32,
                                 lqd as u32, ext_words(&alpha).as_ptr(), chal_words.as_ptr(), words(&data.public_values),
                                 data.public_values.len() as u32, ext_words(&local_sums[i]).as_ptr(), gcs.as_ptr(), &mut out)
            })?;
            let quotient_domain = trace_domains[i].create_disjoint_domain(1 << (n + lqd));
            for (c, qd) in quotient_domain.split_domains(1 << lqd).into_iter().enumerate() {
                chunk_ptrs.push(out + (c as u64) * (1u64 << n) * 16);
                chunk_heights.push(1u64 << n);
                chunk_shifts.push(qd.shift); // GENERATOR * g_{n+lqd}^c
            }
            chunk_bufs.push(out);
        }
        let chunk_widths = vec![D as u32; chunk_ptrs.len()];
        let (quotient_commit, quot_pd) = self.commit_dev(ctx, &chunk_ptrs, &chunk_heights, &chunk_widths, &chunk_shifts)?; // :496-497
        for p in chunk_bufs {
            ck(unsafe { sys::zk_dev_free(ctx.0, p) })?;
        }
        challenger.observe(quotient_commit.clone()); // :498
        let zeta: EF = challenger.sample_ext_element(); // :501

        // opening points (:503-544), rounds [preprocessed (all pk traces), main, permutation, quotient] (:548-553)
        let mut n_points: Vec<u32> = Vec::new();
        let mut points: Vec<u32> = Vec::new();
        let mut push = |pts: &[EF]| {
            n_points.push(pts.len() as u32);
            for p in pts {
                points.extend_from_slice(&ext_words(p));
            }
        };
        for (trace, local_only) in pk.host.traces.iter().zip(pk.host.local_only.iter()) {
            let domain = pcs.natural_domain_for_degree(trace.height());
            if *local_only { push(&[zeta]) } else { push(&[zeta, domain.next_point(zeta).unwrap()]) }
        }
        for (domain, chip) in trace_domains.iter().zip(chips.iter()) {
            if chip.local_only() { push(&[zeta]) } else { push(&[zeta, domain.next_point(zeta).unwrap()]) }
        }
        for domain in trace_domains.iter() {
            push(&[zeta, domain.next_point(zeta).unwrap()]);
        }
        for _ in 0..chunk_ptrs.len() {
            push(&[zeta]);
        }
        let rounds: [*const sys::ZkPdata; 4] = [prep_pd.0, main_pd.0, perm_pd.0, quot_pd.0];
        let fri = pcs.fri_config();
        let n_words = unsafe { sys::zk_pcs_proof_words(4, rounds.as_ptr(), n_points.as_ptr(), lb, fri.num_queries as u32) };
        let mut flat = vec![0u32; n_words as usize];
        let mut z = challenger_to_ffi(challenger);
        ck(unsafe {
            sys::zk_pcs_open(ctx.0, 4, rounds.as_ptr(), n_points.as_ptr(), points.as_ptr(), lb, fri.num_queries as u32,
                             fri.proof_of_work_bits as u32, &mut z, -1, flat.as_mut_ptr(), n_words)
        })?; // :546-556
        challenger_from_ffi(&z, challenger);

        // ---- repackaging (:558-652): flat proof (layout in include/zkgpu.h) -> OpenedValues + FriProof
        let widths_of = |pd: &GpuPdata| -> Vec<usize> {
            (0..unsafe { sys::zk_pdata_num_matrices(pd.0) }).map(|i| unsafe { sys::zk_pdata_width(pd.0, i) } as usize).collect()
        };
        let round_widths = [widths_of(prep_pd), widths_of(main_pd), widths_of(&perm_pd), widths_of(&quot_pd)];
        let round_logmax: Vec<usize> = [prep_pd, main_pd, &perm_pd, &quot_pd].iter()
            .map(|pd| unsafe { sys::zk_pdata_log_max_height(pd.0) } as usize).collect();
        let mut off = 0usize;
        let mut k = 0usize;
        let mut opened: Vec<Vec<Vec<Vec<EF>>>> = Vec::new(); // round -> matrix -> point -> width
        for ws in round_widths.iter() {
            let mut rnd = Vec::new();
            for &w in ws {
                let mut mat = Vec::new();
                for _ in 0..n_points[k] {
                    mat.push((0..w).map(|j| ext(&flat[off + 4 * j..off + 4 * j + 4])).collect::<Vec<EF>>());
                    off += 4 * w;
                }
                rnd.push(mat);
                k += 1;
            }
            opened.push(rnd);
        }
        let log_max = *round_logmax.iter().max().unwrap();
        let n_layers = log_max - lb as usize;
        let commit_phase_commits: Vec<Com<SC>> = (0..n_layers).map(|i| digest(&flat[off + 8 * i..off + 8 * i + 8])).collect();
        off += 8 * n_layers;
        let final_poly = ext(&flat[off..off + 4]);
        off += 4;
        let pow_witness = felt(flat[off]);
        off += 1;
        let path = |flat: &[u32], off: &mut usize, depth: usize| -> Vec<[F; 8]> {
            (0..depth).map(|_| { let d = core::array::from_fn(|i| felt(flat[*off + i])); *off += 8; d }).collect()
        };
        let mut query_proofs = Vec::with_capacity(fri.num_queries);
        for _ in 0..fri.num_queries {
            let mut input_proof = Vec::with_capacity(4);
            for (ws, &lm) in round_widths.iter().zip(round_logmax.iter()) {
                let mut rows = Vec::with_capacity(ws.len());
                for &w in ws {
                    rows.push(flat[off..off + w].iter().map(|&x| felt(x)).collect::<Vec<F>>());
                    off += w;
                }
                input_proof.push(BatchOpening { opened_values: rows, opening_proof: path(&flat, &mut off, lm) });
            }
            let mut steps = Vec::with_capacity(n_layers);
            for i in 0..n_layers {
                let sibling_value = ext(&flat[off..off + 4]);
                off += 4;
                steps.push(CommitPhaseProofStep { sibling_value, opening_proof: path(&flat, &mut off, log_max - i - 1) });
            }
            query_proofs.push(QueryProof { input_proof, commit_phase_openings: steps });
        }
        debug_assert_eq!(off, flat.len());
        let opening_proof = FriProof { commit_phase_commits, query_proofs, final_poly, pow_witness };

        let air_values = |op: &Vec<Vec<EF>>| -> AirOpenedValues<EF> {
            if op.len() == 2 {
                AirOpenedValues { local: op[0].clone(), next: op[1].clone() }
            } else {
                AirOpenedValues { local: op[0].clone(), next: vec![EF::ZERO; op[0].len()] } // local_only, :566-570
            }
        };
        let mut qi = 0usize;
        let mut chip_values = Vec::with_capacity(chips.len());
        for (i, chip) in chips.iter().enumerate() {
            let preprocessed = pk.host.chip_ordering.get(&chip.name())
                .map(|&k| air_values(&opened[0][k]))
                .unwrap_or(AirOpenedValues { local: vec![], next: vec![] });
            let nch = 1usize << lqds[i];
            let quotient = (0..nch).map(|j| opened[3][qi + j][0].clone()).collect::<Vec<_>>();
            qi += nch;
            chip_values.push(ChipOpenedValues {
                preprocessed,
                main: air_values(&opened[1][i]),
                permutation: air_values(&opened[2][i]),
                quotient,
                global_cumulative_sum: global_sums[i],
                local_cumulative_sum: local_sums[i],
                log_degree: log_degrees[i],
            });
        }
        Ok(ShardProof::<SC> {
            commitment: ShardCommitment { main_commit: data.main_commit.clone(), permutation_commit, quotient_commit },
            opened_values: ShardOpenedValues { chips: chip_values },
            opening_proof,
            chip_ordering: data.chip_ordering,
            public_values: data.public_values,
        })
    }

    /// :660-693.  Shards are independent (challenger cloned per shard): rayon workers land on contexts round-robin.
    fn prove(&self, pk: &Self::DeviceProvingKey, mut records: Vec<A::Record>, challenger: &mut Challenger<SC>,
             opts: <A::Record as MachineRecord>::Config) -> Result<MachineProof<SC>, Self::Error> {
        self.machine.generate_dependencies(&mut records, &opts, None).map_err(|_| GpuProverError(-3, "dependencies".into()))?;
        pk.observe_into(challenger);
        let shard_proofs = records
            .into_par_iter()
            .map(|record| {
                let named_traces = self.generate_traces(&record).map_err(|_| GpuProverError(-3, "generate_traces".into()))?;
                let shard_data = self.commit(&record, named_traces);
                self.open(pk, shard_data, &mut challenger.clone())
            })
            .collect::<Result<Vec<_>, _>>()?;
        Ok(MachineProof { shard_proofs })
    }
}

/// Keeps `HashMap` in the public interface the same type the reference uses (hashbrown).
pub type ChipOrdering = HashMap<String, usize>;
