"""Host-side mirror of the reference's per-shard prover, `MachineProver::{setup, commit, open, prove}`
(crates/stark/src/prover.rs:30-184; CpuProver::commit :258-292, ::open :298-653; StarkMachine::setup
crates/stark/src/machine.rs:330-440), driving libzkgpu through its C ABI.

In Ziren this orchestration is Rust (`CpuProver`); `rust/gpu-prover` shows the `GpuProver` that makes the same calls
through the FFI crate.  This Python mirror exists so that the whole commit -> permutation -> quotient -> open flow can
be exercised, verified and timed from the tests and bench.py; the only arithmetic it does itself is O(1) scalar work
per chip (zeta * g, domain shifts), exactly what `Domain::next_point` / `split_domains` do on the Rust side.
Transcript operations run on the device challenger, everything else in the kernels.

Transcript (prover.rs line numbers on every step below):
  [pk.observe_into: commit, pc_start, initial_global_cumulative_sum x/y, 0   (machine.rs:79-86; once per proof)]
  observe public_values[0 .. num_pv_elts]; observe main commit; sample 2 permutation challenges;
  permutation trace for EVERY chip (width 0 without local lookups), commit; observe the commit, then per chip
  local_sum[4], global_sum.x[7], global_sum.y[7] (global sum = last 14 main columns of the last row for Global-scope
  chips, zero otherwise); sample alpha; quotient; observe quotient commit; sample zeta;
  Pcs::open over [preprocessed (ALL pk traces), main, permutation, quotient] with points [zeta, zeta*g] -- [zeta] only
  for `local_only` chips' preprocessed / main traces and for the quotient chunks."""
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import proof as pf
from .native import Challenger, pcs_open

P = 0x7F000001
MONTY_ONE = 0x01FFFFFE
R = 1 << 32


def monty(v):
    return (int(v) % P) * R % P


def two_adic_generator(bits):
    g = pow(3, 127, P)
    for _ in range(bits, 24):
        g = g * g % P
    return g


@dataclass
class Chip:
    """One chip of a shard: traces are (height, width) uint32 Montgomery matrices.  `local_only` and `commit_scope`
    are properties of the chip's AIR (MachineAir::local_only / commit_scope) and must agree with the compiled AIR."""
    name: str
    air: str
    main: Optional[np.ndarray]           # None: the main trace is generated on the device from `events` (see below)
    preprocessed: Optional[np.ndarray] = None
    local_only: bool = False
    commit_scope: str = "local"          # "local" | "global" (LookupScope)
    log_quotient_degree: int = 1
    # device trace generation (MachineAir::generate_trace on the GPU, csrc/tracegen.cuh): the event records of the
    # shard's ExecutionRecord for this chip, the filler's name and the padded height (MachineAir::num_rows)
    events: Optional[np.ndarray] = None
    tracegen: Optional[str] = None       # "AddSub" | "Bitwise" | "Lt" | "ShiftLeft" | "ShiftRight" | "CloClz" | "Poseidon2WideDeg3" | "Poseidon2WideDeg9" | "Poseidon2SkinnyDeg9"
    rows: Optional[int] = None

    @property
    def height(self):
        return int(self.main.shape[0]) if self.main is not None else int(self.rows)

    @property
    def log_degree(self):
        return self.height.bit_length() - 1

    def global_cumulative_sum(self):
        """prover.rs:353-361: zero for Local-scope chips, else the last 14 words of the main trace"""
        if self.commit_scope == "local":
            return np.zeros(14, np.uint32)
        return np.ascontiguousarray(self.main.reshape(-1)[-14:], dtype=np.uint32)


@dataclass
class ProvingKey:
    """StarkProvingKey (machine.rs:56-76): commitment to the preprocessed traces of the MACHINE (every chip that has
    one, ordered by (-height, name), machine.rs:383-384), kept on the device (`pk_to_device`, prover.rs:63)."""
    commit: Optional[np.ndarray]
    pc_start: int                                   # Montgomery word
    initial_global_cumulative_sum: np.ndarray       # 14 words
    traces: List[np.ndarray]
    data: object                                    # PData of the preprocessed round (None: machine without one)
    chip_ordering: Dict[str, int]
    local_only: List[bool]
    constraints_map: Dict[str, int]

    def observe_into(self, ch: Challenger):
        """machine.rs:79-86 (the verifying key observes the same words, :108-115)"""
        if self.commit is not None:
            ch.observe(self.commit)
        ch.observe([self.pc_start])
        ch.observe(self.initial_global_cumulative_sum)
        ch.observe([0])


@dataclass
class ShardMainData:
    """types.rs:16-22"""
    chips: List[Chip]                 # `traces`, in commit order
    main_commit: np.ndarray
    main_data: object                 # PData
    chip_ordering: Dict[str, int]
    public_values: np.ndarray
    global_sums: List[np.ndarray] = field(default_factory=list)   # per chip, prover.rs:353-361
    device_traces: List[int] = field(default_factory=list)        # device buffers the prover data borrows (commit_dev)
    ctx: object = None

    def free(self):
        """drop the prover data and the device traces it points into"""
        if self.main_data is not None:
            self.main_data.free()
            self.main_data = None
        for p in self.device_traces:
            self.ctx.dev_free(p)
        self.device_traces = []


class GpuShardProver:
    def __init__(self, ctx, log_blowup=1, num_queries=84, pow_bits=16, num_pv_elts=0):
        self.ctx, self.log_blowup, self.num_queries, self.pow_bits = ctx, log_blowup, num_queries, pow_bits
        self.num_pv_elts = num_pv_elts  # StarkMachine::num_pv_elts (machine.rs:44,129)
        ctx.keep_traces(True)  # LogUp reads the traces themselves between the commits
        self.phase_ms = {}  # host wall clock per phase of the last commit/open (each phase ends synchronised)

    def _tick(self, name, t0):
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    @staticmethod
    def order(chips):
        """sort by (-height, name): prover.rs:264"""
        return sorted(chips, key=lambda c: (-c.height, c.name))

    def generate_trace(self, chip):
        """MachineAir::generate_trace on the device: (device pointer, width) of the padded main trace of a chip given by
        its event records (zk_tracegen_*, the batched twins of the reference's per-row FFI fillers)."""
        if chip.tracegen in self.ctx.ALU_CHIPS:
            return self.ctx.tracegen_alu(chip.tracegen, chip.events, chip.height)
        if chip.tracegen in ("Poseidon2WideDeg3", "Poseidon2WideDeg9"):
            return self.ctx.tracegen_poseidon2_wide(chip.events, chip.height, chip.tracegen.endswith("3"))
        if chip.tracegen == "Cpu":
            return self.ctx.tracegen_cpu(chip.events, chip.height)
        if chip.tracegen == "Poseidon2SkinnyDeg9":
            return self.ctx.tracegen_poseidon2_skinny(chip.events, chip.height)
        raise ValueError(f"{chip.name}: no device trace filler named {chip.tracegen!r}")

    def setup(self, chips, pc_start=0, initial_global_cumulative_sum=None):
        """StarkMachine::setup (machine.rs:330-440): commit to the preprocessed traces of every chip that has one,
        ordered by (-height, name); chip ordering, local_only flags and the constraint counts."""
        pre = sorted([c for c in chips if c.preprocessed is not None], key=lambda c: (-c.preprocessed.shape[0], c.name))
        root, pd = None, None
        if pre:
            root, pd = self.ctx.commit([c.preprocessed for c in pre], [MONTY_ONE] * len(pre), self.log_blowup)
        igcs = np.zeros(14, np.uint32) if initial_global_cumulative_sum is None else np.asarray(
            initial_global_cumulative_sum, np.uint32)
        return ProvingKey(root, int(pc_start), igcs, [c.preprocessed for c in pre], pd,
                          {c.name: i for i, c in enumerate(pre)}, [c.local_only for c in pre],
                          {c.name: self.ctx.air_info(c.air)["num_constraints"] for c in chips})

    def commit(self, chips, public_values=()):
        """MachineProver::commit (prover.rs:258-292)."""
        t0 = time.perf_counter()
        chips = self.order(chips)
        dev = []
        if all(c.main is not None for c in chips):
            root, pd = self.ctx.commit([c.main for c in chips], [MONTY_ONE] * len(chips), self.log_blowup)
            gsums = [c.global_cumulative_sum() for c in chips]
        else:
            # at least one chip's rows are filled on the device from its events: everything is committed from HBM
            # (zk_commit_dev borrows the buffers; they are released with the ShardMainData)
            shapes, gsums = [], []
            for c in chips:
                if c.main is None:
                    ptr, w = self.generate_trace(c)
                else:
                    ptr, w = self.ctx.upload(c.main), c.main.shape[1]
                dev.append(ptr)
                shapes.append((c.height, w))
                if c.commit_scope == "local":
                    gsums.append(np.zeros(14, np.uint32))
                elif c.main is not None:
                    gsums.append(c.global_cumulative_sum())
                else:
                    gsums.append(self.ctx.download(ptr + 4 * (c.height * w - 14), (14,)))
            root, pd = self.ctx.commit_dev(dev, shapes, [MONTY_ONE] * len(chips), self.log_blowup)
        self._tick("commit_main", t0)
        return ShardMainData(chips, root, pd, {c.name: i for i, c in enumerate(chips)},
                             np.asarray(public_values, np.uint32).reshape(-1), gsums, dev, self.ctx)

    def open(self, pk: ProvingKey, data: ShardMainData, challenger: Challenger, inject_witness=-1) -> pf.ShardProof:
        """MachineProver::open (prover.rs:298-653)."""
        t0 = time.perf_counter()
        ctx, chips = self.ctx, data.chips
        main_pd = data.main_data
        pvs = data.public_values
        log_degrees = [c.log_degree for c in chips]
        # prover.rs:322-323: observe_slice(public_values), observe(main_commit) -- consecutive observations are one
        # device round trip (observing slices one after the other = observing their concatenation)
        challenger.observe_many([pvs[:self.num_pv_elts], data.main_commit])
        perm_challenges = challenger.sample_ext(2)                       # prover.rs:326-329
        # permutation trace of every chip (prover.rs:341-364): generated on the device from the retained traces
        ptrs, shapes, local_sums, global_sums = [], [], [], []
        for i, c in enumerate(chips):
            info = ctx.air_info(c.air)
            wq = 4 * info["perm_width"]                                  # flatten_to_base, prover.rs:393
            if info["num_lookups"] > 0:
                prep = pk.data.trace_ptr(pk.chip_ordering[c.name]) if c.name in pk.chip_ordering else 0
                ptr, lcs = ctx.permutation_trace(c.air, prep, main_pd.trace_ptr(i), c.height, perm_challenges)
            else:
                ptr, lcs = 0, np.zeros(4, np.uint32)                     # width 0: generate_permutation_trace's empty matrix
            ptrs.append(ptr)
            shapes.append((c.height, wq))
            local_sums.append(np.asarray(lcs, np.uint32))
            global_sums.append(data.global_sums[i] if data.global_sums else c.global_cumulative_sum())  # prover.rs:353-361
        perm_root, perm_pd = ctx.commit_dev(ptrs, shapes, [MONTY_ONE] * len(ptrs), self.log_blowup)  # prover.rs:401-403
        for ptr in ptrs:
            if ptr:
                ctx.dev_free(ptr)
        # prover.rs:406-413: observe(perm_root); per chip observe_slice(local sum), observe_slice(global sum x, y)
        parts = [perm_root]
        for lcs, gcs in zip(local_sums, global_sums):
            parts += [lcs, gcs[:7], gcs[7:]]
        challenger.observe_many(parts)
        alpha = challenger.sample_ext()                                  # prover.rs:426
        t0 = self._tick("permutation_and_challenges", t0)
        # quotient values per chip, written as chunk matrices (prover.rs:429-488)
        chunk_ptrs, chunk_shapes, chunk_shifts, chunk_bufs = [], [], [], []
        for i, c in enumerate(chips):
            n, lqd = c.log_degree, c.log_quotient_degree
            assert pk.constraints_map[c.name] == ctx.air_info(c.air)["num_constraints"]
            dptr = ctx.quotient(c.air, (main_pd, i), n, lqd, alpha,
                                prep=(pk.data, pk.chip_ordering[c.name]) if c.name in pk.chip_ordering else None,
                                perm=(perm_pd, i), perm_challenges=perm_challenges, public_values=pvs,
                                local_cumsum=local_sums[i], global_cumsum=global_sums[i])
            g = two_adic_generator(n + lqd)
            for k in range(1 << lqd):
                chunk_ptrs.append(dptr + k * (1 << n) * 16)
                chunk_shapes.append((1 << n, 4))
                chunk_shifts.append(monty(3 * pow(g, k, P)))             # split_domains: shift * g^k
            chunk_bufs.append(dptr)  # (per-call state: Chip objects may be shared between provers / threads)
        ctx.sync()
        t0 = self._tick("quotient", t0)
        q_root, q_pd = ctx.commit_dev(chunk_ptrs, chunk_shapes, chunk_shifts, self.log_blowup)   # prover.rs:496-497
        for dptr in chunk_bufs:
            ctx.dev_free(dptr)
        challenger.observe(q_root)                                       # prover.rs:498
        zeta = challenger.sample_ext()                                   # prover.rs:501
        t0 = self._tick("commit_quotient", t0)

        # opening points (prover.rs:503-544)
        def pts(log_degree, local_only):
            if local_only:
                return [zeta]
            # domain.next_point(zeta) = zeta * g_N: the 4 Montgomery words times the canonical g (host scalars)
            g = two_adic_generator(log_degree)
            return [zeta, np.array([(int(z) * g) % P for z in zeta], np.uint32)]

        rounds, points = [], []
        if pk.data is not None:                                          # prover.rs:503-517: ALL pk traces
            rounds.append(pk.data)
            points += [pts(int(t.shape[0]).bit_length() - 1, lo) for t, lo in zip(pk.traces, pk.local_only)]
        rounds.append(main_pd)
        points += [pts(c.log_degree, c.local_only) for c in chips]       # prover.rs:519-533
        rounds.append(perm_pd)
        points += [pts(c.log_degree, False) for c in chips]              # prover.rs:535-540
        rounds.append(q_pd)
        points += [[zeta] for _ in chunk_ptrs]                           # prover.rs:543-544
        flat = pcs_open(ctx, rounds, points, challenger, self.log_blowup, self.num_queries, self.pow_bits,
                        inject_witness)                                  # prover.rs:546-556
        t0 = self._tick("pcs_open", t0)

        # repackaging (prover.rs:558-652)
        k = 0
        rshapes = []
        for r in rounds:
            nm = r.num_matrices()
            rshapes.append(pf.RoundShape([r.height(i) for i in range(nm)], [r.width(i) for i in range(nm)],
                                         [len(points[k + i]) for i in range(nm)]))
            k += nm
        opened, fri = pf.split_flat_proof(flat, rshapes, self.log_blowup, self.num_queries)
        r0 = 1 if pk.data is not None else 0
        prep_vals = opened[0] if pk.data is not None else []
        main_vals, perm_vals, quot_vals = opened[r0], opened[r0 + 1], opened[r0 + 2]

        def air_values(op):
            if len(op) == 2:
                return pf.AirOpenedValues(op[0], op[1])
            return pf.AirOpenedValues(op[0], np.zeros_like(op[0]))       # local_only: next = zeros (prover.rs:566-570)

        chip_values, qi = [], 0
        for i, c in enumerate(chips):
            nch = 1 << c.log_quotient_degree
            pre = air_values(prep_vals[pk.chip_ordering[c.name]]) if c.name in pk.chip_ordering else pf.AirOpenedValues()
            chip_values.append(pf.ChipOpenedValues(
                pre, air_values(main_vals[i]), air_values(perm_vals[i]),
                [quot_vals[qi + j][0] for j in range(nch)], global_sums[i], local_sums[i], log_degrees[i]))
            qi += nch
        sp = pf.ShardProof(pf.ShardCommitment(data.main_commit, perm_root, q_root), chip_values, fri,
                           dict(data.chip_ordering), pvs)
        perm_pd.free()
        q_pd.free()
        self._tick("repackage", t0)
        return sp

    def prove(self, pk, chips, challenger, public_values=()):
        """MachineProver::prove for one shard (prover.rs:660-693); `challenger` is the per-shard clone."""
        data = self.commit(chips, public_values)
        sp = self.open(pk, data, challenger)
        data.free()
        return sp
