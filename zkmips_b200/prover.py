"""Host-side mirror of the reference's per-shard prover, `MachineProver::{commit, open}`
(crates/stark/src/prover.rs:258-292, 298-653), driving libzkgpu through its C ABI.

In Ziren this orchestration is Rust (`CpuProver`); a `GpuProver` would make the same calls through the FFI
crate shown in INTEGRATION.md.  This Python mirror exists so that the whole commit -> quotient -> open flow
can be exercised, verified and timed from the tests and bench.py; the only arithmetic it does itself is
O(1) scalar work per chip (zeta * g, domain shifts), exactly what `Domain::next_point` / `split_domains`
do on the Rust side.  Transcript operations run on the device challenger, everything else in the kernels.

Transcript order: observe main commit; sample 2 permutation challenges; observe permutation
commit and cumulative sums; sample alpha; observe quotient commit; sample zeta; Pcs::open over the rounds
[preprocessed, main, permutation, quotient] with points [zeta, zeta*g] ([zeta] for quotient chunks)."""
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

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
    """One chip of a shard: traces are (height, width) uint32 Montgomery matrices."""
    name: str
    air: str
    main: np.ndarray
    preprocessed: Optional[np.ndarray] = None
    # host-side permutation trace generator (only for chips whose AIR was not compiled with lookups):
    # (perm_challenges (2,4) Montgomery) -> (permutation trace flattened to base (h, 4*perm_width), local_cumsum[4])
    permutation: Optional[Callable] = None
    has_lookups: bool = False  # LogUp trace generated on the device (zk_permutation_trace)
    public_values: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    global_cumsum: np.ndarray = field(default_factory=lambda: np.zeros(14, np.uint32))
    log_quotient_degree: int = 1

    @property
    def log_degree(self):
        return int(np.log2(self.main.shape[0]))


@dataclass
class ShardProof:
    """Flat image of `ShardProof` (crates/stark/src/types.rs:77-83): commitments, the flat PCS proof
    (opened values + FRI proof, layout in include/zkgpu.h) and the data needed to re-derive the transcript."""
    main_commit: np.ndarray
    perm_commit: Optional[np.ndarray]
    quotient_commit: np.ndarray
    prep_commit: Optional[np.ndarray]
    pcs_proof: np.ndarray
    local_cumsums: List[np.ndarray]
    chip_order: List[str]
    points: list
    shapes: list  # per round: list of (lde_height, width)


class GpuShardProver:
    def __init__(self, ctx, log_blowup=1, num_queries=84, pow_bits=16):
        self.ctx, self.log_blowup, self.num_queries, self.pow_bits = ctx, log_blowup, num_queries, pow_bits
        ctx.keep_traces(True)  # LogUp reads the traces themselves between the commits
        self.phase_ms = {}  # host wall clock per phase of the last commit/open (each phase ends synchronised)

    def _tick(self, name, t0):
        import time
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    @staticmethod
    def order(chips):
        """sort by (-height, name): prover.rs:264"""
        return sorted(chips, key=lambda c: (-c.main.shape[0], c.name))

    def setup(self, chips):
        """StarkMachine::setup's commit to the preprocessed traces (crates/stark/src/machine.rs:383-397)."""
        pre = [c for c in self.order(chips) if c.preprocessed is not None]
        if not pre:
            return None, None
        root, pd = self.ctx.commit([c.preprocessed for c in pre], [MONTY_ONE] * len(pre), self.log_blowup)
        return root, pd

    def commit(self, chips):
        """MachineProver::commit (prover.rs:258-292)."""
        import time
        t0 = time.perf_counter()
        chips = self.order(chips)
        root, pd = self.ctx.commit([c.main for c in chips], [MONTY_ONE] * len(chips), self.log_blowup)
        self._tick("commit_main", t0)
        return chips, root, pd

    def open(self, chips, main_root, main_pd, challenger: Challenger, prep_root=None, prep_pd=None,
             inject_witness=-1):
        """MachineProver::open (prover.rs:298-653).  `chips` in commit order."""
        import time
        t0 = time.perf_counter()
        ctx = self.ctx
        pre_idx = {}
        for c in chips:
            if c.preprocessed is not None:
                pre_idx[c.name] = len(pre_idx)
        challenger.observe(main_root)                                   # prover.rs:323
        perm_challenges = challenger.sample_ext(2)                       # prover.rs:326-329
        lookups = {c.name: ctx.air_info(c.air)["num_lookups"] > 0 for c in chips}
        perm_chips = [c for c in chips if c.permutation is not None or lookups[c.name]]
        perm_pd, perm_root, perm_idx, cumsums = None, None, {}, []
        if perm_chips:
            ptrs, shapes = [], []
            for c in perm_chips:                                         # prover.rs:341-364
                if lookups[c.name]:
                    ptr, lcs = ctx.permutation_trace(
                        c.air, prep_pd.trace_ptr(pre_idx[c.name]) if c.name in pre_idx else 0,
                        main_pd.trace_ptr(chips.index(c)), c.main.shape[0], perm_challenges)
                    wq = 4 * ctx.air_info(c.air)["perm_width"]
                else:
                    tr, lcs = c.permutation(perm_challenges)
                    ptr, wq = ctx.upload(tr), tr.shape[1]
                perm_idx[c.name] = len(ptrs)
                ptrs.append(ptr)
                shapes.append((c.main.shape[0], wq))
                cumsums.append(np.asarray(lcs, np.uint32))
            perm_root, perm_pd = ctx.commit_dev(ptrs, shapes, [MONTY_ONE] * len(ptrs), self.log_blowup)  # prover.rs:401-403
            for ptr in ptrs:
                ctx.dev_free(ptr)
            challenger.observe(perm_root)                                # prover.rs:406
            for c, lcs in zip(perm_chips, cumsums):                      # prover.rs:407-413
                challenger.observe(lcs)
                challenger.observe(c.global_cumsum)
        alpha = challenger.sample_ext()                                  # prover.rs:426
        t0 = self._tick("permutation_and_challenges", t0)
        # quotient values per chip, written as chunk matrices (prover.rs:429-488)
        chunk_ptrs, chunk_shapes, chunk_shifts, chunk_bufs = [], [], [], []
        for c in chips:
            n, lqd = c.log_degree, c.log_quotient_degree
            lcs = cumsums[perm_idx[c.name]] if c.name in perm_idx else None
            dptr = ctx.quotient(c.air, (main_pd, chips.index(c)), n, lqd, alpha,
                                prep=(prep_pd, pre_idx[c.name]) if c.name in pre_idx else None,
                                perm=(perm_pd, perm_idx[c.name]) if c.name in perm_idx else None,
                                perm_challenges=perm_challenges, public_values=c.public_values, local_cumsum=lcs,
                                global_cumsum=c.global_cumsum)
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
        rounds, points = [], []

        def two(c):
            # zeta * g_N: multiply the 4 Montgomery words by the Montgomery form of g (host scalar)
            g = two_adic_generator(c.log_degree)
            zg = np.array([(int(z) * g) % P for z in zeta], np.uint32)   # x_monty * g_canonical = (x*g)_monty
            return [zeta, zg]

        if prep_pd is not None:
            rounds.append(prep_pd)
            points += [two(c) for c in chips if c.name in pre_idx]
        rounds.append(main_pd)
        points += [two(c) for c in chips]
        if perm_pd is not None:
            rounds.append(perm_pd)
            points += [two(c) for c in perm_chips]
        rounds.append(q_pd)
        points += [[zeta] for _ in chunk_ptrs]
        proof = pcs_open(ctx, rounds, points, challenger, self.log_blowup, self.num_queries, self.pow_bits,
                         inject_witness)                                 # prover.rs:546-556
        t0 = self._tick("pcs_open", t0)
        shapes = [[(r.height(i), r.width(i)) for i in range(r.num_matrices())] for r in rounds]
        sp = ShardProof(main_root, perm_root, q_root, prep_root, proof, cumsums, [c.name for c in chips], points, shapes)
        if perm_pd is not None:
            perm_pd.free()
        q_pd.free()
        return sp

    def prove(self, chips, challenger, prep=None):
        """MachineProver::prove for one shard (prover.rs:660-693)."""
        chips, root, pd = self.commit(chips)
        prep_root, prep_pd = prep if prep is not None else (None, None)
        sp = self.open(chips, root, pd, challenger, prep_root, prep_pd)
        pd.free()
        return sp
