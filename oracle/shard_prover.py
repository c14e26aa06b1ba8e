"""oracle/shard_prover.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

CPU restatement of the reference's per-shard prover, `StarkMachine::setup` (crates/stark/src/machine.rs:330-440) and
`CpuProver::{commit, open}` (crates/stark/src/prover.rs:258-292, 298-653), composed from the oracle's own pieces:
  Pcs::commit            oracle/zk_oracle.c   ork_pcs_commit            (C, OpenMP)
  permutation trace      oracle/logup.py      generate_permutation_trace (numpy; permutation.rs:102-196)
  quotient_values        oracle/air_eval.py   quotient_values            (numpy; quotient.rs:19-171)
  Pcs::open              oracle/zk_oracle_fri.c ork_pcs_open             (C)
  DuplexChallenger       oracle/zk_oracle_fri.c
It produces the same structured `ShardProof` (zkmips_b200.proof, a plain container + wire format; no device code) as
the GPU path, so a shard proof can be compared BYTE FOR BYTE (bincode image) -- identical Merkle roots, quotient
commitments, opened values and FRI query openings -- and it is the CPU leg `bench.py` times beside the GPU shard prover
(`cpu_baseline.kind = "port"`).  grind takes the smallest witness, like the library (Plonky3's is non-deterministic)."""
import time

import numpy as np

from zkmips_b200 import proof as pf

from . import air_eval as ae
from . import binding as ob
from . import binding_fri as bf
from . import logup

P = ae.P
MONTY_ONE = 0x01FFFFFE


def _natural(lde_bitrev):
    """committed LDE (rows bit-reversed, Montgomery) -> natural order, canonical  (= get_evaluations_on_domain +
    to_row_major_matrix, prover.rs:437-445, for a quotient domain of the LDE's size)"""
    h = lde_bitrev.shape[0]
    bits = int(h).bit_length() - 1
    idx = np.array([ae.bitrev(i, bits) for i in range(h)], dtype=np.int64) if h > 1 else np.zeros(1, np.int64)
    return ob.from_monty(lde_bitrev[idx]) if lde_bitrev.shape[1] else np.zeros((h, 0), np.uint32)


class OracleProvingKey:
    def __init__(self, commit, pc_start, igcs, traces, data, chip_ordering, local_only, constraints_map):
        self.commit, self.pc_start, self.initial_global_cumulative_sum = commit, pc_start, igcs
        self.traces, self.data, self.chip_ordering, self.local_only = traces, data, chip_ordering, local_only
        self.constraints_map = constraints_map

    def observe_into(self, ch):
        if self.commit is not None:
            bf.observe(ch, self.commit)
        bf.observe(ch, [self.pc_start])
        bf.observe(ch, self.initial_global_cumulative_sum)
        bf.observe(ch, [0])


class OracleShardProver:
    """chips: objects with .name .air .main .preprocessed .local_only .commit_scope .log_quotient_degree
    (zkmips_b200.prover.Chip); airs: {air name: ir.Air}"""

    def __init__(self, airs, log_blowup=1, num_queries=84, pow_bits=16, num_pv_elts=0):
        self.airs, self.log_blowup, self.num_queries, self.pow_bits = airs, log_blowup, num_queries, pow_bits
        self.num_pv_elts = num_pv_elts
        self.phase_s = {}

    def _tick(self, name, t0):
        self.phase_s[name] = self.phase_s.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def setup(self, chips, pc_start=0, initial_global_cumulative_sum=None):
        pre = sorted([c for c in chips if c.preprocessed is not None], key=lambda c: (-c.preprocessed.shape[0], c.name))
        root = tree = None
        if pre:
            tree = ob.pcs_commit([c.preprocessed for c in pre], self.log_blowup)
            root = tree.root
        igcs = np.zeros(14, np.uint32) if initial_global_cumulative_sum is None else np.asarray(
            initial_global_cumulative_sum, np.uint32)
        return OracleProvingKey(root, int(pc_start), igcs, [c.preprocessed for c in pre], tree,
                                {c.name: i for i, c in enumerate(pre)}, [c.local_only for c in pre],
                                {c.name: self.airs[c.air].num_constraints for c in chips})

    def prove(self, pk, chips, ch, public_values=()):
        """commit + open for one shard; `ch` = the shard's clone of the machine challenger (oracle Challenger)."""
        t0 = time.perf_counter()
        lb = self.log_blowup
        chips = sorted(chips, key=lambda c: (-c.main.shape[0], c.name))              # prover.rs:264
        pvs = np.asarray(public_values, np.uint32).reshape(-1)
        main_tree = ob.pcs_commit([c.main for c in chips], lb)                        # prover.rs:277
        t0 = self._tick("commit_main", t0)
        bf.observe(ch, pvs[:self.num_pv_elts])                                        # prover.rs:322
        bf.observe(ch, main_tree.root)                                                # prover.rs:323
        chal = [bf.sample_ext(ch), bf.sample_ext(ch)]                                 # prover.rs:326-329
        chal_c = [ob.from_monty(c) for c in chal]
        perm_traces, local_sums, global_sums = [], [], []
        for c in chips:                                                               # prover.rs:341-364
            air = self.airs[c.air]
            prep_c = ob.from_monty(c.preprocessed) if c.preprocessed is not None else None
            tr, lcs = logup.generate_permutation_trace(air, prep_c, ob.from_monty(c.main), chal_c[0], chal_c[1])
            perm_traces.append(ob.to_monty(tr) if tr.size else np.zeros((c.main.shape[0], 0), np.uint32))
            local_sums.append(ob.to_monty(np.array(lcs, np.uint32)))
            global_sums.append(np.zeros(14, np.uint32) if c.commit_scope == "local"
                               else np.ascontiguousarray(c.main.reshape(-1)[-14:], dtype=np.uint32))
        perm_tree = ob.pcs_commit(perm_traces, lb)                                    # prover.rs:401-403
        bf.observe(ch, perm_tree.root)                                                # prover.rs:406
        for lcs, gcs in zip(local_sums, global_sums):                                 # prover.rs:407-413
            bf.observe(ch, lcs)
            bf.observe(ch, gcs[:7])
            bf.observe(ch, gcs[7:])
        alpha = bf.sample_ext(ch)                                                     # prover.rs:426
        t0 = self._tick("permutation_and_challenges", t0)
        chunks, shifts = [], []
        for i, c in enumerate(chips):                                                 # prover.rs:429-488
            air = self.airs[c.air]
            n, lqd = int(c.main.shape[0]).bit_length() - 1, c.log_quotient_degree
            qsize = 1 << (n + lqd)

            def on_q(tree, idx):
                # the quotient domain is a prefix of the committed LDE in bit-reversed order when lqd <= log_blowup
                return _natural(tree.matrix(idx)[:qsize])

            kw = {}
            if c.name in pk.chip_ordering:
                kw["prep_q"] = on_q(pk.data, pk.chip_ordering[c.name])
            if air.perm_width:
                kw["perm_q"] = on_q(perm_tree, i)
            q = ae.quotient_values(air, n, lqd, on_q(main_tree, i), ob.from_monty(alpha), chal=chal_c,
                                   lcs=ob.from_monty(local_sums[i]), gcs=ob.from_monty(global_sums[i]),
                                   pvs=ob.from_monty(pvs), **kw)
            g = ae.two_adic_generator(n + lqd)
            for k in range(1 << lqd):                                                 # split_evals / split_domains
                chunks.append(ob.to_monty(np.ascontiguousarray(q[k::1 << lqd])))
                shifts.append(int(ob.to_monty(np.array([3 * pow(g, k, P) % P], np.uint32))[0]))
        t0 = self._tick("quotient", t0)
        q_tree = ob.pcs_commit(chunks, lb, shifts)                                    # prover.rs:496-497
        bf.observe(ch, q_tree.root)                                                   # prover.rs:498
        zeta = bf.sample_ext(ch)                                                      # prover.rs:501
        t0 = self._tick("commit_quotient", t0)

        def pts(log_degree, local_only):
            if local_only:
                return [zeta]
            g = ae.two_adic_generator(log_degree)
            return [zeta, np.array([(int(z) * g) % P for z in zeta], np.uint32)]

        trees, points = [], []
        if pk.data is not None:                                                       # prover.rs:503-517
            trees.append(pk.data)
            points += [pts(int(t.shape[0]).bit_length() - 1, lo) for t, lo in zip(pk.traces, pk.local_only)]
        trees.append(main_tree)
        points += [pts(int(c.main.shape[0]).bit_length() - 1, c.local_only) for c in chips]
        trees.append(perm_tree)
        points += [pts(int(c.main.shape[0]).bit_length() - 1, False) for c in chips]
        trees.append(q_tree)
        points += [[zeta] for _ in chunks]
        flat = bf.pcs_open(trees, points, ch, lb, self.num_queries, self.pow_bits)    # prover.rs:546-556
        t0 = self._tick("pcs_open", t0)
        k = 0
        rshapes = []
        for t in trees:
            nm = t.num_matrices
            rshapes.append(pf.RoundShape([t.dims(i)[0] for i in range(nm)], [t.dims(i)[1] for i in range(nm)],
                                         [len(points[k + i]) for i in range(nm)]))
            k += nm
        opened, fri = pf.split_flat_proof(flat, rshapes, lb, self.num_queries)
        r0 = 1 if pk.data is not None else 0
        prep_vals = opened[0] if pk.data is not None else []
        main_vals, perm_vals, quot_vals = opened[r0], opened[r0 + 1], opened[r0 + 2]

        def air_values(op):
            return pf.AirOpenedValues(op[0], op[1]) if len(op) == 2 else pf.AirOpenedValues(op[0], np.zeros_like(op[0]))

        vals, qi = [], 0
        for i, c in enumerate(chips):                                                 # prover.rs:558-652
            nch = 1 << c.log_quotient_degree
            pre = air_values(prep_vals[pk.chip_ordering[c.name]]) if c.name in pk.chip_ordering else pf.AirOpenedValues()
            vals.append(pf.ChipOpenedValues(pre, air_values(main_vals[i]), air_values(perm_vals[i]),
                                            [quot_vals[qi + j][0] for j in range(nch)], global_sums[i], local_sums[i],
                                            int(c.main.shape[0]).bit_length() - 1))
            qi += nch
        return pf.ShardProof(pf.ShardCommitment(main_tree.root, perm_tree.root, q_tree.root), vals, fri,
                             {c.name: i for i, c in enumerate(chips)}, pvs)
