"""Synthetic shards for the AIRs compiled into libzkgpu (valid traces, so the quotients are low degree) and the
oracle-side VERIFIER of a `ShardProof`.

The verifier is written from the reference's verifier alone -- `StarkMachine::verify` (crates/stark/src/machine.rs:
600-655) and `Verifier::verify_shard` / `verify_opening_shape` / `verify_constraints` (crates/stark/src/verifier.rs:
30-246, 248-313, 316-435) -- and consumes only what the reference's verifier consumes: the verifying key, the
machine's chip definitions, the challenger and the structured proof (types.rs:77-83).  It never looks at the prover
mirror's intermediate state, so an observation the prover drops (public values, a chip's cumulative sums, the
`local_only` opening points, the preprocessed round) desynchronises the transcript and the PCS check fails.
TEST INFRASTRUCTURE: uses the oracle's challenger, its transliterated `Pcs::verify` and the numpy/pure-Python
restatement of the constraint identity."""
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from oracle import air_eval as ae
from oracle import binding as ob
from oracle import binding_fri as bf
from zkmips_b200 import proof as pf
from zkmips_b200.air import library
from zkmips_b200.prover import Chip
from zkmips_b200.synth import (LOOKUP_PV3, M, fibonacci_chip, global_chip, local_bool_chip, lookup_chip,  # noqa: F401
                                lookup_side_chips, public_values_for, quintic_chip, wide_chip)

P = ae.P
AIRS = {a.name: a for a in library.all_airs()}


# ---------------------------------------------------------------------------------------------- verifier
@dataclass
class VerifyingKey:
    """StarkVerifyingKey (machine.rs:88-105)"""
    commit: object
    pc_start: int
    initial_global_cumulative_sum: np.ndarray
    chip_information: List[Tuple[str, int, Tuple[int, int]]]   # (name, log2 of the domain size, (height, width))
    chip_ordering: Dict[str, int]


def vk_of(pk):
    """the verifying half of StarkMachine::setup's output (machine.rs:423-438)"""
    info = [None] * len(pk.traces)
    for name, i in pk.chip_ordering.items():
        h, w = pk.traces[i].shape
        info[i] = (name, int(h).bit_length() - 1, (int(h), int(w)))
    return VerifyingKey(pk.commit, pk.pc_start, pk.initial_global_cumulative_sum, info, dict(pk.chip_ordering))


def _ext_from(words_monty):
    return [int(x) for x in ob.from_monty(np.asarray(words_monty, np.uint32))]


def _is_zero(words):
    return not np.asarray(words).any()


def _next_point(zeta, log_n):
    """domain.next_point(zeta) = zeta * g_N (Montgomery words times the canonical generator)"""
    g = ae.two_adic_generator(log_n)
    return np.array([(int(z) * g) % P for z in zeta], np.uint32)


def verify_shard(vk: VerifyingKey, machine_chips: Dict[str, Chip], ch, sp: pf.ShardProof, log_blowup=1, num_queries=84,
                 pow_bits=16):
    """Verifier::verify_shard (verifier.rs:30-246).  `machine_chips`: the machine's chip DEFINITIONS by name (air,
    local_only, commit_scope, log_quotient_degree -- traces are not read); `ch`: the oracle challenger, already past
    vk.observe_into and the public values (machine.rs:609-625).  Returns (ok, reason)."""
    order = sorted(sp.chip_ordering.items(), key=lambda kv: kv[1])          # shard_chips_ordered
    if [i for _, i in order] != list(range(len(order))):
        return False, "chip_ordering is not a permutation"
    if any(name not in machine_chips for name, _ in order):
        return False, "unknown chip in chip_ordering"
    chips = [machine_chips[name] for name, _ in order]
    vals = sp.opened_values
    if len(chips) != len(vals):
        return False, "ChipOpeningLengthMismatch"                          # verifier.rs:51-53
    log_degrees = [v.log_degree for v in vals]
    lqds = [c.log_quotient_degree for c in chips]
    com = sp.commitment
    bf.observe(ch, com.main_commit)                                         # verifier.rs:86
    chal = [bf.sample_ext(ch), bf.sample_ext(ch)]                           # verifier.rs:88-89
    bf.observe(ch, com.permutation_commit)                                  # verifier.rs:91
    for v, c in zip(vals, chips):                                           # verifier.rs:94-115
        bf.observe(ch, v.local_cumulative_sum)
        bf.observe(ch, v.global_cumulative_sum[:7])
        bf.observe(ch, v.global_cumulative_sum[7:])
        if c.commit_scope == "local" and not _is_zero(v.global_cumulative_sum):
            return False, "global cumulative sum is non-zero, but chip is Local"
        air = AIRS[c.air]
        if not (air.sends or air.receives) and not _is_zero(v.local_cumulative_sum):
            return False, "local cumulative sum is non-zero, but no local lookups"
    alpha = bf.sample_ext(ch)                                               # verifier.rs:117
    bf.observe(ch, com.quotient_commit)                                     # verifier.rs:120
    zeta = bf.sample_ext(ch)                                                # verifier.rs:122

    def two(v: pf.AirOpenedValues, log_n, local_only):
        if local_only:
            return [zeta], [v.local]
        return [zeta, _next_point(zeta, log_n)], [v.local, v.next]

    # rounds: (commitment, [(log2 LDE height, points, opened values)])       verifier.rs:124-205
    rounds = []
    if vk.commit is not None:
        mats = []
        for name, log_n, _dims in vk.chip_information:                      # verifier.rs:124-140
            if name not in sp.chip_ordering:
                return False, f"preprocessed chip {name} is not in the shard"
            i = sp.chip_ordering[name]
            mats.append((log_n,) + two(vals[i].preprocessed, log_n, chips[i].local_only))
        rounds.append((vk.commit, mats))
    rounds.append((com.main_commit, [(n,) + two(v.main, n, c.local_only) for v, c, n in zip(vals, chips, log_degrees)]))
    rounds.append((com.permutation_commit, [(n,) + two(v.permutation, n, False) for v, n in zip(vals, log_degrees)]))
    rounds.append((com.quotient_commit, [(n, [zeta], [q]) for v, n in zip(vals, log_degrees) for q in v.quotient]))
    # Pcs::verify (verifier.rs:207-210) through the transliterated PCS verifier, on the flat image of these rounds
    opened = [[list(op) for (_, _, op) in mats] for _, mats in rounds]
    flat = pf.join_flat_proof(opened, sp.opening_proof)
    roots = [r for r, _ in rounds]
    n_mats = [len(m) for _, m in rounds]
    hs = [1 << (n + log_blowup) for _, mats in rounds for (n, _, _) in mats]
    ws = [int(np.asarray(op[0]).shape[0]) for _, mats in rounds for (_, _, op) in mats]
    points = [pts for _, mats in rounds for (_, pts, _) in mats]
    try:
        rc = bf.pcs_verify(roots, n_mats, hs, ws, points, ch, flat, log_blowup, num_queries, pow_bits)
    except Exception as e:  # malformed shapes
        return False, f"pcs_verify raised {e}"
    if rc != 1:
        return False, f"InvalidopeningArgument (pcs_verify rc={rc})"
    zeta_c, alpha_c = _ext_from(zeta), _ext_from(alpha)
    chal_c = [_ext_from(c) for c in chal]
    pvs_c = [int(x) for x in ob.from_monty(sp.public_values)]
    for c, v, n, lqd in zip(chips, vals, log_degrees, lqds):
        air = AIRS[c.air]
        # verify_opening_shape (verifier.rs:248-313)
        if v.preprocessed.local.shape[0] != air.prep_width or v.preprocessed.next.shape[0] != air.prep_width:
            return False, f"OpeningShapeError({c.name}): PreprocessedWidthMismatch"
        if v.main.local.shape[0] != air.main_width or v.main.next.shape[0] != air.main_width:
            return False, f"OpeningShapeError({c.name}): MainWidthMismatch"
        if v.permutation.local.shape[0] != 4 * air.perm_width or v.permutation.next.shape[0] != 4 * air.perm_width:
            return False, f"OpeningShapeError({c.name}): PermutationWidthMismatch"
        if len(v.quotient) != (1 << lqd) or any(q.shape != (4, 4) for q in v.quotient):
            return False, f"OpeningShapeError({c.name}): QuotientWidthMismatch"

        def canon(m):
            return [[int(x) for x in row] for row in ob.from_monty(np.asarray(m, np.uint32)).reshape(-1, 4)]

        def unflatten(m):  # verifier.rs:365-371: ext column j = sum_e X^e * opened[4j + e]
            rows = canon(m)
            cols = []
            for j in range(len(rows) // 4):
                acc = [0, 0, 0, 0]
                for e in range(4):
                    mono = [0, 0, 0, 0]
                    mono[e] = 1
                    acc = ae.ext_add(acc, ae.ext_mul(mono, rows[4 * j + e]))
                cols.append(acc)
            return cols

        op = {"main": [canon(v.main.local), canon(v.main.next)],
              "prep": [canon(v.preprocessed.local), canon(v.preprocessed.next)],
              "perm": [unflatten(v.permutation.local), unflatten(v.permutation.next)]}
        ok = ae.verify_constraints(air, op, [canon(q) for q in v.quotient], n, lqd, zeta_c, alpha_c, chal_c,
                                   _ext_from(v.local_cumulative_sum), _ext_from(v.global_cumulative_sum), pvs_c)
        if not ok:
            return False, f"OodEvaluationMismatch({c.name})"
    if not _is_zero(sp.local_cumulative_sum()):                             # verifier.rs:236-244
        return False, "local cumulative sum is not zero"
    return True, "ok"


def machine_verify(vk, machine_chips, proofs, num_pv_elts, log_blowup=1, num_queries=84, pow_bits=16):
    """StarkMachine::verify (machine.rs:600-655) without the cross-shard global-sum check (curve arithmetic on the
    septic digest is outside this path): vk.observe_into, then per shard a CLONE of the challenger that observes
    public_values[0..num_pv_elts] before verify_shard."""
    ch = bf.new_challenger()
    if vk.commit is not None:
        bf.observe(ch, vk.commit)
    bf.observe(ch, [vk.pc_start])
    bf.observe(ch, vk.initial_global_cumulative_sum)
    bf.observe(ch, [0])
    if not proofs:
        return False, "EmptyProof"
    for i, sp in enumerate(proofs):
        shard_ch = bf.Challenger.from_words(ch.words())
        bf.observe(shard_ch, sp.public_values[:num_pv_elts])
        ok, why = verify_shard(vk, machine_chips, shard_ch, sp, log_blowup, num_queries, pow_bits)
        if not ok:
            return False, f"shard {i}: {why}"
    return True, "ok"
