"""Synthetic shards for the AIRs compiled into libzkgpu (valid traces, so the quotients are low degree) and
an oracle-side verifier of a ShardProof: the transliterated PCS verifier plus the reference's
verify_constraints identity (crates/stark/src/verifier.rs:316-435)."""
import numpy as np

from oracle import air_eval as ae
from oracle import binding as ob
from oracle import binding_fri as bf
from zkmips_b200.air import library
from zkmips_b200.prover import Chip

P = ae.P
AIRS = {a.name: a for a in library.all_airs()}


def M(canon):
    return ob.to_monty(np.asarray(canon, dtype=np.uint64) % P)


def fibonacci_chip(log_n, a=1, b=1, name="Fibonacci"):
    """generate_trace_rows of crates/stark/src/stark_testing.rs:63-81"""
    n = 1 << log_n
    t = np.zeros((n, 2), np.uint64)
    t[0] = (a, b)
    for i in range(1, n):
        t[i, 0] = t[i - 1, 1]
        t[i, 1] = (t[i - 1, 0] + t[i - 1, 1]) % P
    return Chip(name, "fibonacci", M(t), public_values=M([a, b, t[n - 1, 1]]))


def wide_chip(log_n, width=64, seed=1, name=None):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    g = width // 4
    a = rng.integers(0, 2, (n, g))
    b = rng.integers(0, 2, (n, g))
    c = rng.integers(0, 2, (n, g))
    a[1:] = c[:-1]  # next.a = c on transitions
    t = np.zeros((n, width), np.uint64)
    t[:, 0::4], t[:, 1::4], t[:, 2::4], t[:, 3::4] = a, b, a ^ b, c
    return Chip(name or f"Wide{width}", f"wide_bitwise_{width}", M(t))


def quintic_chip(log_n, seed=5, name="Quintic"):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    a = (np.arange(n, dtype=np.uint64) + 3) % P
    b = rng.integers(0, P, n).astype(np.uint64)
    b[0] = 1
    a2 = a * a % P
    d = a2 * a2 % P * b % P
    return Chip(name, "quintic", M(np.stack([a, b, d], axis=1)), log_quotient_degree=2)


def lookup_chip(log_n, seed=3, name="Lookup"):
    """valid trace for library.lookup_pair (its LogUp permutation trace is generated on the device)"""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    pv0 = 7
    p = rng.integers(0, P, (n, 2)).astype(np.uint64)
    m = np.zeros((n, 5), np.uint64)
    m[0, 0] = 5
    for i in range(1, n):
        m[i, 0] = (m[i - 1, 0] + pv0) % P
    m[:, 1] = rng.integers(0, P, n)
    m[:, 2] = (m[:, 0] * m[:, 1] + p[:, 0]) % P
    m[:, 3] = rng.integers(0, 2, n)
    m[:, 4] = rng.integers(0, 5, n)
    c = Chip(name, "lookup_pair", M(m), preprocessed=M(p), public_values=M([pv0]), global_cumsum=M(np.arange(1, 15)))
    c.has_lookups = True
    c.canon = (p, m)
    return c


def _ext_from(words_monty):
    return [int(x) for x in ob.from_monty(np.asarray(words_monty, np.uint32))]


def verify_shard(sp, chips, challenger_words, log_blowup=1, num_queries=84, pow_bits=16):
    """Verifier side (crates/stark/src/verifier.rs:30-246 restricted to what this repo proves): re-derive the
    challenges from the transcript, run the transliterated Pcs::verify, then verify_constraints per chip.
    chips: in commit order (sp.chip_order)."""
    ch = bf.Challenger.from_words(challenger_words)
    bf.observe(ch, sp.main_commit)
    chal = [bf.sample_ext(ch), bf.sample_ext(ch)]
    perm_chips = [c for c in chips if c.permutation is not None or c.has_lookups]
    if sp.perm_commit is not None:
        bf.observe(ch, sp.perm_commit)
        for c, lcs in zip(perm_chips, sp.local_cumsums):
            bf.observe(ch, lcs)
            bf.observe(ch, c.global_cumsum)
    alpha = bf.sample_ext(ch)
    bf.observe(ch, sp.quotient_commit)
    zeta = bf.sample_ext(ch)
    roots = ([sp.prep_commit] if sp.prep_commit is not None else []) + [sp.main_commit] + \
        ([sp.perm_commit] if sp.perm_commit is not None else []) + [sp.quotient_commit]
    n_mats = [len(s) for s in sp.shapes]
    hs = [h for s in sp.shapes for h, _ in s]
    ws = [w for s in sp.shapes for _, w in s]
    # the prover's points must be the ones the verifier derives
    for pts in sp.points:
        assert (np.asarray(pts[0]) == zeta).all()
    rc = bf.pcs_verify(roots, n_mats, hs, ws, sp.points, ch, sp.pcs_proof, log_blowup, num_queries, pow_bits)
    if rc != 1:
        return False, f"pcs_verify rc={rc}"
    # split the opened values: round -> matrix -> point -> width ext
    off = 0
    opened = []
    k = 0
    for s in sp.shapes:
        rnd = []
        for (_, w) in s:
            mat = []
            for _ in sp.points[k]:
                vals = ob.from_monty(sp.pcs_proof[off:off + 4 * w]).reshape(w, 4)
                mat.append([[int(x) for x in v] for v in vals])
                off += 4 * w
            rnd.append(mat)
            k += 1
        opened.append(rnd)
    r = 0
    prep_round = None
    if sp.prep_commit is not None:
        prep_round, r = opened[r], r + 1
    main_round, r = opened[r], r + 1
    perm_round = None
    if sp.perm_commit is not None:
        perm_round, r = opened[r], r + 1
    quot_round = opened[r]
    zeta_c, alpha_c = _ext_from(zeta), _ext_from(alpha)
    chal_c = [_ext_from(c) for c in chal]
    pi = qi = ppi = 0
    for ci, c in enumerate(chips):
        air = AIRS[c.air]
        op = {"main": main_round[ci]}
        if c.preprocessed is not None:
            op["prep"] = prep_round[ppi]
            ppi += 1
        lcs = (0, 0, 0, 0)
        if c.permutation is not None or c.has_lookups:
            # unflatten: ext column j = sum_e X^e * opened[4j+e]  (verifier.rs:365-371)
            rows = []
            for prow in perm_round[pi]:
                cols = []
                for j in range(len(prow) // 4):
                    acc = [0, 0, 0, 0]
                    for e in range(4):
                        mono = [0, 0, 0, 0]
                        mono[e] = 1
                        acc = ae.ext_add(acc, ae.ext_mul(mono, prow[4 * j + e]))
                    cols.append(acc)
                rows.append(cols)
            op["perm"] = rows
            lcs = _ext_from(sp.local_cumsums[pi])
            pi += 1
        nch = 1 << c.log_quotient_degree
        chunks = [quot_round[qi + k][0] for k in range(nch)]
        qi += nch
        ok = ae.verify_constraints(air, op, chunks, c.log_degree, c.log_quotient_degree, zeta_c, alpha_c, chal_c, lcs,
                                   _ext_from(c.global_cumsum), [int(x) for x in ob.from_monty(c.public_values)])
        if not ok:
            return False, f"constraint identity failed for {c.name}"
    return True, "ok"
