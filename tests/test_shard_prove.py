"""Quotient kernels and the whole shard flow (commit -> quotient -> quotient commit -> open) against the
oracle: quotient values bit-exact with the numpy restatement of quotient.rs, commitments bit-exact with the
oracle PCS, and the resulting proof accepted by the verifier restatement (PCS verifier transliterated from
crates/recursion/circuit/src/fri.rs + the constraint identity of crates/stark/src/verifier.rs:316-435)."""
import numpy as np
import pytest

from oracle import air_eval as ae
from oracle import binding as ob
from tests import backends, shard_util as su
from zkmips_b200 import Challenger
from zkmips_b200.prover import MONTY_ONE, GpuShardProver

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _backend(name):
    return backends.gpu() if name == "gpu" else backends.emu()


def _natural(lde_bitrev):
    h = lde_bitrev.shape[0]
    bits = int(np.log2(h))
    idx = np.array([ae.bitrev(i, bits) for i in range(h)])
    return ob.from_monty(lde_bitrev[idx])


@pytest.mark.parametrize("be", BACKENDS)
@pytest.mark.parametrize("which", ["fibonacci", "wide", "lookup", "wide1024", "wide4096"])
def test_quotient_values_match_oracle(be, which):
    """`wide1024` (2^10 rows) and `wide4096` (2^8 rows) are the chips bench.py's shard-prove legs time: their
    constraint programs are cut into several kernels (codegen parts of <= 1500 nodes) that ACCUMULATE into the
    quotient buffer (quotient.cuh accumulate path), which the small chips never reach."""
    ctx = _backend(be)
    chip = {"fibonacci": lambda: su.fibonacci_chip(5), "wide": lambda: su.wide_chip(4, 64),
            "lookup": lambda: su.lookup_chip(4), "wide1024": lambda: su.wide_chip(10, 1024, seed=21),
            "wide4096": lambda: su.wide_chip(8, 4096, seed=22)}[which]()
    if which in ("wide1024", "wide4096"):
        assert ctx.air_info(chip.air)["num_kernels"] > 1, "this case must exercise the multi-part accumulate path"
    air = su.AIRS[chip.air]
    n = chip.log_degree
    _, main_pd = ctx.commit([chip.main], [MONTY_ONE], 1)
    prep_pd = perm_pd = None
    chal = su.M(np.arange(10, 18).reshape(2, 4))
    alpha = su.M([3, 1, 4, 1])
    kw = {}
    okw = {}
    if chip.preprocessed is not None:
        _, prep_pd = ctx.commit([chip.preprocessed], [MONTY_ONE], 1)
        kw["prep"] = (prep_pd, 0)
        okw["prep_q"] = _natural(prep_pd.lde(0))
    lcs = None
    if chip.has_lookups:
        # device LogUp trace (zk_permutation_trace) against the numpy restatement of permutation.rs:102-196
        from oracle import logup
        p_c, m_c = chip.canon
        exp_tr, exp_lcs = logup.generate_permutation_trace(air, p_c, m_c, ob.from_monty(chal[0]), ob.from_monty(chal[1]))
        dptr, lcs = ctx.permutation_trace(chip.air, ctx.upload(chip.preprocessed), ctx.upload(chip.main), 1 << n, chal)
        tr = ctx.download(dptr, exp_tr.shape)
        assert (ob.from_monty(tr) == exp_tr).all()
        assert list(ob.from_monty(lcs)) == exp_lcs
        _, perm_pd = ctx.commit([tr], [MONTY_ONE], 1)
        kw["perm"] = (perm_pd, 0)
        okw["perm_q"] = _natural(perm_pd.lde(0))
        okw["lcs"] = ob.from_monty(lcs)
    dptr = ctx.quotient(chip.air, (main_pd, 0), n, 1, alpha, perm_challenges=chal, public_values=chip.public_values,
                        local_cumsum=lcs, global_cumsum=chip.global_cumsum, **kw)
    got = ob.from_monty(ctx.download(dptr, (2, 1 << n, 4)))
    ctx.dev_free(dptr)
    exp = ae.quotient_values(air, n, 1, _natural(main_pd.lde(0)), ob.from_monty(alpha), chal=ob.from_monty(chal),
                             gcs=ob.from_monty(chip.global_cumsum), pvs=ob.from_monty(chip.public_values), **okw)
    # chunk c holds the rows i = c (mod 2)
    assert (got[0] == exp[0::2]).all() and (got[1] == exp[1::2]).all()
    # the quotient of a valid trace is a polynomial of degree < 2N: its top half of coefficients vanish, which
    # the FRI low-degree test checks in test_shard_proof_verifies; here: not identically zero
    assert exp.any()


@pytest.mark.parametrize("be", BACKENDS)
def test_shard_proof_verifies(be):
    """multi-chip shard (BASELINE config 1 stand-in): Fibonacci + wide + lookup chips of different heights."""
    ctx = _backend(be)
    nq, pw = (8, 6) if be == "emu" else (84, 16)
    chips = [su.fibonacci_chip(6), su.wide_chip(4, 64), su.lookup_chip(5), su.fibonacci_chip(3, 2, 5, name="Fib2")]
    prover = GpuShardProver(ctx, 1, nq, pw)
    prep_root, prep_pd = prover.setup(chips)
    ch = Challenger(ctx)
    ch.observe(prep_root)  # stands in for the vk/pc_start observations of machine.rs:79-86
    start = ch.w.copy()
    ordered, root, pd = prover.commit(chips)
    assert [c.name for c in ordered] == ["Fibonacci", "Lookup", "Wide64", "Fib2"]
    # commitment parity with the oracle PCS
    assert (root == ob.pcs_commit([c.main for c in ordered], 1).root).all()
    sp = prover.open(ordered, root, pd, ch, prep_root, prep_pd)
    ok, why = su.verify_shard(sp, ordered, start, 1, nq, pw)
    assert ok, why
    # tampering with an opened value or a commitment must be rejected
    bad = sp.pcs_proof.copy()
    bad[3] ^= 1
    sp_bad = type(sp)(**{**sp.__dict__, "pcs_proof": bad})
    assert not su.verify_shard(sp_bad, ordered, start, 1, nq, pw)[0]
    pd.free()
    prep_pd.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_invalid_trace_is_rejected(be):
    """a trace that violates its AIR still yields well-formed commitments and a FRI proof (any N values
    interpolate to a degree < N chunk), but the verifier's constraint identity at zeta fails."""
    ctx = _backend(be)
    chip = su.fibonacci_chip(5)
    chip.main = chip.main.copy()
    chip.main[7, 1] = (int(chip.main[7, 1]) + 1) % su.P
    prover = GpuShardProver(ctx, 1, 4, 4)
    ch = Challenger(ctx)
    start = ch.w.copy()
    ordered, root, pd = prover.commit([chip])
    sp = prover.open(ordered, root, pd, ch)
    ok, why = su.verify_shard(sp, ordered, start, 1, 4, 4)
    assert not ok and "constraint identity" in why
    pd.free()


@pytest.mark.parametrize("be,log_n", [pytest.param("emu", 11, id="emu-2^11"),
                                      pytest.param("gpu", 11, id="gpu-2^11", marks=pytest.mark.gpu),
                                      pytest.param("gpu", 16, id="gpu-2^16", marks=pytest.mark.gpu)])
def test_permutation_trace_multi_block_scan(be, log_n):
    """running-sum column over more rows than one scan block (1024): block scan + totals + add"""
    from oracle import logup
    ctx = _backend(be)
    chip = su.lookup_chip(log_n, seed=9)
    air = su.AIRS[chip.air]
    chal = su.M(np.arange(20, 28).reshape(2, 4))
    p_c, m_c = chip.canon
    exp_tr, exp_lcs = logup.generate_permutation_trace(air, p_c, m_c, ob.from_monty(chal[0]), ob.from_monty(chal[1]))
    pp, mp = ctx.upload(chip.preprocessed), ctx.upload(chip.main)
    dptr, lcs = ctx.permutation_trace(chip.air, pp, mp, 1 << log_n, chal)
    tr = ctx.download(dptr, exp_tr.shape)
    for p in (pp, mp, dptr):
        ctx.dev_free(p)
    assert (ob.from_monty(tr) == exp_tr).all()
    assert list(ob.from_monty(lcs)) == exp_lcs


@pytest.mark.gpu
def test_two_contexts_prove_concurrently():
    """two shards in flight on one GPU (two contexts, two host threads, shared Chip objects): both proofs must
    equal the proof computed alone -- guards the library's per-context state and the mirror's per-call state."""
    import threading
    from zkmips_b200 import native
    lib = native.load()
    chips = [su.fibonacci_chip(12), su.wide_chip(10, 64), su.lookup_chip(11)]
    ctxs = [lib.ctx_create(0), lib.ctx_create(0)]
    provers = [GpuShardProver(c, 1, 20, 8) for c in ctxs]
    preps = [p.setup(chips) for p in provers]

    def prove(k):
        ch = Challenger(ctxs[k])
        ordered, root, pd = provers[k].commit(chips)
        sp = provers[k].open(ordered, root, pd, ch, *preps[k])
        pd.free()
        return sp

    ref = prove(0)
    out = [[], []]

    def worker(k):
        for _ in range(4):
            out[k].append(prove(k))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(2):
        assert len(out[k]) == 4
        for sp in out[k]:
            assert (sp.pcs_proof == ref.pcs_proof).all()
            assert (sp.quotient_commit == ref.quotient_commit).all()
    for k in range(2):
        preps[k][1].free()
        ctxs[k].destroy()


@pytest.mark.gpu
def test_one_process_two_devices():
    """INTEGRATION.md's dispatch keeps one context per GPU inside ONE process: per-device state (constant
    twiddles, function attributes, pools) must be set up for every device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from zkmips_b200 import native
    lib = native.load()
    roots = []
    m = su.M(np.arange(1 << 14 * 1).reshape(-1, 1) % 1000 + np.arange(24))  # 2^14 x 24
    for dev in (0, 1, 0):
        ctx = lib.ctx_create(dev)
        root, pd = ctx.commit([m], [MONTY_ONE], 1)
        roots.append(root)
        pd.free()
        ctx.destroy()
    assert (roots[0] == roots[1]).all() and (roots[0] == roots[2]).all()
    assert (roots[0] == ob.pcs_commit([m], 1).root).all()


@pytest.mark.parametrize("be", BACKENDS)
def test_quotient_degree_four_chunks(be):
    """log_quotient_degree = 2 with blowup 4 (the shrink configuration): four quotient chunks on shifted domains,
    quotient values against the numpy restatement and the whole proof against the verifier identity
    (recompute_quotient with four chunk domains, crates/stark/src/verifier.rs:400-435)."""
    ctx = _backend(be)
    chip = su.quintic_chip(5)
    air = su.AIRS[chip.air]
    assert air.max_degree() == 5
    n = chip.log_degree
    _, main_pd = ctx.commit([chip.main], [MONTY_ONE], 2)
    alpha = su.M([9, 8, 7, 6])
    dptr = ctx.quotient(chip.air, (main_pd, 0), n, 2, alpha)
    got = ob.from_monty(ctx.download(dptr, (4, 1 << n, 4)))
    ctx.dev_free(dptr)
    lde_nat = _natural(main_pd.lde(0))  # quotient domain = the whole LDE here (n + 2 bits)
    exp = ae.quotient_values(air, n, 2, lde_nat, ob.from_monty(alpha))
    for c in range(4):
        assert (got[c] == exp[c::4]).all()
    main_pd.free()
    nq, pw = (6, 4) if be == "emu" else (42, 16)
    prover = GpuShardProver(ctx, 2, nq, pw)
    ch = Challenger(ctx)
    start = ch.w.copy()
    ordered, root, pd = prover.commit([chip, su.fibonacci_chip(4)])
    sp = prover.open(ordered, root, pd, ch)
    ok, why = su.verify_shard(sp, ordered, start, 2, nq, pw)
    assert ok, why
    pd.free()


@pytest.mark.parametrize("name,log_blowup,num_queries", [("default", 1, 84), ("compressed", 2, 42), ("ultra_compressed", 3, 28)])
@pytest.mark.parametrize("be", BACKENDS)
def test_reference_fri_configs(be, name, log_blowup, num_queries):
    """The three FRI configurations of the reference (crates/stark/src/kb31_poseidon2.rs:203-241: default_fri_config
    for core shards, compressed_fri_config for compress, ultra_compressed_fri_config for shrink/wrap; 16 proof-of-work
    bits each): a shard with chips of different heights, lookups and a preprocessed trace proves and verifies under
    every one of them, and the main commitment equals the oracle's for that blowup."""
    ctx = _backend(be)
    nq, pw = (min(num_queries, 6), 5) if be == "emu" else (num_queries, 16)
    chips = [su.fibonacci_chip(6), su.wide_chip(4, 64), su.lookup_chip(5)]
    prover = GpuShardProver(ctx, log_blowup, nq, pw)
    prep_root, prep_pd = prover.setup(chips)
    ch = Challenger(ctx)
    ch.observe(prep_root)
    start = ch.w.copy()
    ordered, root, pd = prover.commit(chips)
    assert (root == ob.pcs_commit([c.main for c in ordered], log_blowup).root).all()
    sp = prover.open(ordered, root, pd, ch, prep_root, prep_pd)
    ok, why = su.verify_shard(sp, ordered, start, log_blowup, nq, pw)
    assert ok, f"{name}: {why}"
    # a proof made for one configuration must not verify under another blowup
    other = 1 if log_blowup != 1 else 2
    assert not su.verify_shard(sp, ordered, start, other, nq, pw)[0]
    pd.free()
    prep_pd.free()
