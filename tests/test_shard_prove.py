"""Quotient kernels and the whole shard flow (commit -> quotient -> quotient commit -> open) against the
oracle: quotient values bit-exact with the numpy restatement of quotient.rs, commitments bit-exact with the
oracle PCS, and the resulting proof accepted by the verifier restatement (PCS verifier transliterated from
crates/recursion/circuit/src/fri.rs + the constraint identity of crates/stark/src/verifier.rs:316-435)."""
import numpy as np
import pytest

from oracle import air_eval as ae
from oracle import binding as ob
from tests import backends, shard_util as su
from zkmips_b200 import Challenger, synth
from zkmips_b200.prover import MONTY_ONE, GpuShardProver

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _backend(name):
    return backends.gpu() if name == "gpu" else backends.emu()


def _natural(lde_bitrev):
    h = lde_bitrev.shape[0]
    bits = int(np.log2(h))
    idx = np.array([ae.bitrev(i, bits) for i in range(h)])
    return ob.from_monty(lde_bitrev[idx])


def p2w_host_chip(log_n, degree, **kw):
    """Poseidon2WideDeg{3,9} with a HOST main trace from the oracle filler (the CPU provers and the quotient oracle need
    one; the product fills these rows on the device)"""
    c = synth.poseidon2_wide_chip(log_n, degree, **kw)
    c.main = ob.poseidon2_wide_trace(c.events, c.rows, degree == 3)
    c.canon = (ob.from_monty(c.preprocessed).astype(np.uint64), ob.from_monty(c.main).astype(np.uint64))
    return c


COMPRESS_SMALL = dict(log_var=7, log_ext=4, log_sel=5, log_bf=5, log_exp=5, pv=True)


@pytest.mark.parametrize("be", BACKENDS)
@pytest.mark.parametrize("which", ["fibonacci", "wide", "lookup", "wide1024", "wide4096", "global", "local_bool", "AddSub",
                                   "Lt", "Bitwise", "Poseidon2WideDeg3", "Poseidon2WideDeg9", "MemoryConst", "BaseAlu",
                                   "MemoryVar", "ExtAlu", "Select", "BatchFRI", "ExpReverseBitsLen", "PublicValues", "FriFold", "Poseidon2SkinnyDeg9", "MovCond", "Jump", "Branch", "ShiftLeft", "CloClz", "Byte", "Program", "SyscallCore", "SyscallPrecompile", "MemoryLocal", "ShiftRight", "Mul", "Cpu", "DivRem"])
def test_quotient_values_match_oracle(be, which):
    """`wide1024` (2^10 rows) and `wide4096` (2^8 rows) are the chips bench.py's shard-prove legs time: their
    constraint programs are cut into several kernels (codegen parts of <= 1500 nodes) that ACCUMULATE into the
    quotient buffer (quotient.cuh accumulate path), which the small chips never reach."""
    if which == "Byte" and be == "emu":
        pytest.skip("the 2^16-row table takes the emulator half a minute; its GPU twin runs")
    ctx = _backend(be)
    chip = {"fibonacci": lambda: su.fibonacci_chip(5), "wide": lambda: su.wide_chip(4, 64),
            "lookup": lambda: su.lookup_chip(4), "wide1024": lambda: su.wide_chip(10, 1024, seed=21),
            "wide4096": lambda: su.wide_chip(8, 4096, seed=22), "global": lambda: su.global_chip(5),
            "local_bool": lambda: su.local_bool_chip(4),
            # real Ziren chips transcribed from their Air::eval (library.add_sub / lt / bitwise), real lookups
            "AddSub": lambda: synth.add_sub_chip(6), "Lt": lambda: synth.lt_chip(6),
            "Bitwise": lambda: synth.bitwise_chip(5),
            # recursion chip (library.poseidon2_wide): 313 / 172 main + 49 preprocessed columns, 32 memory sends;
            # DEGREE 9 has log_quotient_degree 3 (8 chunks, LogUp batches of 8)
            "Poseidon2WideDeg3": lambda: p2w_host_chip(5, 3), "Poseidon2WideDeg9": lambda: p2w_host_chip(3, 9),
            "MemoryConst": lambda: synth.recursion_program_chips(5, 4, 5)[0],
            "BaseAlu": lambda: synth.recursion_program_chips(5, 4, 5)[1],
            "Select": lambda: synth.recursion_program_chips(5, 4, 5, log_var=6, log_ext=4, log_sel=5)[3],
            "MemoryVar": lambda: synth.recursion_program_chips(5, 4, 5, log_var=6, log_ext=4, log_sel=5)[4],
            "ExtAlu": lambda: synth.recursion_program_chips(5, 4, 5, log_var=6, log_ext=4, log_sel=5)[5],
            # the chips that complete the compress machine; they read the NEXT row (transition constraints) and
            # PublicValues the shard's 231 public values
            "BatchFRI": lambda: synth.recursion_program_chips(5, 4, 5, **COMPRESS_SMALL)[6],
            "ExpReverseBitsLen": lambda: synth.recursion_program_chips(5, 4, 5, **COMPRESS_SMALL)[7],
            "PublicValues": lambda: synth.recursion_program_chips(5, 4, 5, **COMPRESS_SMALL)[8],
            "FriFold": lambda: synth.fri_fold_program_chips()[2],
            "Poseidon2SkinnyDeg9": lambda: synth.skinny_program_chips(5)[1],
            # more chips of the core machine: misc/mov_cond, control_flow/jump, control_flow/branch, alu/sll
            "MovCond": lambda: synth.mov_cond_chip(5), "Jump": lambda: synth.jump_chip(5),
            "Branch": lambda: synth.branch_chip(6), "ShiftLeft": lambda: synth.shift_left_chip(6),
            "CloClz": lambda: synth.clo_clz_chip(5),
            # the 2^16-row byte table with the multiplicities of two small chips' lookups
            "Byte": lambda: synth.byte_chip_for([synth.bitwise_chip(4), synth.lt_chip(4)]),
            "Program": lambda: synth.program_chip(6), "SyscallCore": lambda: synth.syscall_chip(5, "Core"),
            "SyscallPrecompile": lambda: synth.syscall_chip(4, "Precompile"),
            "MemoryLocal": lambda: synth.memory_local_chip(4), "ShiftRight": lambda: synth.shift_right_chip(6),
            "Mul": lambda: synth.mul_chip(6), "Cpu": lambda: synth.core_program_chips(6)[0][0],
            "DivRem": lambda: synth.div_rem_chip(6)}[which]()
    lqd = chip.log_quotient_degree
    if which in ("wide1024", "wide4096"):
        assert ctx.air_info(chip.air)["num_kernels"] > 1, "this case must exercise the multi-part accumulate path"
    air = su.AIRS[chip.air]
    n = chip.log_degree
    _, main_pd = ctx.commit([chip.main], [MONTY_ONE], lqd)
    prep_pd = perm_pd = None
    chal = su.M(np.arange(10, 18).reshape(2, 4))
    alpha = su.M([3, 1, 4, 1])
    kw = {}
    okw = {}
    if chip.preprocessed is not None:
        _, prep_pd = ctx.commit([chip.preprocessed], [MONTY_ONE], lqd)
        kw["prep"] = (prep_pd, 0)
        okw["prep_q"] = _natural(prep_pd.lde(0))
    lcs = None
    if air.sends or air.receives:
        # device LogUp trace (zk_permutation_trace) against the numpy restatement of permutation.rs:102-196
        from oracle import logup
        p_c, m_c = chip.canon
        exp_tr, exp_lcs = logup.generate_permutation_trace(air, p_c, m_c, ob.from_monty(chal[0]), ob.from_monty(chal[1]))
        dptr, lcs = ctx.permutation_trace(chip.air, ctx.upload(chip.preprocessed) if chip.preprocessed is not None else 0,
                                          ctx.upload(chip.main), 1 << n, chal)
        tr = ctx.download(dptr, exp_tr.shape)
        assert (ob.from_monty(tr) == exp_tr).all()
        assert list(ob.from_monty(lcs)) == exp_lcs
        _, perm_pd = ctx.commit([tr], [MONTY_ONE], lqd)
        kw["perm"] = (perm_pd, 0)
        okw["perm_q"] = _natural(perm_pd.lde(0))
        okw["lcs"] = ob.from_monty(lcs)
    pvs = su.public_values_for([chip], 231 if which in ("PublicValues", "Cpu") else 8)
    gcs = su.M(np.arange(1, 15))
    dptr = ctx.quotient(chip.air, (main_pd, 0), n, lqd, alpha, perm_challenges=chal, public_values=pvs,
                        local_cumsum=lcs, global_cumsum=gcs, **kw)
    got = ob.from_monty(ctx.download(dptr, (1 << lqd, 1 << n, 4)))
    ctx.dev_free(dptr)
    exp = ae.quotient_values(air, n, lqd, _natural(main_pd.lde(0)), ob.from_monty(alpha), chal=ob.from_monty(chal),
                             gcs=ob.from_monty(gcs), pvs=ob.from_monty(pvs), **okw)
    # chunk c holds the rows i = c (mod 2^lqd)
    for c in range(1 << lqd):
        assert (got[c] == exp[c::1 << lqd]).all(), c
    # the quotient of a valid trace is a polynomial of degree < 2N: its top half of coefficients vanish, which
    # the FRI low-degree test checks in test_shard_proof_verifies; here: not identically zero
    assert exp.any()


NUM_PV = 8  # StarkMachine::num_pv_elts of the synthetic machine


def _machine(chips):
    return {c.name: c for c in chips}


def _prove(ctx, chips, log_blowup, nq, pw, pc_start=0x1234, mutate=None, num_pv=None):
    """setup + the machine-level challenger (pk.observe_into) + prove one shard; returns what the verifier needs"""
    NUM_PV = num_pv or globals()["NUM_PV"]
    prover = GpuShardProver(ctx, log_blowup, nq, pw, num_pv_elts=NUM_PV)
    pk = prover.setup(chips, pc_start=su.M([pc_start])[0], initial_global_cumulative_sum=su.M(np.arange(101, 115)))
    ch = Challenger(ctx)
    pk.observe_into(ch)                                  # machine.rs:79-86, once per proof; shards get a clone
    pvs = su.public_values_for(chips, NUM_PV)
    data = prover.commit(chips, pvs)
    if mutate:
        mutate(prover, pk, data)
    sp = prover.open(pk, data, Challenger(ctx, ch.w))
    return prover, pk, data, sp


@pytest.mark.parametrize("be", BACKENDS)
def test_shard_proof_verifies(be):
    """multi-chip shard (BASELINE config 1 stand-in) proven with the reference's transcript and checked by the
    verifier written from crates/stark/src/verifier.rs alone: Fibonacci (public values), a wide chip, a balanced LogUp
    pair (`local_only`, so their main traces open at zeta only), a Global-scope chip whose cumulative sum is read
    from its last 14 main columns, a `local_only` chip without lookups and a chip with a preprocessed trace."""
    ctx = _backend(be)
    nq, pw = (8, 6) if be == "emu" else (84, 16)
    send, recv = su.lookup_side_chips(5)
    chips = [su.fibonacci_chip(6), su.wide_chip(4, 64), send, recv, su.global_chip(3), su.local_bool_chip(4)]
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    assert [c.name for c in data.chips] == ["Fibonacci", "LookupRecv", "LookupSend", "LocalBool", "Wide64", "GlobalTail"]
    # commitment parity with the oracle PCS
    assert (data.main_commit == ob.pcs_commit([c.main for c in data.chips], 1).root).all()
    vk = su.vk_of(pk)
    ok, why = su.machine_verify(vk, _machine(chips), [sp], NUM_PV, 1, nq, pw)
    assert ok, why
    # every chip has a permutation matrix (width 0 without lookups) and the Global chip's sum is its last 14 columns
    gi = sp.chip_ordering["GlobalTail"]
    assert (sp.opened_values[gi].global_cumulative_sum == data.chips[gi].main.reshape(-1)[-14:]).all()
    assert all(v.permutation.local.shape[0] == 4 * su.AIRS[c.air].perm_width for v, c in zip(sp.opened_values, data.chips))
    li = sp.chip_ordering["LookupSend"]
    assert not sp.opened_values[li].main.next.any() and sp.opened_values[li].main.local.any()   # local_only: next = 0
    assert sp.opened_values[li].local_cumulative_sum.any() and not sp.local_cumulative_sum().any()
    # tampering with an opened value, a commitment, a cumulative sum or a public value must be rejected
    import copy
    for what in ("opened", "commit", "lcs", "pv"):
        bad = copy.deepcopy(sp)
        if what == "opened":
            bad.opened_values[0].main.local[0, 0] ^= 1
        elif what == "commit":
            bad.commitment.quotient_commit[3] ^= 1
        elif what == "lcs":
            bad.opened_values[li].local_cumulative_sum[1] ^= 1
        else:
            bad.public_values[5] ^= 1
        assert not su.machine_verify(vk, _machine(chips), [bad], NUM_PV, 1, nq, pw)[0], what
    # the wire format round-trips and the decoded proof verifies
    from zkmips_b200 import proof as pf
    blob = pf.to_bincode(sp)
    sp2 = pf.from_bincode(blob)
    assert pf.to_bincode(sp2) == blob
    ok, why = su.machine_verify(vk, _machine(chips), [sp2], NUM_PV, 1, nq, pw)
    assert ok, why
    data.main_data.free()
    pk.data and pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_shard_proof_bit_exact_with_oracle_prover(be):
    """north_star: "identical Merkle roots, quotient commitments, FRI query openings".  The CPU restatement of the
    reference prover (oracle/shard_prover.py: C oracle commits and Pcs::open, numpy LogUp and quotient) and the GPU
    path prove the same shard from the same challenger; the two `ShardProof`s must have the same bincode image --
    every commitment, opened value, cumulative sum, FRI layer root, the pow witness (both take the smallest) and every
    query opening and Merkle path."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (8, 6) if be == "emu" else (84, 16)
    send, recv = su.lookup_side_chips(5)
    chips = [su.fibonacci_chip(6), su.wide_chip(4, 64), send, recv, su.global_chip(3), su.local_bool_chip(4),
             su.lookup_chip(5)]
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    assert (opk.commit == pk.commit).all()
    och = bf.new_challenger()
    opk.observe_into(och)
    osp_proof = op.prove(opk, chips, och, su.public_values_for(chips, NUM_PV))
    assert (osp_proof.commitment.main_commit == sp.commitment.main_commit).all()
    assert (osp_proof.commitment.permutation_commit == sp.commitment.permutation_commit).all()
    assert (osp_proof.commitment.quotient_commit == sp.commitment.quotient_commit).all()
    assert osp_proof.opening_proof.pow_witness == sp.opening_proof.pow_witness
    assert pf.to_bincode(osp_proof) == pf.to_bincode(sp)
    data.main_data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_real_alu_chips_shard(be):
    """a shard of three REAL Ziren chips -- AddSub, Lt, Bitwise, transcribed from their Air::eval with their
    real byte-table sends and instruction-bus receives -- proven on the GPU path: the proof is byte-identical to the CPU
    restatement's, and the verifier accepts every per-chip check (PCS openings at [zeta] only: all three are
    local_only; constraint identity including the LogUp constraints).  Their lookups are answered by the Cpu and Byte
    chips, which are not in this shard, so the final shard-sum check must be the one that fails."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    logs = (7, 6, 5) if be == "emu" else (12, 11, 10)
    chips = [synth.add_sub_chip(logs[0]), synth.lt_chip(logs[1]), synth.bitwise_chip(logs[2]),
             synth.mov_cond_chip(logs[2]), synth.jump_chip(logs[2] - 1), synth.branch_chip(logs[1]),
             synth.shift_left_chip(logs[1] - 1), synth.clo_clz_chip(logs[2]), synth.shift_right_chip(logs[2]),
             synth.mul_chip(logs[2]), synth.div_rem_chip(logs[2] - 1)]
    if be != "emu":
        chips.append(synth.byte_chip_for(chips))          # the 2^16-row Byte table answering their byte lookups (GPU only: size)
    for c in chips:
        air = su.AIRS[c.air]
        assert c.main.shape[1] + 4 * air.perm_width + 8 == {"AddSub": 47, "Lt": 56, "Bitwise": 42, "MovCond": 48, "Jump": 82,
                                                            "Branch": 90, "ShiftLeft": 68, "CloClz": 46, "ShiftRight": 135, "Mul": 110, "DivRem": 162, "Byte": 54 - 12}[c.air]  # mips_costs.json
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 1, nq, pw)
    assert not ok and why.endswith("local cumulative sum is not zero"), why
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, chips, och, su.public_values_for(chips, NUM_PV))) == pf.to_bincode(sp)
    data.main_data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_recursion_chips_from_events_shard(be):
    """a compress-shaped shard of Poseidon2WideDeg3 (the recursion machine's widest chip) and the ALU
    chips, whose main traces never exist on the host: the prover gets EVENTS (16-word permutation inputs, 7-word
    AluEvents), fills the rows on the device (zk_tracegen_*) and commits them from HBM.  The proof must be byte-identical
    to the CPU prover's, which is given the rows the oracle fillers produce (those are checked against the reference's
    own C++ fillers in tests/test_tracegen.py)."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    logs = (5, 6, 4) if be == "emu" else (12, 13, 11)
    dev_chips = [synth.poseidon2_wide_chip(logs[0]), synth.add_sub_chip(logs[1], device=True),
                 synth.lt_chip(logs[2], device=True), su.fibonacci_chip(4)]
    assert all(c.main is None for c in dev_chips[:3])
    host_chips = [p2w_host_chip(logs[0], 3), synth.add_sub_chip(logs[1]), synth.lt_chip(logs[2]), su.fibonacci_chip(4)]
    prover, pk, data, sp = _prove(ctx, dev_chips, 1, nq, pw)
    assert len(data.device_traces) == 4
    assert (data.main_commit == ob.pcs_commit([c.main for c in prover.order(host_chips)], 1).root).all()
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(host_chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    assert (opk.commit == pk.commit).all()
    och = bf.new_challenger()
    opk.observe_into(och)
    want = op.prove(opk, host_chips, och, su.public_values_for(host_chips, NUM_PV))
    assert pf.to_bincode(want) == pf.to_bincode(sp)
    # every per-chip check of the verifier passes (the memory / byte / instruction buses are answered by chips that are
    # not in this shard, so only the final sum check can fail)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(host_chips), [sp], NUM_PV, 1, nq, pw)
    assert not ok and why.endswith("local cumulative sum is not zero"), why
    data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_recursion_program_shard_verifies_completely(be):
    """A toy recursion PROGRAM over six real RecursionAir chips -- MemoryConst (constants written to memory), BaseAlu,
    Select and Poseidon2WideDeg3 (reading them), MemoryVar (extension operands) and ExtAlu (reading those) -- whose
    memory bus balances: the verifier restatement accepts the proof
    COMPLETELY (PCS openings, constraint identity of every chip including LogUp, and the zero shard sum), and it is
    byte-identical with the CPU prover's.  Dropping one read from the program unbalances the bus and is rejected."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    logs = (5, 4, 5) if be == "emu" else (12, 11, 11)
    more = dict(log_var=6, log_ext=4, log_sel=5) if be == "emu" else dict(log_var=13, log_ext=11, log_sel=12)
    chips = synth.recursion_program_chips(*logs, **more) + [su.fibonacci_chip(4)]
    assert [c.air for c in chips[:6]] == ["MemoryConst", "BaseAlu", "Poseidon2WideDeg3", "Select", "MemoryVar", "ExtAlu"]
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    host = list(chips)
    host[2] = p2w_host_chip(logs[0], 3)                                # same events (same seed / fill), rows from the oracle
    host[2].preprocessed = chips[2].preprocessed
    host[2].main = ob.poseidon2_wide_trace(chips[2].events, chips[2].rows, True)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(host), [sp], NUM_PV, 1, nq, pw)
    assert ok, why
    assert not sp.local_cumulative_sum().any()
    assert all(sp.opened_values[sp.chip_ordering[c.name]].local_cumulative_sum.any() for c in chips[:6])
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(host, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, host, och, su.public_values_for(host, NUM_PV))) == pf.to_bincode(sp)
    data.free()
    pk.data.free()
    # one multiplicity off by one: every per-chip check still passes, the shard sum does not
    bad = synth.recursion_program_chips(*logs, **more) + [su.fibonacci_chip(4)]
    pre = ob.from_monty(bad[0].preprocessed).astype(np.uint64)
    pre[0, 5] = (pre[0, 5] + 1) % ae.P
    bad[0].preprocessed = ob.to_monty(pre.astype(np.uint32))
    prover, pk, data, sp = _prove(ctx, bad, 1, nq, pw)
    host[0] = bad[0]
    ok, why = su.machine_verify(su.vk_of(pk), _machine(host), [sp], NUM_PV, 1, nq, pw)
    assert not ok and why.endswith("local cumulative sum is not zero"), why
    data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_compress_machine_shard_verifies_completely(be):
    """The toy program on ALL NINE chips of the reference's compress machine (recursion/core/src/machine.rs:112-128:
    MemoryConst, MemoryVar, BaseAlu, ExtAlu, Poseidon2Wide, BatchFRI, Select, ExpReverseBitsLen, PublicValues), with
    the machine's 231 public values (PROOF_MAX_NUM_PVS, stark/src/types.rs:73) observed and the digest constrained by
    PublicValues: BatchFRI accumulators feed ExtAlu, ExpReverseBitsLen results feed BaseAlu, the memory bus balances.
    Complete verification, byte-identical with the CPU prover; a wrong digest in the public values is rejected by
    PublicValues' constraint alone."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    logs = (5, 4, 5) if be == "emu" else (12, 11, 11)
    more = COMPRESS_SMALL if be == "emu" else dict(log_var=14, log_ext=11, log_sel=12, log_bf=12, log_exp=11, pv=True)
    chips = synth.recursion_program_chips(*logs, **more)
    assert sorted(c.air for c in chips) == sorted(["MemoryConst", "MemoryVar", "BaseAlu", "ExtAlu", "Poseidon2WideDeg3",
                                                   "BatchFRI", "Select", "ExpReverseBitsLen", "PublicValues"])
    npv = 231
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw, num_pv=npv)
    host = list(chips)
    host[2] = p2w_host_chip(logs[0], 3)
    host[2].preprocessed = chips[2].preprocessed
    host[2].main = ob.poseidon2_wide_trace(chips[2].events, chips[2].rows, True)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(host), [sp], npv, 1, nq, pw)
    assert ok, why
    assert not sp.local_cumulative_sum().any()
    assert all(sp.opened_values[sp.chip_ordering[c.name]].local_cumulative_sum.any() for c in chips)
    # chips with transition constraints are opened at zeta and zeta * g; the local_only ones at zeta alone, their `next`
    # row is the zero vector the reference puts there (prover.rs:576-600)
    for c in chips:
        v = sp.opened_values[sp.chip_ordering[c.name]]
        assert (not np.asarray(v.preprocessed.next).any()) == su.AIRS[c.air].local_only, c.name
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=npv)
    opk = op.setup(host, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, host, och, su.public_values_for(host, npv))) == pf.to_bincode(sp)
    # a public-values vector whose digest differs from what the program committed: same prover, rejected
    bad_pvs = su.public_values_for(chips, npv).copy()
    bad_pvs[225] = su.M([5])[0]
    sp2 = prover.open(pk, prover.commit(chips, bad_pvs), _machine_challenger(ctx, pk))
    ok, why = su.machine_verify(su.vk_of(pk), _machine(host), [sp2], npv, 1, nq, pw)
    assert not ok, "a wrong digest must violate PublicValues' constraint"
    data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_fri_fold_program_shard_verifies_completely(be):
    """FriFold, the chip machine_wide_with_all_chips has beyond the compress machine (machine.rs:68-87), on a toy program
    with MemoryConst and MemoryVar: balanced bus, complete verification, byte-identical with the CPU prover."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    chips = synth.fri_fold_program_chips() if be == "emu" else synth.fri_fold_program_chips(14, 17, 13)
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 1, nq, pw)
    assert ok, why
    assert not sp.local_cumulative_sum().any()
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, chips, och, su.public_values_for(chips, NUM_PV))) == pf.to_bincode(sp)
    data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_poseidon2_skinny_program_shard_verifies_completely(be):
    """Poseidon2SkinnyDeg9 -- the wrap machine's Poseidon2 chip (machine.rs:138-153), log_quotient_degree 3, eleven rows
    per permutation -- with MemoryConst (log_quotient_degree 1) in one shard: balanced bus, complete verification,
    byte-identical with the CPU prover.  The prover gets the permutation INPUTS and fills the rows on the device
    (zk_tracegen_poseidon2_skinny); the verifier and the CPU prover get the rows of the numpy filler, which is pinned
    against the reference's poseidon2_skinny.hpp (tests/test_tracegen.py)."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    size = () if be == "emu" else (15, 11)
    chips = synth.skinny_program_chips(*size)
    dev_chips = synth.skinny_program_chips(*size, device=True)
    assert dev_chips[1].main is None and dev_chips[1].events.shape[1] == 16
    prover, pk, data, sp = _prove(ctx, dev_chips, 3, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 3, nq, pw)
    assert ok, why
    assert not sp.local_cumulative_sum().any()
    op = osp.OracleShardProver(su.AIRS, 3, nq, pw, num_pv_elts=NUM_PV)
    opk = op.setup(chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, chips, och, su.public_values_for(chips, NUM_PV))) == pf.to_bincode(sp)
    data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_core_program_shard(be):
    """The toy core-machine program on fifteen real chips -- Cpu, Program, AddSub, Bitwise, Lt, ShiftLeft, ShiftRight,
    CloClz, Mul, DivRem, MovCond, Jump, Branch, MemoryLocal, Byte (GPU only: 2^16 rows) -- with the core machine's 231 public values (start_pc, next_pc,
    execution_shard constrained by the CPU chip): the proof is byte-identical with the CPU prover's and every per-chip
    check of the verifier passes; the memory, program, instruction and byte buses cancel (tests/test_air_ir.py), so what
    is left in the shard's cumulative sum are MemoryLocal's Global-kind forwards."""
    from oracle import binding_fri as bf
    from oracle import shard_prover as osp
    from zkmips_b200 import proof as pf
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    chips, _ = synth.core_program_chips(7 if be == "emu" else 12)
    dev_chips, _ = synth.core_program_chips(7 if be == "emu" else 12, device=True)   # the CPU chip as events (zk_tracegen_cpu)
    assert dev_chips[0].main is None and dev_chips[0].events.shape[1] == 22
    if be == "emu":
        chips, dev_chips = chips[:-1], dev_chips[:-1]
    npv = 231
    prover, pk, data, sp = _prove(ctx, dev_chips, 1, nq, pw, num_pv=npv)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], npv, 1, nq, pw)
    assert not ok and why.endswith("local cumulative sum is not zero"), why
    op = osp.OracleShardProver(su.AIRS, 1, nq, pw, num_pv_elts=npv)
    opk = op.setup(chips, pc_start=pk.pc_start, initial_global_cumulative_sum=pk.initial_global_cumulative_sum)
    och = bf.new_challenger()
    opk.observe_into(och)
    assert pf.to_bincode(op.prove(opk, chips, och, su.public_values_for(chips, npv))) == pf.to_bincode(sp)
    # wrong public next_pc: the CPU chip's own constraint fails
    bad_pvs = su.public_values_for(chips, npv).copy()
    bad_pvs[41] = su.M([7])[0]
    sp2 = prover.open(pk, prover.commit(dev_chips, bad_pvs), _machine_challenger(ctx, pk))
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp2], npv, 1, nq, pw)
    assert not ok and not why.endswith("local cumulative sum is not zero"), why
    data.free()
    pk.data.free()


def _machine_challenger(ctx, pk):
    ch = Challenger(ctx)
    pk.observe_into(ch)
    return Challenger(ctx, ch.w)


@pytest.mark.parametrize("drop", ["public_values", "vk", "local_sum", "global_sum", "perm_commit", "local_only"])
@pytest.mark.parametrize("be", BACKENDS)
def test_dropped_observation_fails_verification(be, drop, monkeypatch):
    """VERDICT round 1, task 2: a prover that skips any one transcript step of prover.rs:322,406-413 -- or opens a
    `local_only` chip at two points -- must produce a proof the (independent) verifier rejects."""
    ctx = _backend(be)
    nq, pw = 6, 4
    send, recv = su.lookup_side_chips(4)
    chips = [su.fibonacci_chip(5), send, recv, su.global_chip(3)]
    prover = GpuShardProver(ctx, 1, nq, pw, num_pv_elts=NUM_PV)
    pk = prover.setup(chips, pc_start=su.M([77])[0])
    ch = Challenger(ctx)
    if drop != "vk":
        pk.observe_into(ch)
    pvs = su.public_values_for(chips, NUM_PV)
    data = prover.commit(chips, pvs)
    if drop == "public_values":
        prover.num_pv_elts = 0                       # skips prover.rs:322
    if drop == "local_only":
        for c in data.chips:
            c.local_only = False                     # opens [zeta, zeta*g] where the verifier expects [zeta]
    real_many = Challenger.observe_many
    seen = {"n": 0}

    def observe_many(self, parts):
        parts = list(parts)
        seen["n"] += 1
        # inside open(): call 1 = [public values, main commit]; call 2 = [permutation commit, then per chip local sum,
        # global x, global y]
        if seen["n"] == 2 and drop in ("perm_commit", "local_sum", "global_sum"):
            parts.pop({"perm_commit": 0, "local_sum": 1, "global_sum": 2}[drop])
        real_many(self, parts)

    monkeypatch.setattr(Challenger, "observe_many", observe_many)
    sp = prover.open(pk, data, Challenger(ctx, ch.w))
    monkeypatch.setattr(Challenger, "observe_many", real_many)
    if drop == "local_only":
        for c in chips:
            c.local_only = su.AIRS[c.air].local_only
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 1, nq, pw)
    assert not ok, f"dropping {drop} went unnoticed"
    data.main_data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_unbalanced_lookups_are_rejected(be):
    """lookup_pair's sends and receives do not cancel: every per-chip check passes (PCS, constraint identity with the
    permutation constraints) and the verifier's LAST check -- the shard's local cumulative sum is zero,
    verifier.rs:236-244 -- rejects the proof, as the reference would."""
    ctx = _backend(be)
    nq, pw = (6, 4) if be == "emu" else (84, 16)
    chips = [su.fibonacci_chip(5), su.lookup_chip(4)]
    prover, pk, data, sp = _prove(ctx, chips, 1, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 1, nq, pw)
    assert not ok and why.endswith("local cumulative sum is not zero"), why
    data.main_data.free()
    pk.data.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_invalid_trace_is_rejected(be):
    """a trace that violates its AIR still yields well-formed commitments and a FRI proof (any N values
    interpolate to a degree < N chunk), but the verifier's constraint identity at zeta fails."""
    ctx = _backend(be)
    chip = su.fibonacci_chip(5)
    chip.main = chip.main.copy()
    chip.main[7, 1] = (int(chip.main[7, 1]) + 1) % su.P
    prover, pk, data, sp = _prove(ctx, [chip], 1, 4, 4)
    ok, why = su.machine_verify(su.vk_of(pk), _machine([chip]), [sp], NUM_PV, 1, 4, 4)
    assert not ok and "OodEvaluationMismatch" in why
    data.main_data.free()


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
    provers = [GpuShardProver(c, 1, 20, 8, num_pv_elts=NUM_PV) for c in ctxs]
    pks = [p.setup(chips) for p in provers]
    pvs = su.public_values_for(chips, NUM_PV)
    from zkmips_b200 import proof as pf

    def prove(k):
        ch = Challenger(ctxs[k])
        pks[k].observe_into(ch)
        data = provers[k].commit(chips, pvs)
        sp = provers[k].open(pks[k], data, ch)
        data.main_data.free()
        return pf.to_bincode(sp)

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
        for blob in out[k]:
            assert blob == ref
    for k in range(2):
        pks[k].data.free()
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
    chips = [chip, su.fibonacci_chip(4)]
    prover, pk, data, sp = _prove(ctx, chips, 2, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk), _machine(chips), [sp], NUM_PV, 2, nq, pw)
    assert ok, why
    data.main_data.free()


@pytest.mark.parametrize("name,log_blowup,num_queries", [("default", 1, 84), ("compressed", 2, 42), ("ultra_compressed", 3, 28)])
@pytest.mark.parametrize("be", BACKENDS)
def test_reference_fri_configs(be, name, log_blowup, num_queries):
    """The three FRI configurations of the reference (crates/stark/src/kb31_poseidon2.rs:203-241: default_fri_config
    for core shards, compressed_fri_config for compress, ultra_compressed_fri_config for shrink/wrap; 16 proof-of-work
    bits each): a shard with chips of different heights, lookups and a preprocessed trace proves and verifies under
    every one of them, and the main commitment equals the oracle's for that blowup."""
    ctx = _backend(be)
    nq, pw = (min(num_queries, 6), 5) if be == "emu" else (num_queries, 16)
    send, recv = su.lookup_side_chips(5)
    chips = [su.fibonacci_chip(6), su.wide_chip(4, 64), send, recv, su.lookup_chip(3)]
    prover, pk, data, sp = _prove(ctx, chips, log_blowup, nq, pw)
    assert (data.main_commit == ob.pcs_commit([c.main for c in data.chips], log_blowup).root).all()
    vk = su.vk_of(pk)
    ok, why = su.machine_verify(vk, _machine(chips), [sp], NUM_PV, log_blowup, nq, pw)
    # lookup_pair is unbalanced on purpose: everything up to the last check (shard sum == 0) must pass
    assert not ok and why.endswith("local cumulative sum is not zero"), f"{name}: {why}"
    chips2 = chips[:4]
    prover, pk2, data2, sp2 = _prove(ctx, chips2, log_blowup, nq, pw)
    ok, why = su.machine_verify(su.vk_of(pk2), _machine(chips2), [sp2], NUM_PV, log_blowup, nq, pw)
    assert ok, f"{name}: {why}"
    # a proof made for one configuration must not verify under another blowup
    other = 1 if log_blowup != 1 else 2
    assert not su.machine_verify(su.vk_of(pk2), _machine(chips2), [sp2], NUM_PV, other, nq, pw)[0]
    data.main_data.free()
    data2.main_data.free()
    pk.data.free()
