"""The constraint IR: JSON round trip of the real chips (the Rust-side exporter's target format) and the transcribed
chips' shapes against the reference's cost table."""
import json

import numpy as np
import os

from zkmips_b200.air import codegen, library
from zkmips_b200.air.ir import Air

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPORTED = os.path.join(ROOT, "zkmips_b200", "air", "exported")


def test_exported_json_reproduces_the_generated_kernels():
    """Loading zkmips_b200/air/exported/<Chip>.json (what a recording-builder exporter would write) must give the same
    program as the hand transcription: identical JSON again, and byte-identical generated CUDA."""
    for make in (library.add_sub, library.lt, library.bitwise, lambda: library.poseidon2_wide(3),
                 lambda: library.poseidon2_wide(9), library.memory_const, library.base_alu, library.memory_var,
                 library.ext_alu, library.select, library.batch_fri, library.exp_reverse_bits_len,
                 library.public_values_chip, library.fri_fold, library.poseidon2_skinny, library.mov_cond, library.jump,
                 library.branch, library.shift_left, library.clo_clz, library.byte_chip, library.program_chip,
                 lambda: library.syscall_chip("Core"), lambda: library.syscall_chip("Precompile"), library.memory_local,
                 library.shift_right, library.mul, library.cpu, library.div_rem):
        air = make()
        text = open(os.path.join(EXPORTED, air.name + ".json")).read()
        assert json.loads(text) == json.loads(air.to_json()), f"{air.name}.json is stale: run tools/export_airs.py"
        loaded = Air.from_json(text)
        assert loaded.to_json() == air.to_json()
        assert loaded.commit_scope == air.commit_scope and loaded.local_only == air.local_only
        assert codegen.generate([loaded]) == codegen.generate([air])


def test_real_chip_shapes_match_mips_costs():
    """committed columns per row = main + 4 * permutation + 4 * quotient chunks (Chip::cost, crates/stark/src/chip.rs:
    151-162) must equal crates/core/executor/src/artifacts/mips_costs.json: AddSub 47, Lt 56, Bitwise 42; constraint
    counts as StarkMachine::setup computes them (own + count_permutation_constraints, permutation.rs:355-388)."""
    want = {"AddSub": (19, 47, 14, 8), "Lt": (36, 56, 32, 4), "Bitwise": (18, 42, 5, 5), "MovCond": (32, 48, 43, 1),
            "Jump": (66, 82, 60, 2), "Branch": (62, 90, 60, 8), "ShiftLeft": (44, 68, 64, 5), "CloClz": (22, 46, 20, 5),
            "ShiftRight": (71, 135, 83, 26), "Mul": (58, 110, 41, 19), "DivRem": (106, 162, 126, 21)}
    for make in (library.add_sub, library.lt, library.bitwise, library.mov_cond, library.jump, library.branch,
                 library.shift_left, library.clo_clz, library.shift_right, library.mul, library.div_rem):
        air = make()
        width, cost, own, n_lookups = want[air.name]
        assert air.main_width == width
        assert air.main_width + 4 * air.perm_width + 4 * 2 == cost
        assert len(air.sends) + len(air.receives) == n_lookups
        assert air.perm_width == -(-n_lookups // 2) + 1
        assert air.num_constraints == own + (air.perm_width - 1) + 3
        assert air.local_only == (air.name not in ("CloClz", "ShiftRight")) and air.commit_scope == "local"   # no local_only() there
        assert air.max_degree() == 3  # log_quotient_degree 1


def _constraints_on_trace(air, main, prep=None, seed=5, pvs=()):
    """every constraint of `air` (own + LogUp) on the trace domain itself: rows local / next, selectors as indicator
    vectors, permutation trace from the numpy LogUp restatement.  All must vanish on a valid trace."""
    import numpy as np
    from oracle import air_eval as ae, logup
    rng = np.random.default_rng(seed)
    n = main.shape[0]
    chal = [[int(x) for x in rng.integers(1, ae.P, 4)] for _ in range(2)]
    perm, lcs = logup.generate_permutation_trace(air, prep, main, chal[0], chal[1])
    nxt = (np.arange(n) + 1) % n
    rows = {"main": (main, main[nxt]), "perm": (perm, perm[nxt])}
    if prep is not None:
        rows["prep"] = (prep, prep[nxt])
    first, last = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    first[0], last[n - 1] = 1, 1
    sel = {"first": first, "last": last, "trans": 1 - last}
    out = []
    for c in ae.eval_rows(air, rows, sel, chal, lcs, (0,) * 14, tuple(int(x) for x in pvs)):
        out.append(np.stack(c.c) if isinstance(c, ae.VExt) else np.asarray(c))
    return out


def test_poseidon2_wide_air_vanishes_on_the_reference_fillers_rows():
    """Poseidon2WideChip<3> / <9> as transcribed (library.poseidon2_wide) against rows produced by the oracle filler,
    which is itself checked against the reference's C++ filler (tests/test_tracegen.py): a valid trace satisfies every
    constraint, a corrupted cell breaks at least one.  Shapes: 313 / 172 main columns, 49 preprocessed columns
    (chips/poseidon2_wide/columns), 32 memory lookups."""
    import numpy as np
    from oracle import binding as ob
    from zkmips_b200.proof import to_monty
    rng = np.random.default_rng(11)
    n_ev, rows = 13, 16
    x = to_monty(rng.integers(0, ae_P, (n_ev, 16), dtype=np.uint64))
    instrs = to_monty(rng.integers(0, ae_P, (n_ev, 48), dtype=np.uint64))
    prep = ob.from_monty(ob.poseidon2_wide_prep(instrs, rows)).astype(np.uint64)
    for degree, width, own in ((3, 313, 298), (9, 172, 157)):
        air = library.poseidon2_wide(degree)
        assert (air.main_width, air.prep_width, len(air.sends), len(air.receives)) == (width, 49, 32, 0)
        assert air.max_degree() == degree and air.local_only
        bs = 2 if degree == 3 else 8
        assert air.perm_width == 32 // bs + 1 and air.num_constraints == own + (air.perm_width - 1) + 3
        main = ob.from_monty(ob.poseidon2_wide_trace(x, rows, degree == 3)).astype(np.uint64)
        vals = _constraints_on_trace(air, main, prep)
        assert len(vals) == air.num_constraints and all(not v.any() for v in vals)
        bad = main.copy()
        bad[3, 130] = (bad[3, 130] + 1) % ae_P           # one cell of internal_rounds_state
        assert any(v.any() for v in _constraints_on_trace(air, bad, prep)[:own])


def test_alu_airs_vanish_on_their_fillers_rows():
    from zkmips_b200 import synth
    import numpy as np
    for make, events, rows in ((library.add_sub, synth.add_sub_events, synth.add_sub_rows),
                               (library.lt, synth.lt_events, synth.lt_rows),
                               (library.bitwise, synth.bitwise_events, synth.bitwise_rows),
                               (library.mov_cond, synth.mov_cond_events, synth.mov_cond_rows),
                               (library.jump, synth.jump_events, synth.jump_rows),
                               (library.branch, synth.branch_events, synth.branch_rows),
                               (library.shift_left, synth.shift_left_events, synth.shift_left_rows),
                               (library.clo_clz, synth.clo_clz_events, synth.clo_clz_rows),
                               (library.shift_right, synth.shift_right_events, synth.shift_right_rows),
                               (library.mul, synth.mul_events, synth.mul_rows),
                               (library.div_rem, synth.div_rem_events, synth.div_rem_rows)):
        ev, n = events(5)
        vals = _constraints_on_trace(make(), rows(ev, n))
        assert all(not v.any() for v in vals), make.__name__
    # one wrong cell each: MovCond's is_zero result, Jump's link address, Branch's decision, ShiftLeft's carry
    for make, events, rows, cell in ((library.mov_cond, synth.mov_cond_events, synth.mov_cond_rows, (2, 28)),
                                     (library.jump, synth.jump_events, synth.jump_rows, (3, 37)),
                                     (library.branch, synth.branch_events, synth.branch_rows, (1, 59)),
                                     (library.shift_left, synth.shift_left_events, synth.shift_left_rows, (2, 35)),
                                     (library.clo_clz, synth.clo_clz_events, synth.clo_clz_rows, (2, 15)),
                                     (library.shift_right, synth.shift_right_events, synth.shift_right_rows, (3, 43)),
                                     (library.mul, synth.mul_events, synth.mul_rows, (2, 20)),
                                     (library.div_rem, synth.div_rem_events, synth.div_rem_rows, (2, 38))):
        ev, n = events(5)
        bad = rows(ev, n)
        bad[cell] = (bad[cell] + 1) % ae_P
        assert any(v.any() for v in _constraints_on_trace(make(), bad)), make.__name__
    # ShiftLeft's first rows are the (a, b, c) triples of the reference's own test (alu/sll/mod.rs prove_koalabear)
    ev, n = synth.shift_left_events(5)
    assert [(int(e[4]), int(e[5]), int(e[6])) for e in ev[:19]] == synth.SLL_REFERENCE_CASES
    ev, n = synth.clo_clz_events(5)                                    # and CloClz's (alu/clo_clz/mod.rs prove_koalabear)
    assert [(int(e[2]), int(e[4]), int(e[5])) for e in ev[:6]] == synth.CLOCLZ_REFERENCE_CASES


ae_P = 0x7F000001


def test_recursion_program_chips_satisfy_their_airs():
    """MemoryConst (1 + 12 columns, 2 block writes per row), BaseAlu (12 + 32 columns, 4 operations per row: 5
    constraints, 2 reads and 1 write each), MemoryVar (8 + 4), ExtAlu (48 + 32: 17 constraints per operation, block
    reads / writes) and Select (5 + 8) as transcribed, on the toy program of synth.recursion_program_chips"""
    from zkmips_b200 import synth
    mem, alu, _, sel, var, ext = synth.recursion_program_chips(5, 4, 5, log_var=6, log_ext=4, log_sel=5)
    for chip, air, shape in ((mem, library.memory_const(), (1, 12, 2, 0, 0)), (alu, library.base_alu(), (12, 32, 4, 8, 20)),
                             (var, library.memory_var(), (8, 4, 2, 0, 0)), (ext, library.ext_alu(), (48, 32, 4, 8, 68)),
                             (sel, library.select(), (5, 8, 2, 3, 2))):
        assert (air.main_width, air.prep_width, len(air.sends), len(air.receives)) == shape[:4]
        assert air.num_constraints == shape[4] + ((air.perm_width - 1) + 3 if air.perm_width else 0)
        vals = _constraints_on_trace(air, chip.canon[1], chip.canon[0])
        assert all(not v.any() for v in vals)
    bad = alu.canon[1].copy()
    bad[0, 0] = (bad[0, 0] + 1) % ae_P                      # a wrong result
    assert any(v.any() for v in _constraints_on_trace(library.base_alu(), bad, alu.canon[0])[:20])


def test_compress_machine_chips_satisfy_their_airs():
    """The three chips that complete the reference's compress machine (recursion/core/src/machine.rs:112-128) beside the
    six above: BatchFRI (13 + 6 columns: dummy + 12 accumulator constraints, 3 reads, 1 write), ExpReverseBitsLen
    (7 + 10: 8 constraints, 3 sends with signed multiplicities) and PublicValues (1 + 10, 231 public values, digest at
    223..231), on the toy program of synth.recursion_program_chips; all three read the NEXT row or none of the
    `local_only` chips' shortcuts.  A corrupted cell breaks a constraint of each."""
    from zkmips_b200 import synth
    chips = synth.recursion_program_chips(5, 4, 5, log_var=7, log_ext=4, log_sel=5, log_bf=5, log_exp=5, pv=True)
    assert [c.air for c in chips] == ["MemoryConst", "BaseAlu", "Poseidon2WideDeg3", "Select", "MemoryVar", "ExtAlu",
                                      "BatchFRI", "ExpReverseBitsLen", "PublicValues"]
    pvs = np.zeros(231, np.uint64)
    pvs[223:231] = chips[8].pv_digest
    for chip, air, shape, cell in ((chips[6], library.batch_fri(3), (13, 6, 1, 3, 13), (2, 1)),
                                   (chips[7], library.exp_reverse_bits_len(3), (7, 10, 3, 0, 8), (3, 4)),
                                   (chips[8], library.public_values_chip(), (1, 10, 1, 0, 8), (5, 0))):
        assert (air.main_width, air.prep_width, len(air.sends), len(air.receives)) == shape[:4]
        assert air.num_constraints == shape[4] + (air.perm_width - 1) + 3
        assert not air.local_only and air.max_degree() <= 3
        assert all(not v.any() for v in _constraints_on_trace(air, chip.canon[1], chip.canon[0], pvs=pvs))
        bad = chip.canon[1].copy()
        bad[cell] = (bad[cell] + 1) % ae_P
        assert any(v.any() for v in _constraints_on_trace(air, bad, chip.canon[0], pvs=pvs)[:shape[4]]), air.name
    # FriFold (33 + 20 columns, nine signed-multiplicity sends, dummy + 17 constraints): the tenth chip of
    # machine_wide_with_all_chips (machine.rs:68-87), on its own toy program
    mem, var, ff = synth.fri_fold_program_chips()
    air = library.fri_fold(3)
    assert (air.main_width, air.prep_width, len(air.sends), len(air.receives), air.num_constraints) == (33, 20, 9, 0, 18 + 5 + 3)
    assert all(not v.any() for v in _constraints_on_trace(air, ff.canon[1], ff.canon[0]))
    bad = ff.canon[1].copy()
    bad[1, 30] = (bad[1, 30] + 1) % ae_P                               # ro_output
    assert any(v.any() for v in _constraints_on_trace(air, bad, ff.canon[0])[:18])
    for chip, a in ((mem, library.memory_const()), (var, library.memory_var())):
        assert all(not v.any() for v in _constraints_on_trace(a, chip.canon[1], chip.canon[0]))
    # Poseidon2SkinnyDeg9, the Poseidon2 chip of the wrap machine (machine.rs:138-153): 28 + 51 columns, eleven rows per
    # permutation tied by next-row constraints, dummy + 60 constraints of degree <= 9, LogUp batches of 8; its rows are
    # pinned against the reference's poseidon2_skinny.hpp in tests/test_tracegen.py
    mem, sk = synth.skinny_program_chips()
    air = library.poseidon2_skinny(9)
    assert (air.main_width, air.prep_width, len(air.sends), len(air.receives)) == (28, 51, 16, 0)
    assert air.num_constraints == 61 + 2 + 3 and air.max_degree() == 9 and air.perm_width == 3 and not air.local_only
    assert all(not v.any() for v in _constraints_on_trace(air, sk.canon[1], sk.canon[0]))
    for cell in ((5, 20), (3, 7), (0, 2)):                              # an s0 column, an external round, the input row
        bad = sk.canon[1].copy()
        bad[cell] = (bad[cell] + 1) % ae_P
        assert any(v.any() for v in _constraints_on_trace(air, bad, sk.canon[0])[:61]), cell
    # the six older chips are unchanged by the additions (the accumulators / results only replace operands)
    for chip, air in ((chips[0], library.memory_const()), (chips[1], library.base_alu()), (chips[4], library.memory_var()),
                      (chips[5], library.ext_alu())):
        assert all(not v.any() for v in _constraints_on_trace(air, chip.canon[1], chip.canon[0]))


def test_byte_chip_answers_the_core_chips_byte_lookups():
    """ByteChip (bytes/air.rs:22-74: 12 preprocessed + 10 multiplicity columns, ten receives, cost 54) with the
    multiplicities ByteChip::generate_trace would count (synth.byte_chip_for: every byte lookup the eleven transcribed core
    chips send is looked up in the table and CHECKED against it).  The byte bus then balances: with the lookups of the
    other kinds removed, the LogUp cumulative sums of the nine chips add up to zero for random challenges -- and no
    longer do when one multiplicity is off by one."""
    import copy
    from oracle import logup
    from zkmips_b200 import synth
    chips = [synth.add_sub_chip(6), synth.lt_chip(6), synth.bitwise_chip(5), synth.mov_cond_chip(5), synth.jump_chip(5),
             synth.branch_chip(6), synth.shift_left_chip(5), synth.clo_clz_chip(5), synth.shift_right_chip(6),
             synth.mul_chip(6), synth.div_rem_chip(6)]
    byte = synth.byte_chip_for(chips)
    air = library.byte_chip()
    assert (air.main_width, air.prep_width, len(air.sends), len(air.receives), air.num_constraints) == (10, 12, 0, 10, 5 + 3)
    assert air.prep_width + air.main_width + 4 * air.perm_width + 8 == 54          # Chip::cost counts preprocessed columns
    assert byte.canon[1].sum() > 500 and byte.canon[1][:, [0, 1, 2, 4, 5, 6, 7, 8, 9]].any(axis=0).all()
    assert all(not v.any() for v in _constraints_on_trace(air, byte.canon[1], byte.canon[0]))
    # ProgramChip: the other table chip of the core machine (14 preprocessed + 1 multiplicity column, cost 31)
    pair = library.program_chip()
    prog = synth.program_chip(5)
    assert (pair.main_width, pair.prep_width, len(pair.receives), pair.num_constraints) == (1, 14, 1, 1 + 3)
    assert pair.prep_width + pair.main_width + 4 * pair.perm_width + 8 == 31
    assert all(not v.any() for v in _constraints_on_trace(pair, prog.canon[1], prog.canon[0]))
    # SyscallCore / SyscallPrecompile (6 columns, one constraint, a syscall lookup and a Global-table send each, cost 22)
    for kind in ("Core", "Precompile"):
        sc, ch = library.syscall_chip(kind), synth.syscall_chip(5, kind)
        assert (sc.main_width, len(sc.sends), len(sc.receives), sc.num_constraints) == (6, 1 + (kind != "Core"), kind == "Core", 1 + 4)
        assert sc.main_width + 4 * sc.perm_width + 8 == 22 and not sc.local_only
        assert all(not v.any() for v in _constraints_on_trace(sc, ch.canon[1]))
    # MemoryLocal (four cells of 14 columns per row, one constraint and four lookups per cell, cost 100)
    ml, ch = library.memory_local(), synth.memory_local_chip(4)
    assert (ml.main_width, len(ml.sends), len(ml.receives), ml.num_constraints) == (56, 12, 4, 4 + 8 + 3)
    assert ml.main_width + 4 * ml.perm_width + 8 == 100 and not ml.local_only
    assert all(not v.any() for v in _constraints_on_trace(ml, ch.canon[1]))
    rng = np.random.default_rng(3)
    alpha, beta = ([int(x) for x in rng.integers(1, ae_P, 4)] for _ in range(2))

    def byte_bus_sum(all_chips):
        total = np.zeros(4, np.uint64)
        for ch in all_chips:
            only = copy.copy({a.name: a for a in [getattr(library, f)() for f in
                                                  ("add_sub", "lt", "bitwise", "mov_cond", "jump", "branch", "shift_left",
                                                   "clo_clz", "shift_right", "mul", "div_rem", "byte_chip")]}[ch.air])
            only.sends = [l for l in only.sends if l["kind"] == 4]
            only.receives = [l for l in only.receives if l["kind"] == 4]
            _, lcs = logup.generate_permutation_trace(only, ch.canon[0], ch.canon[1], alpha, beta)
            total = (total + np.asarray(lcs, np.uint64)) % ae_P
        return total
    assert not byte_bus_sum(chips + [byte]).any()
    assert byte_bus_sum(chips).any()
    byte.canon[1][np.nonzero(byte.canon[1][:, 4])[0][0], 4] += 1
    assert byte_bus_sum(chips + [byte]).any()


def test_core_program_chips_interlock():
    """A toy core-machine program (synth.core_program_chips: ALU instructions, conditional moves, branches with their delay
    slots, jumps, and MULT / DIV / MOD with the HI register over registers 1..31, executed in Python) on FIFTEEN real chips:
    Cpu (cpu/air/mod.rs: 67 columns, 57 constraints, 19 lookups, cost 119), Program, AddSub, Bitwise, Lt, ShiftLeft,
    ShiftRight, CloClz, Mul, DivRem, MovCond, Jump, Branch, MemoryLocal and Byte.  Every constraint of every chip vanishes on the rows, and the chips interlock as in the
    reference's machine: taken kind by kind, the LogUp sums of the MEMORY bus (register accesses chained from
    MemoryLocal's initial to its final state), the PROGRAM bus (instruction fetches), the INSTRUCTION bus (CPU -> the chip
    implementing the opcode, with shard and clk for the instructions that write HI, plus the chips' own dependencies:
    CloClz's SRL on ShiftRight, Branch's two SLT on Lt and its target ADD on AddSub, JumpDirect's ADD, DivRem's product
    on Mul, its absolute values on AddSub and its remainder check on Lt) and the BYTE bus cancel across the shard; only MemoryLocal's Global-kind
    forwards have no partner (the Global chip is not transcribed)."""
    import copy
    from oracle import logup
    from zkmips_b200 import synth
    chips, pv_of = synth.core_program_chips(9)
    assert [c.air for c in chips] == ["Cpu", "Program", "AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Mul",
                                      "DivRem", "MovCond", "Jump", "Branch", "MemoryLocal", "Byte"]
    airs = {a.name: a for a in (library.cpu(), library.program_chip(), library.add_sub(), library.bitwise(), library.lt(),
                                library.shift_left(), library.shift_right(), library.clo_clz(), library.mul(),
                                library.div_rem(), library.mov_cond(), library.jump(), library.branch(),
                                library.memory_local(), library.byte_chip())}
    branch = chips[12].canon[1]
    assert 0 < branch[:, 59].sum() < branch[:, 53:59].sum()            # taken and not-taken branches both occur
    assert chips[0].canon[1][:, 22].sum() > 0 and (chips[0].canon[1][:, 25] == 0).sum() > 0   # is_rw_a rows, non-sequential rows
    assert chips[0].canon[1][:, 23].sum() > 0 and chips[9].canon[1][:, 57:61].sum() > 0       # is_check_memory rows, divisions
    cpu = airs["Cpu"]
    assert (cpu.main_width, len(cpu.sends), len(cpu.receives), cpu.num_constraints) == (67, 16, 3, 57 + 10 + 3)
    assert cpu.main_width + 4 * cpu.perm_width + 8 == 119 and not cpu.local_only and cpu.uses_next_row()
    pvs = np.zeros(231, np.uint64)
    for k, v in pv_of.items():
        pvs[k] = v
    for ch in chips:
        assert all(not v.any() for v in _constraints_on_trace(airs[ch.air], ch.canon[1], ch.canon[0], pvs=pvs)), ch.name
    for cell in ((3, 6), (2, 38), (5, 45), (4, 0)):                    # next_pc, the written value, a timestamp limb, shard
        bad = chips[0].canon[1].copy()
        bad[cell] = (bad[cell] + 1) % ae_P
        assert any(v.any() for v in _constraints_on_trace(cpu, bad, None, pvs=pvs)[:57]), cell
    wrong = pvs.copy()
    wrong[41] += 4                                                     # public next_pc
    assert any(v.any() for v in _constraints_on_trace(cpu, chips[0].canon[1], None, pvs=wrong)[:57])
    rng = np.random.default_rng(3)
    alpha, beta = ([int(x) for x in rng.integers(1, ae_P, 4)] for _ in range(2))

    def bus_sum(kind, these):
        total = np.zeros(4, np.uint64)
        for ch in these:
            only = copy.copy(airs[ch.air])
            only.sends = [l for l in only.sends if l["kind"] == kind]
            only.receives = [l for l in only.receives if l["kind"] == kind]
            if only.sends or only.receives:
                _, lcs = logup.generate_permutation_trace(only, ch.canon[0], ch.canon[1], alpha, beta)
                total = (total + np.asarray(lcs, np.uint64)) % ae_P
        return total
    for kind in (1, 2, 3, 4):                                          # Memory, Program, Instruction, Byte
        assert not bus_sum(kind, chips).any(), kind
        assert bus_sum(kind, chips[1:]).any(), kind                    # without the CPU every one of them is open
    assert bus_sum(7, chips).any()                                     # Global: MemoryLocal's forwards


def test_exporter_output_without_logup_constraints_loads_to_the_same_program():
    """rust/air-export writes a chip's OWN constraints and its lookups; `Air.from_exported_json` appends the LogUp
    constraints.  Simulated here by stripping them from the hand transcriptions: same constraints, same generated
    CUDA, for chips with sends and receives, with a preprocessed trace, and with 8-lookup batches (DEGREE 9)."""
    for make in (library.add_sub, library.lt, library.base_alu, library.ext_alu, lambda: library.poseidon2_wide(3),
                 lambda: library.poseidon2_wide(9)):
        air = make()
        d = json.loads(air.to_json())
        n_logup = (air.perm_width - 1) + 3
        d["constraints"] = d["constraints"][:-n_logup]
        d["permutation_constraints_included"] = False
        d["perm_width"] = 0
        # the exporter never emits permutation / challenge / cumulative-sum leaves: cut the node list where they start
        cut = min(i for i, n in enumerate(d["nodes"]) if n[0] in ("perm", "chal", "lcs"))
        assert all(c < cut for c in d["constraints"])
        d["nodes"] = d["nodes"][:cut]
        loaded = Air.from_exported_json(json.dumps(d))
        assert loaded.num_constraints == air.num_constraints and loaded.perm_width == air.perm_width
        assert codegen.generate([loaded]) == codegen.generate([air])
