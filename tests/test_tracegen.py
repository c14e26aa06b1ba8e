"""Device trace generation (csrc/tracegen.cuh) against the oracle, and the oracle against the REFERENCE's own C++ row
fillers compiled into oracle/_ref (poseidon2_wide.hpp event_to_row / instr_to_row, add_sub.hpp event_to_row).

The reference's tests for these fillers compare the FFI rows with the Rust rows (crates/recursion/core/src/chips/
poseidon2_wide/trace.rs tests `generate_trace_deg_3` / `generate_trace_deg_9` / `generate_preprocessed_trace`,
crates/core/machine/src/alu/add_sub/mod.rs `test_generate_trace_ffi_eq_rust`); here the oracle plays the Rust side."""
import numpy as np
import pytest

from oracle import binding as ob
from zkmips_b200 import synth
from zkmips_b200.proof import to_monty
from . import backends

P = 0x7F000001


def _inputs(n, seed=1):
    rng = np.random.default_rng(seed)
    x = to_monty(rng.integers(0, P, (n, 16), dtype=np.uint64))
    if n:
        x[0] = 0                 # the padding row's input as a real event
    if n > 1:
        x[1] = to_monty(np.full(16, P - 1, np.uint64))
    return x


def _instrs(n, seed=2):
    rng = np.random.default_rng(seed)
    return to_monty(rng.integers(0, P, (n, 48), dtype=np.uint64))


# ------------------------------------------------------------------------------------------------ oracle vs reference
@pytest.mark.parametrize("sbox", [True, False])
def test_oracle_poseidon2_wide_rows_match_reference_cpp(sbox):
    R = ob.ref()
    if R is None or not hasattr(R, "ref_poseidon2_wide_event_to_row"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    x = _inputs(40)
    got = ob.poseidon2_wide_trace(x, 64, sbox)
    w = got.shape[1]
    zero = np.zeros(16, np.uint32)
    for r in range(64):
        row = np.zeros(w, np.uint32)
        inp = np.ascontiguousarray(x[r] if r < 40 else zero)
        R.ref_poseidon2_wide_event_to_row(ob._ptr(inp), ob._ptr(row), int(sbox))
        assert np.array_equal(row, got[r]), r
    # output_state is the permutation of the input (poseidon2_wide/trace.rs:311-318 asserts the same)
    assert np.array_equal(got[5, 156:172], ob.permute(x[5]))


def test_oracle_poseidon2_wide_prep_matches_reference_cpp():
    R = ob.ref()
    if R is None or not hasattr(R, "ref_poseidon2_wide_instr_to_row"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    ins = _instrs(11)
    got = ob.poseidon2_wide_prep(ins, 16)
    for r in range(11):
        row = np.zeros(49, np.uint32)
        R.ref_poseidon2_wide_instr_to_row(ob._ptr(np.ascontiguousarray(ins[r])), ob._ptr(row))
        assert np.array_equal(row, got[r])
    assert not got[11:].any()


def test_poseidon2_skinny_rows_match_reference_cpp():
    """The numpy filler of Poseidon2SkinnyChip (synth.poseidon2_skinny_rows / _prep_rows; the transcribed AIR is checked on
    its rows in tests/test_air_ir.py) against the reference's own poseidon2_skinny.hpp event_to_row / instr_to_row:
    eleven 28-word main rows per permutation, eleven 51-word preprocessed rows per instruction."""
    R = ob.ref()
    if R is None or not hasattr(R, "ref_poseidon2_skinny_event_to_rows"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    x = _inputs(5)
    got = to_monty(synth.poseidon2_skinny_rows(ob.from_monty(x), 64))
    for k in range(5):
        rows = np.zeros((11, 28), np.uint32)
        R.ref_poseidon2_skinny_event_to_rows(ob._ptr(np.ascontiguousarray(x[k])), ob._ptr(rows))
        assert np.array_equal(rows, got[11 * k:11 * k + 11]), k
        assert np.array_equal(rows[10, :16], ob.permute(x[k]))
    assert not got[55:].any()
    ins = ob.from_monty(_instrs(5)).astype(np.uint64)
    prep = to_monty(synth.poseidon2_skinny_prep_rows(ins[:, 0:16], ins[:, 16:32], ins[:, 32:48], 64))
    for k in range(5):
        for i in range(11):
            row = np.zeros(51, np.uint32)
            R.ref_poseidon2_skinny_instr_to_row(ob._ptr(np.ascontiguousarray(to_monty(ins[k]))), i, ob._ptr(row))
            assert np.array_equal(row, prep[11 * k + i]), (k, i)


def test_oracle_add_sub_rows_match_reference_cpp_and_numpy():
    ev, n = synth.add_sub_events(7)
    got = ob.add_sub_trace(ev, n)
    assert np.array_equal(got, to_monty(synth.add_sub_rows(ev, n)))
    R = ob.ref()
    if R is None or not hasattr(R, "ref_add_sub_event_to_row"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for r in range(len(ev)):
        row = np.zeros(19, np.uint32)
        R.ref_add_sub_event_to_row(ob._ptr(np.ascontiguousarray(ev[r])), ob._ptr(row))
        assert np.array_equal(row, got[r]), r


# ------------------------------------------------------------------------------------------------ device vs oracle
def _check_poseidon2_wide(ctx, n_events, rows, sbox):
    x = _inputs(n_events, seed=rows + n_events)
    dptr, w = ctx.tracegen_poseidon2_wide(x, rows, sbox)
    got = ctx.download(dptr, (rows, w))
    ctx.dev_free(dptr)
    assert np.array_equal(got, ob.poseidon2_wide_trace(x, rows, sbox))


def _check_alu(ctx, chip, log_n, fill):
    events_of = {"AddSub": synth.add_sub_events, "Bitwise": synth.bitwise_events, "Lt": synth.lt_events,
                 "ShiftLeft": synth.shift_left_events, "ShiftRight": synth.shift_right_events,
                 "CloClz": synth.clo_clz_events}[chip]
    rows_of = {"AddSub": synth.add_sub_rows, "Bitwise": synth.bitwise_rows, "Lt": synth.lt_rows,
               "ShiftLeft": synth.shift_left_rows, "ShiftRight": synth.shift_right_rows, "CloClz": synth.clo_clz_rows}[chip]
    ev, n = events_of(log_n, fill=fill)
    dptr, w = ctx.tracegen_alu(chip, ev, n)
    got = ctx.download(dptr, (n, w))
    ctx.dev_free(dptr)
    assert np.array_equal(got, to_monty(rows_of(ev, n)))


def _check_prep(ctx, n, rows):
    ins = _instrs(n)
    dptr, w = ctx.tracegen_poseidon2_wide_prep(ins, rows)
    got = ctx.download(dptr, (rows, w))
    ctx.dev_free(dptr)
    assert np.array_equal(got, ob.poseidon2_wide_prep(ins, rows))


def _check_poseidon2_skinny(ctx, n_events, rows):
    """device filler against the numpy filler, which is pinned against the reference's poseidon2_skinny.hpp above"""
    x = _inputs(n_events, seed=rows + n_events)
    dptr, w = ctx.tracegen_poseidon2_skinny(x, rows)
    got = ctx.download(dptr, (rows, w))
    ctx.dev_free(dptr)
    assert w == 28 and np.array_equal(got, to_monty(synth.poseidon2_skinny_rows(ob.from_monty(x), rows)))
    ins = _instrs(n_events, seed=3 + n_events)
    dptr, w = ctx.tracegen_poseidon2_skinny_prep(ins, rows)
    got = ctx.download(dptr, (rows, w))
    ctx.dev_free(dptr)
    c = ob.from_monty(ins).astype(np.uint64)
    assert w == 51 and np.array_equal(got, to_monty(synth.poseidon2_skinny_prep_rows(c[:, 0:16], c[:, 16:32], c[:, 32:48], rows)))


def _check_cpu(ctx, log_cpu):
    """device filler of the CPU chip against the numpy rows of the toy core-machine program (which satisfy the transcribed
    CPU AIR and balance the machine's buses, tests/test_air_ir.py)"""
    chips, _ = synth.core_program_chips(log_cpu)
    cpu = chips[0]
    dptr, w = ctx.tracegen_cpu(cpu.cpu_events, cpu.height)
    got = ctx.download(dptr, (cpu.height, w))
    ctx.dev_free(dptr)
    assert w == 67 and np.array_equal(got, to_monty(cpu.canon[1]))


def test_tracegen_emu():
    """kernel index math on the CPU emulator (test-only build of the same sources)"""
    ctx = backends.emu()
    _check_poseidon2_wide(ctx, 100, 256, True)
    _check_poseidon2_wide(ctx, 3, 8, False)
    for chip in ("AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz"):
        _check_alu(ctx, chip, 8, 0.7)
    _check_prep(ctx, 5, 8)
    _check_poseidon2_skinny(ctx, 40, 512)      # two CTAs, the second one partial and partly padding
    _check_poseidon2_skinny(ctx, 1, 16)
    _check_cpu(ctx, 8)


@pytest.mark.gpu
@pytest.mark.parametrize("sbox", [True, False])
@pytest.mark.parametrize("n_events,rows", [(0, 1), (1, 1), (5, 8), (127, 128), (129, 256), (40000, 1 << 16), (1 << 14, 1 << 14)])
def test_tracegen_poseidon2_wide_gpu(n_events, rows, sbox):
    _check_poseidon2_wide(backends.gpu(), n_events, rows, sbox)


@pytest.mark.gpu
@pytest.mark.parametrize("log_cpu", [2, 7, 13])
def test_tracegen_cpu_gpu(log_cpu):
    _check_cpu(backends.gpu(), log_cpu)


@pytest.mark.gpu
@pytest.mark.parametrize("n_events,rows", [(0, 1), (1, 16), (2, 32), (32, 512), (93, 1024), (5957, 1 << 16), (23000, 1 << 18)])
def test_tracegen_poseidon2_skinny_gpu(n_events, rows):
    _check_poseidon2_skinny(backends.gpu(), n_events, rows)


@pytest.mark.gpu
@pytest.mark.parametrize("chip", ["AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz"])
@pytest.mark.parametrize("log_n,fill", [(0, 1.0), (3, 0.5), (7, 1.0), (12, 0.75), (17, 0.9)])
def test_tracegen_alu_gpu(chip, log_n, fill):
    _check_alu(backends.gpu(), chip, log_n, fill)


@pytest.mark.gpu
def test_tracegen_prep_and_commit_gpu():
    """the generated trace goes straight into zk_commit_dev: root equals the oracle's commit of the oracle's trace"""
    ctx = backends.gpu()
    _check_prep(ctx, 1000, 1024)
    x = _inputs(3000)
    rows = 4096
    dptr, w = ctx.tracegen_poseidon2_wide(x, rows, True)
    one = int(to_monty(np.array([1]))[0])  # trace-domain shift 1
    root, pd = ctx.commit_dev([dptr], [(rows, w)], [one], 1)
    tree = ob.pcs_commit([ob.poseidon2_wide_trace(x, rows, True)], 1)
    assert np.array_equal(root, tree.root)
    pd.free()
    ctx.dev_free(dptr)


@pytest.mark.gpu
def test_tracegen_errors_gpu():
    from zkmips_b200.native import ZkError
    ctx = backends.gpu()
    with pytest.raises(ZkError):
        ctx.tracegen_poseidon2_wide(_inputs(9), 8, True)     # more events than rows
    with pytest.raises(ZkError):
        ctx.tracegen_alu("AddSub", synth.add_sub_events(4)[0], 12)  # not a power of two
