"""Synthetic workloads (numpy only, no oracle): the BASELINE config-2 traces, valid traces for the AIRs compiled into
libzkgpu, and the chip shapes of Ziren's maximal execution shards.  Used by bench.py and by the tests."""
import numpy as np

from .proof import to_monty
from .prover import Chip

P = 0x7F000001


def M(canon):
    """canonical integers -> Montgomery words"""
    return to_monty(np.asarray(canon, dtype=np.uint64) % P)


def splitmix64(seed, n, offset=0):
    """Uniform canonical field elements < p from splitmix64: outputs
    offset .. offset + n - 1 of the stream (counter based, so a large matrix can be produced in chunks)."""
    x = np.uint64(seed)
    with np.errstate(over="ignore"):
        idx = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = x + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(P)).astype(np.uint32)


def config2_trace(kind, log_rows=20, cols=256, seed=0x5A4B4D49):
    """BASELINE config 2 inputs in Montgomery form:
    'a' = from_canonical((r * W + c) mod p)  (mirrors recursion/circuit/src/fri.rs:832-835),
    'b' = uniform canonical values from splitmix64(seed = 0x5A4B4D49)."""
    n = (1 << log_rows) * cols
    out = np.empty(n, np.uint32)
    step = 1 << 24
    for off in range(0, n, step):
        m = min(step, n - off)
        if kind == "a":
            canon = (np.arange(off, off + m, dtype=np.uint64) % P).astype(np.uint32)
        else:
            canon = splitmix64(seed, m, off)
        out[off:off + m] = to_monty(canon)
    return out.reshape(1 << log_rows, cols)


# ---------------------------------------------------------------------------------------------- chips
def fibonacci_chip(log_n, a=1, b=1, name="Fibonacci"):
    """generate_trace_rows of crates/stark/src/stark_testing.rs:63-81; `chip.pvs` = (a, b, last) are the public values
    0..2 the AIR reads (the shard's public-values vector must start with them)."""
    n = 1 << log_n
    t = np.zeros((n, 2), np.uint64)
    t[0] = (a, b)
    for i in range(1, n):
        t[i, 0] = t[i - 1, 1]
        t[i, 1] = (t[i - 1, 0] + t[i - 1, 1]) % P
    c = Chip(name, "fibonacci", M(t))
    c.pvs = [a, b, int(t[n - 1, 1])]
    return c


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


LOOKUP_PV3 = 7  # public value 3: the increment lookup_pair's column 0 takes per row


def lookup_chip(log_n, seed=3, name="Lookup"):
    """valid trace for library.lookup_pair (its LogUp permutation trace is generated on the device).  Its sends and
    receives do NOT balance, so a shard holding it fails the verifier's final check (local cumulative sum != 0)."""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    p = rng.integers(0, P, (n, 2)).astype(np.uint64)
    m = np.zeros((n, 5), np.uint64)
    m[0, 0] = 5
    for i in range(1, n):
        m[i, 0] = (m[i - 1, 0] + LOOKUP_PV3) % P
    m[:, 1] = rng.integers(0, P, n)
    m[:, 2] = (m[:, 0] * m[:, 1] + p[:, 0]) % P
    m[:, 3] = rng.integers(0, 2, n)
    m[:, 4] = rng.integers(0, 5, n)
    c = Chip(name, "lookup_pair", M(m), preprocessed=M(p))
    c.canon = (p, m)
    return c


def lookup_side_chips(log_n, seed=9):
    """`LookupSend` / `LookupRecv`: the same (x, y, m) columns on both sides, so their LogUp sums cancel."""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 4), np.uint64)
    t[:, 0] = rng.integers(0, P, n)
    t[:, 1] = rng.integers(0, P, n)
    t[:, 2] = rng.integers(0, 4, n)
    return (Chip("LookupSend", "lookup_send", M(t), local_only=True),
            Chip("LookupRecv", "lookup_recv", M(t.copy()), local_only=True))


def global_chip(log_n, seed=13, name="GlobalTail"):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 16), np.uint64)
    t[:, 0] = rng.integers(0, 2, n)
    t[1:, 1] = np.cumsum(t[:-1, 0]) % P
    t[:, 2:] = rng.integers(0, P, (n, 14))
    return Chip(name, "global_tail", M(t), commit_scope="global")


def local_bool_chip(log_n, seed=17, name="LocalBool"):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 8), np.uint64)
    for g in range(2):
        a, b = rng.integers(0, 2, n), rng.integers(0, 2, n)
        x = a ^ b
        c = np.where(x == a, rng.integers(0, 2, n), 0)  # c * x = c * a
        t[:, 4 * g], t[:, 4 * g + 1], t[:, 4 * g + 2], t[:, 4 * g + 3] = a, b, x, c
    return Chip(name, "local_bool", M(t), local_only=True)


def public_values_for(chips, n=8):
    """the shard's public-values vector (Montgomery): Fibonacci's (a, b, last) at 0..2, lookup_pair's increment at 3"""
    pv = np.zeros(n, np.uint64)
    for c in chips:
        if c.air == "fibonacci":
            pv[0:3] = c.pvs
    pv[3] = LOOKUP_PV3
    for c in chips:
        for k, v in (getattr(c, "core_pvs", None) or {}).items():      # start_pc, next_pc, execution_shard of the core machine
            pv[k] = v
        if getattr(c, "pv_digest", None) is not None:                  # RecursionPublicValues.digest (air/public_values.rs:144)
            pv[223:231] = c.pv_digest
    return M(pv)




# ------------------------------------------------------------------------------------------------------------------
# Real Ziren ALU chips (library.add_sub / lt / bitwise): numpy restatements of their `event_to_row` on random events.
# `fill` of the 2^log_n rows are real events, the rest is the zero padding the reference leaves (is_real = 0).
# ------------------------------------------------------------------------------------------------------------------
def _bytes(x):
    x = np.asarray(x, np.uint64)
    return np.stack([(x >> np.uint64(8 * k)) & np.uint64(0xFF) for k in range(4)], axis=1)


def _events(log_n, seed, fill):
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    real = max(1, int(n * fill)) if n > 1 else 1
    pc = (0x1000 + 4 * np.arange(real, dtype=np.uint64)) % P
    b = rng.integers(0, 1 << 32, real, dtype=np.uint64)
    c = rng.integers(0, 1 << 32, real, dtype=np.uint64)
    return n, real, rng, pc, b, c


def _alu_event_array(pc, opcode, a, b, c):
    """`AluEvent` records in their #[repr(C)] layout (crates/core/executor/src/events/instr.rs:10-26): u32 pc, next_pc,
    u8 opcode (+ 3 padding bytes), u32 hi, a, b, c = 7 words -- what `zk_tracegen_alu` takes."""
    ev = np.zeros((len(pc), 7), np.uint32)
    ev[:, 0], ev[:, 1], ev[:, 2] = pc, (np.asarray(pc, np.uint64) + 4) % P, opcode
    ev[:, 4], ev[:, 5], ev[:, 6] = a, b, c
    return ev


def add_sub_events(log_n, seed=21, fill=0.75):
    """random ADD / SUB events: (events [real, 7], padded height)"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    is_add = rng.integers(0, 2, real).astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    a = np.where(is_add == 1, (b + c) & mask, (b - c) & mask)       # ADD: a = b + c;  SUB: a = b - c
    return _alu_event_array(pc, np.where(is_add == 1, 0, 1), a, b, c), n


def add_sub_rows(events, n):
    """AddSubChip::event_to_row + AddOperation::populate (alu/add_sub/mod.rs:150-172, operations/add.rs:26-60):
    canonical rows, zero padding"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    is_add = (ev[:, 2] & np.uint64(0xFF)) == 0
    mask = np.uint64(0xFFFFFFFF)
    op1 = np.where(is_add, ev[:, 5], ev[:, 4])                      # operand_1: b for ADD, a for SUB
    op2 = ev[:, 6]
    value = (op1 + op2) & mask
    x, y = _bytes(op1), _bytes(op2)
    carry = np.zeros((real, 3), np.uint64)
    carry[:, 0] = (x[:, 0] + y[:, 0]) > 255
    carry[:, 1] = (x[:, 1] + y[:, 1] + carry[:, 0]) > 255
    carry[:, 2] = (x[:, 2] + y[:, 2] + carry[:, 1]) > 255
    t = np.zeros((n, 19), np.uint64)
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:9] = _bytes(value), carry
    t[:real, 9:13], t[:real, 13:17] = x, y
    t[:real, 17], t[:real, 18] = is_add, ~is_add
    return t


def add_sub_chip(log_n, seed=21, fill=0.75, name="AddSub", device=False):
    """`device=True`: no host trace -- the chip carries its AluEvent records and the prover fills the rows on the GPU
    (zk_tracegen_alu), so only 28 bytes per event cross PCIe."""
    ev, n = add_sub_events(log_n, seed, fill)
    if device:
        return Chip(name, "AddSub", None, local_only=True, events=ev, tracegen="AddSub", rows=n)
    t = add_sub_rows(ev, n)
    ch = Chip(name, "AddSub", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


def bitwise_events(log_n, seed=22, fill=0.75):
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    op = rng.integers(0, 4, real)  # 0 nor, 1 xor, 2 or, 3 and
    mask = np.uint64(0xFFFFFFFF)
    a = np.select([op == 0, op == 1, op == 2], [~(b | c) & mask, b ^ c, b | c], b & c)
    return _alu_event_array(pc, np.array([18, 17, 16, 15])[op], a, b, c), n   # Opcode::NOR / XOR / OR / AND


def bitwise_rows(events, n):
    """BitwiseChip::event_to_row (alu/bitwise/mod.rs:141-170)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 18), np.uint64)
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10], t[:real, 10:14] = _bytes(ev[:, 4]), _bytes(ev[:, 5]), _bytes(ev[:, 6])
    for k, opc in enumerate((18, 17, 16, 15)):                      # is_nor, is_xor, is_or, is_and
        t[:real, 14 + k] = (ev[:, 2] & np.uint64(0xFF)) == opc
    return t


def bitwise_chip(log_n, seed=22, fill=0.75, name="Bitwise", device=False):
    """`device=True`: no host trace -- the chip carries its AluEvent records and the prover fills the rows on the GPU
    (zk_tracegen_alu), so only 28 bytes per event cross PCIe."""
    ev, n = bitwise_events(log_n, seed, fill)
    if device:
        return Chip(name, "Bitwise", None, local_only=True, events=ev, tracegen="Bitwise", rows=n)
    t = bitwise_rows(ev, n)
    ch = Chip(name, "Bitwise", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


def lt_events(log_n, seed=23, fill=0.75):
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    eq = rng.integers(0, 8, real) == 0
    c = np.where(eq, b, c)                                          # some equal operands (is_comp_eq = 1)
    near = rng.integers(0, 4, real) == 0
    c = np.where(near & ~eq, (b & np.uint64(0xFFFF0000)) | (c & np.uint64(0xFFFF)), c)  # differ in a LOW byte only
    is_slt = rng.integers(0, 2, real).astype(np.uint64)
    bs = np.where(b >= (1 << 31), b.astype(np.int64) - (1 << 32), b.astype(np.int64))
    cs = np.where(c >= (1 << 31), c.astype(np.int64) - (1 << 32), c.astype(np.int64))
    a = np.where(is_slt == 1, bs < cs, b < c).astype(np.uint64)
    return _alu_event_array(pc, np.where(is_slt == 1, 13, 14), a, b, c), n   # Opcode::SLT / SLTU


def lt_rows(events, n):
    """LtChip::event_to_row (alu/lt/mod.rs:179-262)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    b, c = ev[:, 5], ev[:, 6]
    is_slt = ((ev[:, 2] & np.uint64(0xFF)) == 13).astype(np.uint64)
    bb, cb = _bytes(b), _bytes(c)
    b_masked, c_masked = bb[:, 3] & np.uint64(0x7F), cb[:, 3] & np.uint64(0x7F)
    b_comp, c_comp = bb.copy(), cb.copy()
    b_comp[:, 3] = np.where(is_slt == 1, b_masked, bb[:, 3])
    c_comp[:, 3] = np.where(is_slt == 1, c_masked, cb[:, 3])
    t = np.zeros((n, 36), np.uint64)
    flags = np.zeros((real, 4), np.uint64)
    sltu = np.zeros(real, np.uint64)
    inv = np.zeros(real, np.uint64)
    cmp_bytes = np.zeros((real, 2), np.uint64)
    done = np.zeros(real, bool)
    for k in (3, 2, 1, 0):                                          # most significant differing byte
        hit = ~done & (b_comp[:, k] != c_comp[:, k])
        flags[hit, k] = 1
        sltu[hit] = b_comp[hit, k] < c_comp[hit, k]
        cmp_bytes[hit, 0], cmp_bytes[hit, 1] = b_comp[hit, k], c_comp[hit, k]
        done |= hit
    diff = (cmp_bytes[:, 0] + np.uint64(P) - cmp_bytes[:, 1]) % np.uint64(P)
    inv[done] = [pow(int(d), P - 2, P) for d in diff[done]]
    msb_b, msb_c = bb[:, 3] >> np.uint64(7), cb[:, 3] >> np.uint64(7)
    is_sign_eq = np.where(is_slt == 1, msb_b == msb_c, 1).astype(np.uint64)
    bit_b, bit_c = msb_b * is_slt, msb_c * is_slt
    a0 = bit_b * (1 - bit_c) + is_sign_eq * sltu
    t[:real, 0], t[:real, 1], t[:real, 2], t[:real, 3] = ev[:, 0], ev[:, 1], is_slt, 1 - is_slt
    t[:real, 4] = a0
    t[:real, 8:12], t[:real, 12:16], t[:real, 16:20] = bb, cb, flags
    t[:real, 20], t[:real, 21], t[:real, 22] = b_masked, c_masked, inv
    t[:real, 23], t[:real, 24], t[:real, 25], t[:real, 26] = msb_b, msb_c, bit_b, bit_c
    t[:real, 27], t[:real, 28], t[:real, 29] = sltu, ~done, is_sign_eq
    t[:real, 30:32] = cmp_bytes
    return t


def lt_chip(log_n, seed=23, fill=0.75, name="Lt", device=False):
    """`device=True`: no host trace -- the chip carries its AluEvent records and the prover fills the rows on the GPU
    (zk_tracegen_alu), so only 28 bytes per event cross PCIe."""
    ev, n = lt_events(log_n, seed, fill)
    if device:
        return Chip(name, "Lt", None, local_only=True, events=ev, tracegen="Lt", rows=n)
    t = lt_rows(ev, n)
    assert np.array_equal(t[:len(ev), 4], np.asarray(ev[:, 4], np.uint64)), "LtChip row disagrees with the event's a"
    ch = Chip(name, "Lt", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


_BYTE_INVERSES = np.array([0] + [pow(x, P - 2, P) for x in range(1, 256)], np.uint64)


def mov_cond_events(log_n, seed=24, fill=0.75):
    """random MEQ / MNE / WSBH events: (pc, next_pc, opcode, a, b, c, prev_a) per row (MovCondEvent,
    crates/core/executor/src/events/instr.rs), c == 0 in about a third of the conditional moves"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    prev_a = rng.integers(0, 1 << 32, real, dtype=np.uint64)
    c = np.where(rng.integers(0, 3, real) == 0, 0, c).astype(np.uint64)
    c = np.where(rng.integers(0, 5, real) == 0, c & np.uint64(0xFF00), c)         # zero bytes inside a non-zero word
    op = rng.integers(0, 3, real)                                                 # 0 MEQ, 1 MNE, 2 WSBH
    wsbh = ((b & np.uint64(0x00FF00FF)) << np.uint64(8)) | ((b & np.uint64(0xFF00FF00)) >> np.uint64(8))
    a = np.select([op == 0, op == 1], [np.where(c == 0, b, prev_a), np.where(c != 0, b, prev_a)], wsbh)
    ev = np.stack([pc, (pc + 4) % P, np.array([50, 51, 52], np.uint64)[op], a, b, c, prev_a], axis=1).astype(np.uint64)
    return ev, n


def mov_cond_rows(events, n):
    """MovCondChip::event_to_row + IsZeroWordOperation::populate (misc/mov_cond/mod.rs:125-143, operations/is_zero_word.rs:
    26-42, is_zero.rs:31-43): canonical rows, zero padding"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 32), np.uint64)
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10] = _bytes(ev[:, 3]), _bytes(ev[:, 6])
    t[:real, 10:14], t[:real, 14:18] = _bytes(ev[:, 4]), _bytes(ev[:, 5])
    cb = _bytes(ev[:, 5])
    inv = _BYTE_INVERSES[cb.astype(np.int64)]
    zero = (cb == 0).astype(np.uint64)
    t[:real, 18:26:2], t[:real, 19:26:2] = inv, zero
    t[:real, 26], t[:real, 27] = zero[:, 0] * zero[:, 1], zero[:, 2] * zero[:, 3]
    t[:real, 28] = ev[:, 5] == 0
    t[:real, 29], t[:real, 30], t[:real, 31] = ev[:, 2] == 51, ev[:, 2] == 50, ev[:, 2] == 52
    return t


def mov_cond_chip(log_n, seed=24, fill=0.75, name="MovCond"):
    ev, n = mov_cond_events(log_n, seed, fill)
    t = mov_cond_rows(ev, n)
    ch = Chip(name, "MovCond", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


def _kb_range_cols(v):
    """KoalaBearWordRangeChecker::populate (operations/koala_bear_word.rs:28-43): bits of the top byte and their running
    AND from bit 0"""
    v = np.asarray(v, np.uint64)
    bits = np.stack([(v >> np.uint64(24 + i)) & np.uint64(1) for i in range(8)], axis=1)
    ands = np.zeros((len(v), 6), np.uint64)
    acc = bits[:, 0] * bits[:, 1]
    ands[:, 0] = acc
    for k in range(1, 6):
        acc = acc * bits[:, k + 1]
        ands[:, k] = acc
    return np.concatenate([bits, ands], axis=1)


def jump_events(log_n, seed=25, fill=0.75):
    """random Jump / Jumpi / JumpDirect events (JumpEvent: pc, next_pc, next_next_pc, opcode, a, b, c): a = next_pc + 4;
    Jump / Jumpi jump to b, JumpDirect to next_pc + b; every address is below p (one row sits on the largest)"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    next_pc = (pc + 4) % P
    op = rng.integers(0, 3, real)                                                  # 0 Jump, 1 Jumpi, 2 JumpDirect
    target = rng.integers(0, P - 8, real, dtype=np.uint64)
    target[0] = P - 1                                                              # 0x7F000000: the range checker's edge
    b = np.where(op == 2, (target + (1 << 32) - next_pc) & np.uint64(0xFFFFFFFF), target)
    ev = np.stack([pc, next_pc, target, np.array([27, 28, 29], np.uint64)[op], next_pc + 4, b, c], axis=1).astype(np.uint64)
    return ev, n


def jump_rows(events, n):
    """JumpChip::event_to_row (control_flow/jump/trace.rs:81-102)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 66), np.uint64)
    t[:real, 0] = ev[:, 0]
    t[:real, 1:5], t[:real, 5:19] = _bytes(ev[:, 1]), _kb_range_cols(ev[:, 1])
    t[:real, 19:23], t[:real, 23:37] = _bytes(ev[:, 2]), _kb_range_cols(ev[:, 2])
    t[:real, 37:41], t[:real, 41:45], t[:real, 45:49] = _bytes(ev[:, 4]), _bytes(ev[:, 5]), _bytes(ev[:, 6])
    t[:real, 49], t[:real, 50], t[:real, 51] = ev[:, 3] == 27, ev[:, 3] == 28, ev[:, 3] == 29
    t[:real, 52:66] = _kb_range_cols(ev[:, 4])
    return t


def jump_chip(log_n, seed=25, fill=0.75, name="Jump"):
    ev, n = jump_events(log_n, seed, fill)
    t = jump_rows(ev, n)
    ch = Chip(name, "Jump", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


def branch_events(log_n, seed=26, fill=0.75):
    """random BEQ / BNE / BLTZ / BLEZ / BGTZ / BGEZ events (BranchEvent: pc, next_pc, next_next_pc, opcode, a, b, c);
    b = 0 for the compare-with-zero forms, a == b in a third of the BEQ / BNE rows; the offset c is signed"""
    n, real, rng, pc, b, _ = _events(log_n, seed, fill)
    pc = pc + 0x10000
    next_pc = pc + 4
    op = rng.integers(0, 6, real)                                                  # index into the opcode list below
    opc = np.array([21, 26, 25, 24, 23, 22], np.uint64)[op]                        # BEQ BNE BLTZ BLEZ BGTZ BGEZ
    a = rng.integers(0, 1 << 32, real, dtype=np.uint64)
    a = np.where(rng.integers(0, 4, real) == 0, 0, a).astype(np.uint64)
    b = np.where(op >= 2, 0, np.where(rng.integers(0, 3, real) == 0, a, b)).astype(np.uint64)
    sa, sb = a.astype(np.uint32).view(np.int32), b.astype(np.uint32).view(np.int32)
    lt, gt, eq = sa < sb, sa > sb, sa == sb
    branching = np.select([op == 0, op == 1, op == 2, op == 3, op == 4], [eq, ~eq, lt, lt | eq, gt], eq | gt)
    off = rng.integers(-0x4000, 0x4000, real) * 4
    c = (off & 0xFFFFFFFF).astype(np.uint64)
    target = (next_pc + c) & np.uint64(0xFFFFFFFF)
    nnpc = np.where(branching, target, next_pc + 4)
    return np.stack([pc, next_pc, nnpc, opc, a, b, c], axis=1).astype(np.uint64), n


def branch_rows(events, n):
    """BranchChip::event_to_row (control_flow/branch/trace.rs:81-131)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    a, b = ev[:, 4].astype(np.uint32).view(np.int32), ev[:, 5].astype(np.uint32).view(np.int32)
    lt, gt, eq = a < b, a > b, a == b
    opc = ev[:, 3]
    branching = np.select([opc == 21, opc == 26, opc == 25, opc == 24, opc == 23], [eq, ~eq, lt, lt | eq, gt], eq | gt)
    t = np.zeros((n, 62), np.uint64)
    t[:real, 0] = ev[:, 0]
    t[:real, 1:5], t[:real, 5:19] = _bytes(ev[:, 1]), _kb_range_cols(ev[:, 1])
    t[:real, 19:23] = _bytes((ev[:, 1] + ev[:, 6]) & np.uint64(0xFFFFFFFF))
    t[:real, 23:27], t[:real, 27:41] = _bytes(ev[:, 2]), _kb_range_cols(ev[:, 2])
    t[:real, 41:45], t[:real, 45:49], t[:real, 49:53] = _bytes(ev[:, 4]), _bytes(ev[:, 5]), _bytes(ev[:, 6])
    for k, o in enumerate((21, 26, 25, 24, 23, 22)):                               # is_beq, is_bne, is_bltz, is_blez, is_bgtz, is_bgez
        t[:real, 53 + k] = opc == o
    t[:real, 59], t[:real, 60], t[:real, 61] = branching, gt, lt
    return t


def branch_chip(log_n, seed=26, fill=0.75, name="Branch"):
    ev, n = branch_events(log_n, seed, fill)
    t = branch_rows(ev, n)
    ch = Chip(name, "Branch", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


# the (a, b, c) triples of the reference's own ShiftLeft test (alu/sll/mod.rs `prove_koalabear`): a = b << (c mod 32)
SLL_REFERENCE_CASES = [
    (0x00000002, 0x00000001, 1), (0x00000080, 0x00000001, 7), (0x00004000, 0x00000001, 14), (0x80000000, 0x00000001, 31),
    (0xffffffff, 0xffffffff, 0), (0xfffffffe, 0xffffffff, 1), (0xffffff80, 0xffffffff, 7), (0xffffc000, 0xffffffff, 14),
    (0x80000000, 0xffffffff, 31), (0x21212121, 0x21212121, 0), (0x42424242, 0x21212121, 1), (0x90909080, 0x21212121, 7),
    (0x48484000, 0x21212121, 14), (0x80000000, 0x21212121, 31), (0x21212121, 0x21212121, 0xffffffe0),
    (0x42424242, 0x21212121, 0xffffffe1), (0x90909080, 0x21212121, 0xffffffe7), (0x48484000, 0x21212121, 0xffffffee),
    (0x00000000, 0x21212120, 0xffffffff)]


def shift_left_events(log_n, seed=27, fill=0.75):
    """SLL AluEvents: the reference test's cases first, then random ones (a = b << (c mod 32))"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    c = np.where(rng.integers(0, 2, real) == 0, c & np.uint64(31), c)
    k = min(real, len(SLL_REFERENCE_CASES))
    b[:k] = [t[1] for t in SLL_REFERENCE_CASES[:k]]
    c[:k] = [t[2] for t in SLL_REFERENCE_CASES[:k]]
    a = (b << (c & np.uint64(31))) & np.uint64(0xFFFFFFFF)
    assert all(int(a[i]) == SLL_REFERENCE_CASES[i][0] for i in range(k))
    return _alu_event_array(pc, 9, a, b, c), n


def shift_left_rows(events, n):
    """ShiftLeft::event_to_row (alu/sll/mod.rs:150-215) and the non-zero padding row of generate_trace (:95-106)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 44), np.uint64)
    t[:, 22], t[:, 39], t[:, 30] = 1, 1, 1                                         # padding: shift by 0 bits / 0 bytes
    a, b, c = ev[:, 4], ev[:, 5], ev[:, 6]
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10], t[:real, 10:14] = _bytes(a), _bytes(b), _bytes(c)
    for i in range(8):
        t[:real, 14 + i] = (c >> np.uint64(i)) & np.uint64(1)
        t[:real, 22 + i] = (c % np.uint64(8)) == i
    mult = np.uint64(1) << (c % np.uint64(8))
    t[:real, 30] = mult
    bb = _bytes(b)
    carry = np.zeros(real, np.uint64)
    for i in range(4):
        v = bb[:, i] * mult + carry
        carry = v >> np.uint64(8)
        t[:real, 31 + i], t[:real, 35 + i] = v & np.uint64(0xFF), carry
    nb = (c & np.uint64(31)) >> np.uint64(3)
    for i in range(4):
        t[:real, 39 + i] = nb == i
    t[:real, 43] = 1
    return t


def shift_left_chip(log_n, seed=27, fill=0.75, name="ShiftLeft", device=False):
    ev, n = shift_left_events(log_n, seed, fill)
    if device:                                                    # rows filled on the GPU from the AluEvents (zk_tracegen_alu)
        return Chip(name, "ShiftLeft", None, local_only=True, events=ev, tracegen="ShiftLeft", rows=n)
    t = shift_left_rows(ev, n)
    ch = Chip(name, "ShiftLeft", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


# the (opcode, a, b) cases of the reference's own CloClz test (alu/clo_clz/mod.rs `prove_koalabear`)
CLOCLZ_REFERENCE_CASES = [(19, 32, 0), (19, 8, 0x00800000), (19, 0, 0xffffffff), (20, 32, 0xffffffff), (20, 8, 0xff7fffff),
                          (20, 0, 0)]


def clo_clz_events(log_n, seed=28, fill=0.75):
    """CLZ / CLO AluEvents (c = 0): the reference test's cases first, then random words with a random number of leading
    zeros / ones"""
    n, real, rng, pc, b, _ = _events(log_n, seed, fill)
    op = rng.integers(0, 2, real)                                                  # 0 CLZ, 1 CLO
    b = b >> rng.integers(0, 33, real).astype(np.uint64)                           # CLZ operand with that many zeros
    k = min(real, len(CLOCLZ_REFERENCE_CASES))
    bb = b.copy()
    b = np.where(op == 1, np.uint64(0xFFFFFFFF) - b, b)
    for i, (o, _, bv) in enumerate(CLOCLZ_REFERENCE_CASES[:k]):
        op[i], b[i] = o - 19, bv
        bb[i] = bv if o == 19 else 0xFFFFFFFF - bv
    a = (32 - np.where(bb == 0, 0, np.floor(np.log2(np.maximum(bb, 1).astype(np.float64))).astype(np.int64) + 1)).astype(np.uint64)
    assert all(int(a[i]) == CLOCLZ_REFERENCE_CASES[i][1] for i in range(k))
    return _alu_event_array(pc, 19 + op, a, b, np.zeros(real, np.uint64)), n


def clo_clz_rows(events, n):
    """CloClzChip::generate_trace (alu/clo_clz/mod.rs:64-128): padding rows are CLZ of zero"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 22), np.uint64)
    t[:, 2], t[:, 19], t[:, 14] = 32, 1, 1
    a, b = ev[:, 4], ev[:, 5]
    clo = (ev[:, 2] & np.uint64(0xFF)) == 20
    bb = np.where(clo, np.uint64(0xFFFFFFFF) - b, b)
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10], t[:real, 10:14] = _bytes(a), _bytes(b), _bytes(bb)
    t[:real, 14] = bb == 0
    sr1 = np.where(bb == 0, 0, bb >> (np.uint64(31) - np.minimum(a, 31)))
    t[:real, 15:19] = _bytes(sr1)
    t[:real, 19], t[:real, 20], t[:real, 21] = ~clo, clo, 1
    return t


def clo_clz_chip(log_n, seed=28, fill=0.75, name="CloClz", device=False):
    ev, n = clo_clz_events(log_n, seed, fill)
    if device:                                                    # rows filled on the GPU from the AluEvents (zk_tracegen_alu)
        return Chip(name, "CloClz", None, events=ev, tracegen="CloClz", rows=n)
    t = clo_clz_rows(ev, n)
    ch = Chip(name, "CloClz", M(t))
    ch.canon, ch.events = (None, t), ev
    return ch


def byte_table():
    """ByteChip::trace (crates/core/machine/src/bytes/mod.rs:36-104), canonical: one row per byte pair (b, c), index
    b * 256 + c: b, c, and, or, xor, nor, sll, shr, shr_carry, ltu, msb, value_u16"""
    bc = np.arange(1 << 16, dtype=np.uint64)
    b, c = bc >> np.uint64(8), bc & np.uint64(0xFF)
    k = c & np.uint64(7)
    shr = b >> k
    carry = np.where(k == 0, 0, ((b << (np.uint64(8) - k)) & np.uint64(0xFF)) >> (np.uint64(8) - k))
    return np.stack([b, c, b & c, b | c, b ^ c, (~(b | c)) & np.uint64(0xFF), (b << k) & np.uint64(0xFF), shr, carry,
                     b < c, b >> np.uint64(7), bc], axis=1).astype(np.uint64)


_AIR_CACHE = {}


def _air(name):
    if not _AIR_CACHE:
        from .air import library
        _AIR_CACHE.update({a.name: a for a in library.all_airs()})
    return _AIR_CACHE[name]


def _lf_values(lf, prep, main, h):
    """VirtualPairCol::apply over all rows of a canonical trace (an affine form of the row's columns)"""
    c, terms = lf
    acc = np.full(h, int(c) % P, np.uint64)
    for (t, col, w) in terms:
        src = main if t == "main" else prep
        acc = (acc + np.asarray(src[:, col], np.uint64) * np.uint64(int(w) % P)) % np.uint64(P)
    return acc


def byte_chip_for(chips):
    """ByteChip::generate_trace (bytes/trace.rs:46-67) for a shard of `chips` (each with canonical rows in .canon): every
    byte lookup the chips SEND -- read off their AIRs' recorded lookups, row by row -- is looked up in the table
    (U16Range by value, everything else by (b, c)), checked against it, and counted in the multiplicity column of its
    opcode.  Returns the Byte chip (preprocessed = the table)."""
    table = byte_table()
    mult = np.zeros((1 << 16, 10), np.uint64)
    col_of = {0: 2, 1: 3, 2: 4, 3: 6, 5: 7, 6: 9, 7: 10, 9: 5}                    # opcode -> table column of a1
    for ch in chips:
        air = _air(ch.air)
        prep, main = ch.canon
        h = main.shape[0]
        for l in air.sends:
            if l["kind"] != 4:
                continue
            m = _lf_values(l["mult"], prep, main, h)
            op, a1, a2, b, c = (_lf_values(v, prep, main, h) for v in l["values"])
            live = m != 0
            assert (m[live] < (1 << 20)).all(), f"{ch.name}: byte lookup with a negative multiplicity"
            row = np.where(op == 8, a1, (b << np.uint64(8)) + c)
            assert (row[live] < (1 << 16)).all() and (op[live] < 10).all(), f"{ch.name}: byte lookup outside the table"
            r, o = row[live].astype(np.int64), op[live].astype(np.int64)
            for code, col in col_of.items():                                       # the looked-up value must be the table's
                sel = o == code
                assert (table[r[sel], col] == a1[live][sel]).all(), f"{ch.name}: byte lookup {code} disagrees with the table"
            sel = o == 5
            assert (table[r[sel], 8] == a2[live][sel]).all(), f"{ch.name}: ShrCarry carry disagrees with the table"
            np.add.at(mult, (r, o), m[live])
    byte = Chip("Byte", "Byte", M(mult % np.uint64(P)), preprocessed=M(table))
    byte.canon = (table, mult % np.uint64(P))
    return byte


def program_chip(log_n, seed=29, fill=0.75, name="Program"):
    """ProgramChip::generate_preprocessed_trace / generate_trace (program/mod.rs:56-81 and the multiplicity count): a random
    program of `fill * 2^log_n` instructions at pc = 4 i and how often each was executed"""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    real = max(1, int(n * fill))
    prep = np.zeros((n, 14), np.uint64)
    prep[:real, 0] = 4 * np.arange(real)
    prep[:real, 1] = rng.integers(0, 56, real)                                     # opcode
    op_a = rng.integers(0, 34, real)
    prep[:real, 2] = op_a
    prep[:real, 3:7], prep[:real, 7:11] = _bytes(rng.integers(0, 1 << 32, real, dtype=np.uint64)), _bytes(rng.integers(0, 1 << 32, real, dtype=np.uint64))
    prep[:real, 11], prep[:real, 12], prep[:real, 13] = op_a == 0, rng.integers(0, 2, real), rng.integers(0, 2, real)
    main = np.zeros((n, 1), np.uint64)
    main[:real, 0] = rng.integers(0, 1000, real)
    ch = Chip(name, "Program", M(main), preprocessed=M(prep))
    ch.canon = (prep, main)
    return ch


def syscall_chip(log_n, kind="Core", seed=30, fill=0.75):
    """SyscallChip::generate_trace (syscall/chip.rs:100-170): one row per syscall event: shard, clk, syscall_id, arg1, arg2"""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    real = max(1, int(n * fill))
    t = np.zeros((n, 6), np.uint64)
    t[:real, 0] = 1 + rng.integers(0, 4, real)
    t[:real, 1] = 5 * np.arange(real) + 24
    t[:real, 2] = rng.integers(0, 0x130, real)
    t[:real, 3], t[:real, 4] = 4 * rng.integers(0, 1 << 20, real), 4 * rng.integers(0, 1 << 20, real)
    t[:real, 5] = 1
    ch = Chip("Syscall" + kind, "Syscall" + kind, M(t))
    ch.canon = (None, t)
    return ch


def memory_local_chip(log_n, seed=31, fill=0.75):
    """MemoryLocalChip::generate_trace (memory/local.rs:125-175): four MemoryLocalEvents per row (the last real row may be
    partly filled)"""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    cells = max(1, int(4 * n * fill) - 1)
    t = np.zeros((4 * n, 14), np.uint64)
    t[:cells, 0] = 4 * rng.permutation(1 << 22)[:cells]
    t[:cells, 1] = rng.integers(0, 3, cells)
    t[:cells, 2] = 3
    t[:cells, 3], t[:cells, 4] = rng.integers(0, 1 << 20, cells), rng.integers(1 << 20, 1 << 21, cells)
    t[:cells, 5:9], t[:cells, 9:13] = _bytes(rng.integers(0, 1 << 32, cells, dtype=np.uint64)), _bytes(rng.integers(0, 1 << 32, cells, dtype=np.uint64))
    t[:cells, 13] = 1
    t = t.reshape(n, 56)
    ch = Chip("MemoryLocal", "MemoryLocal", M(t))
    ch.canon = (None, t)
    return ch


def shift_right_events(log_n, seed=32, fill=0.75):
    """random SRL / SRA / ROR AluEvents; half of the shift amounts carry garbage above bit 4 (MIPS takes c mod 32)"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    c = np.where(rng.integers(0, 2, real) == 0, c & np.uint64(31), c)
    op = rng.integers(0, 3, real)                                                  # 0 SRL, 1 SRA, 2 ROR
    sh = c & np.uint64(31)
    srl = b >> sh
    sra = (b.astype(np.uint32).view(np.int32).astype(np.int64) >> sh.astype(np.int64)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    ror = ((b >> sh) | (b << (np.uint64(32) - sh))) & np.uint64(0xFFFFFFFF)
    a = np.select([op == 0, op == 1], [srl, sra], ror)
    return _alu_event_array(pc, 10 + op, a, b, c), n


def shift_right_rows(events, n):
    """ShiftRightChip::event_to_row (alu/sr/mod.rs:153-247) and the padding row of generate_trace (:108-111)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 71), np.uint64)
    t[:, 14], t[:, 22] = 1, 1
    a, b, c = ev[:, 4], ev[:, 5], ev[:, 6]
    opc = ev[:, 2] & np.uint64(0xFF)
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10], t[:real, 10:14] = _bytes(a), _bytes(b), _bytes(c)
    nbits, nbytes = (c & np.uint64(31)) % np.uint64(8), (c & np.uint64(31)) // np.uint64(8)
    for i in range(8):
        t[:real, 14 + i] = nbits == i
        t[:real, 59 + i] = (c >> np.uint64(i)) & np.uint64(1)
    for i in range(4):
        t[:real, 22 + i] = nbytes == i
    msb = (b >> np.uint64(31)) & np.uint64(1)
    t[:real, 58] = msb
    bb = _bytes(b)
    hi = np.where((opc == 11)[:, None], (msb * np.uint64(0xFF))[:, None] * np.ones(4, np.uint64)[None, :],
                  np.where((opc == 12)[:, None], bb, 0))
    ext = np.concatenate([bb, hi], axis=1).astype(np.uint64)                       # the 8-byte extension of b
    byte_res = np.zeros((real, 8), np.uint64)
    for i in range(8):
        src = i + nbytes.astype(np.int64)
        ok = src < 8
        byte_res[ok, i] = ext[np.nonzero(ok)[0], src[ok]]
    k = nbits
    shifted = byte_res >> k[:, None]
    carry = np.where((k == 0)[:, None], 0, ((byte_res << (np.uint64(8) - k)[:, None]) & np.uint64(0xFF)) >> (np.uint64(8) - k)[:, None])
    mult = np.uint64(1) << (np.uint64(8) - k)
    bit_res = np.zeros((real, 8), np.uint64)
    last = np.zeros(real, np.uint64)
    for i in reversed(range(8)):
        bit_res[:, i] = (shifted[:, i] + last * mult) & np.uint64(0xFF)
        last = carry[:, i]
    assert (bit_res[:, :4] == _bytes(a)).all(), "ShiftRight filler disagrees with the event's a"
    t[:real, 26:34], t[:real, 34:42], t[:real, 42:50], t[:real, 50:58] = byte_res, bit_res, carry, shifted
    t[:real, 67], t[:real, 68], t[:real, 69], t[:real, 70] = opc == 10, opc == 12, opc == 11, 1
    return t


def shift_right_chip(log_n, seed=32, fill=0.75, name="ShiftRight", device=False):
    ev, n = shift_right_events(log_n, seed, fill)
    if device:                                                    # rows filled on the GPU from the AluEvents (zk_tracegen_alu)
        return Chip(name, "ShiftRight", None, events=ev, tracegen="ShiftRight", rows=n)
    t = shift_right_rows(ev, n)
    ch = Chip(name, "ShiftRight", M(t))
    ch.canon, ch.events = (None, t), ev
    return ch


def mul_events(log_n, seed=33, fill=0.75):
    """random MUL / MULT / MULTU events (CompAluEvent: pc, next_pc, opcode, hi, a, b, c, shard, clk, hi_record_is_real and
    the HI register's write record): columns pc, next_pc, opcode, hi, a, b, c, hi_record_is_real, shard, clk, prev_hi,
    prev_shard, prev_clk.  MUL keeps only the low word; MULT / MULTU write HI, with the register write checked in about
    half of them (from this shard or an earlier one)."""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    small = rng.integers(0, 4, real) == 0
    b = np.where(small, b & np.uint64(0xFF), b)
    op = rng.integers(0, 3, real)                                                  # 0 MUL, 1 MULT, 2 MULTU
    sb = np.where(op == 1, b.astype(np.uint32).view(np.int32).astype(np.int64), b.astype(np.int64))
    sc = np.where(op == 1, c.astype(np.uint32).view(np.int32).astype(np.int64), c.astype(np.int64))
    prod = [(int(x) * int(y)) & 0xFFFFFFFFFFFFFFFF for x, y in zip(sb, sc)]
    a = np.array([p & 0xFFFFFFFF for p in prod], np.uint64)
    hi = np.where(op == 0, 0, np.array([p >> 32 for p in prod], np.uint64))
    hi_real = (op != 0) & (rng.integers(0, 2, real) == 1)
    shard = np.where(hi_real, 3, 0)
    clk = np.where(hi_real, 24 + 5 * np.arange(real), 0)
    same = rng.integers(0, 2, real) == 1
    prev_shard = np.where(hi_real, np.where(same, 3, rng.integers(0, 3, real)), 0)
    prev_clk = np.where(hi_real, np.where(same, clk - rng.integers(0, 20, real), rng.integers(0, 1 << 20, real)), 0)
    prev_hi = np.where(hi_real, rng.integers(0, 1 << 32, real, dtype=np.uint64), 0)
    ev = np.stack([pc, (pc + 4) % P, 2 + op, hi, a, b, c, hi_real, shard, clk, prev_hi, prev_shard, prev_clk], axis=1)
    return ev.astype(np.uint64), n


def mul_rows(events, n):
    """MulChip::event_to_row (alu/mul/mod.rs:221-325) with MemoryReadWriteCols::populate_write
    (memory/consistency/trace.rs:43-105)"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 58), np.uint64)
    opc, hi, a, b, c, hi_real, shard, clk, prev_hi, prev_shard, prev_clk = (ev[:, k] for k in range(2, 13))
    t[:real, 0], t[:real, 1] = ev[:, 0], ev[:, 1]
    t[:real, 2:6], t[:real, 6:10], t[:real, 10:14], t[:real, 14:18] = _bytes(hi), _bytes(a), _bytes(b), _bytes(c)
    b_msb, c_msb = (b >> np.uint64(31)) & np.uint64(1), (c >> np.uint64(31)) & np.uint64(1)
    b_sx, c_sx = (opc == 3) & (b_msb == 1), (opc == 3) & (c_msb == 1)
    be = np.concatenate([_bytes(b), np.where(b_sx[:, None], 0xFF, 0) * np.ones((real, 4), np.uint64)], axis=1).astype(np.uint64)
    ce = np.concatenate([_bytes(c), np.where(c_sx[:, None], 0xFF, 0) * np.ones((real, 4), np.uint64)], axis=1).astype(np.uint64)
    prod = np.zeros((real, 8), np.uint64)
    for i in range(8):
        for j in range(8 - i):
            prod[:, i + j] += be[:, i] * ce[:, j]
    for i in range(8):
        carry = prod[:, i] >> np.uint64(8)
        prod[:, i] &= np.uint64(0xFF)
        if i + 1 < 8:
            prod[:, i + 1] += carry
        t[:real, 18 + i] = carry
    t[:real, 26:34] = prod
    assert (prod[:, :4] == _bytes(a)).all() and (prod[opc != 2][:, 4:] == _bytes(hi[opc != 2])).all(), "Mul filler disagrees with the events"
    t[:real, 34], t[:real, 35], t[:real, 36], t[:real, 37] = b_msb, c_msb, b_sx, c_sx
    t[:real, 38], t[:real, 39], t[:real, 40], t[:real, 41] = opc == 2, opc == 3, opc == 4, 1
    live = hi_real == 1
    t[:real, 42:46] = np.where(live[:, None], _bytes(prev_hi), 0)
    t[:real, 46:50] = np.where(live[:, None], _bytes(hi), 0)
    compare = live & (prev_shard == shard)
    cur = np.where(compare, clk + 4, shard)                                        # the HI access happens at clk + 4
    prv = np.where(compare, prev_clk, prev_shard)
    diff = np.where(live, cur - prv - 1, 0)
    t[:real, 50], t[:real, 51], t[:real, 52] = np.where(live, prev_shard, 0), np.where(live, prev_clk, 0), compare
    t[:real, 53], t[:real, 54] = diff & np.uint64(0xFFFF), (diff >> np.uint64(16)) & np.uint64(0xFF)
    t[:real, 55], t[:real, 56], t[:real, 57] = hi_real, shard, clk
    return t


def mul_chip(log_n, seed=33, fill=0.75, name="Mul"):
    ev, n = mul_events(log_n, seed, fill)
    t = mul_rows(ev, n)
    ch = Chip(name, "Mul", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


def _quotient_remainder(b, c, signed):
    """get_quotient_and_remainder (crates/core/executor/src/utils.rs:33-43): c = 0 gives (0xFFFFFFFF, b); signed division
    truncates toward zero and wraps on i32::MIN / -1"""
    M32 = 0xFFFFFFFF
    if c == 0:
        return M32, b
    if not signed:
        return b // c, b % c
    sb, sc = (b ^ 0x80000000) - 0x80000000, (c ^ 0x80000000) - 0x80000000
    q = abs(sb) // abs(sc)
    if (sb < 0) != (sc < 0):
        q = -q
    r = sb - q * sc
    return q & M32, r & M32


def div_rem_events(log_n, seed=34, fill=0.75):
    """random DIV / DIVU / MOD / MODU events, including division by zero, i32::MIN / -1, small divisors and negative
    operands: columns pc, next_pc, opcode, b, c, shard, clk, prev_hi, prev_shard, prev_clk (the HI register's write record
    of DIV / DIVU; zero for MOD / MODU)"""
    n, real, rng, pc, b, c = _events(log_n, seed, fill)
    pick = rng.integers(0, 8, real)
    c = np.where(pick == 0, 0, np.where(pick == 1, c & np.uint64(0xFF), np.where(pick == 2, np.uint64(0xFFFFFFFF), c)))
    b = np.where(pick == 2, np.where(rng.integers(0, 2, real) == 0, np.uint64(0x80000000), b), b)
    op = rng.integers(0, 4, real)                                                  # DIV 5, DIVU 6, MOD 7, MODU 8
    has_hi = op < 2
    shard = np.where(has_hi, 3, 0)
    clk = np.where(has_hi, 24 + 5 * np.arange(real), 0)
    same = rng.integers(0, 2, real) == 1
    prev_shard = np.where(has_hi, np.where(same, 3, rng.integers(0, 3, real)), 0)
    prev_clk = np.where(has_hi, np.where(same, clk - rng.integers(0, 20, real), rng.integers(0, 1 << 20, real)), 0)
    prev_hi = np.where(has_hi, rng.integers(0, 1 << 32, real, dtype=np.uint64), 0)
    ev = np.stack([pc, (pc + 4) % P, 5 + op, b, c, shard, clk, prev_hi, prev_shard, prev_clk], axis=1)
    return ev.astype(np.uint64), n


def div_rem_rows(events, n):
    """DivRemChip::generate_trace (alu/divrem/mod.rs:112-340); padding rows are zero"""
    ev = np.asarray(events, np.uint64)
    real = len(ev)
    t = np.zeros((n, 106), np.uint64)
    M32 = 0xFFFFFFFF

    def word(v):
        return [(int(v) >> (8 * k)) & 0xFF for k in range(4)]

    def is_zero_word(vals):                                            # IsZeroWordOperation::populate_from_field_element
        inv = [pow(int(x), P - 2, P) if x else 0 for x in vals]
        z = [int(x == 0) for x in vals]
        out = []
        for i in range(4):
            out += [inv[i], z[i]]
        return out + [z[0] * z[1], z[2] * z[3], int(all(z))]
    for r in range(real):
        pc, next_pc, opc, b, c, shard, clk, prev_hi, prev_shard, prev_clk = (int(x) for x in ev[r])
        signed = opc in (5, 7)
        q, rem = _quotient_remainder(b, c, signed)
        row = t[r]
        row[0], row[1] = pc, next_pc
        row[2:6], row[6:10], row[10:14], row[14:18] = word(b), word(c), word(q), word(rem)
        sc, srem = (c ^ 0x80000000) - 0x80000000, (rem ^ 0x80000000) - 0x80000000
        abs_rem, abs_c = (abs(srem) & M32, abs(sc) & M32) if signed else (rem, c)
        row[18:22], row[22:26], row[26:30] = word(abs_rem), word(abs_c), word(max(1, abs_c))
        sq = (q ^ 0x80000000) - 0x80000000
        ctq = ((sq * sc) if signed else (q * c)) & 0xFFFFFFFFFFFFFFFF
        ctq_b = [(ctq >> (8 * k)) & 0xFF for k in range(8)]
        rem64 = (srem if signed else rem) & 0xFFFFFFFFFFFFFFFF
        rem_b = [(rem64 >> (8 * k)) & 0xFF for k in range(8)]
        carry = 0
        for i in range(8):
            carry = (ctq_b[i] + rem_b[i] + carry) >> 8
            row[38 + i] = carry
        row[30:38] = ctq_b
        row[46:57] = is_zero_word(word(c))
        row[57], row[58], row[59], row[60] = opc == 5, opc == 6, opc == 7, opc == 8
        row[61] = signed and b == 0x80000000 and c == M32
        row[62:73] = is_zero_word([(x - k) % P for x, k in zip(word(b), (0, 0, 0, 0x80))])
        row[73:84] = is_zero_word([(x - 0xFF) % P for x in word(c)])
        b_msb, rem_msb, c_msb = b >> 31, rem >> 31, c >> 31
        row[84], row[85], row[86] = b_msb, rem_msb, c_msb
        if signed:
            row[87], row[88], row[89] = b_msb, rem_msb, c_msb
        row[90] = int(c != 0)
        if opc in (5, 6):
            compare = prev_shard == shard
            diff = ((clk + 4 if compare else shard) - (prev_clk if compare else prev_shard) - 1) & M32
            row[91:95], row[95:99] = word(prev_hi), word(rem)
            row[99:104] = [prev_shard, prev_clk, int(compare), diff & 0xFFFF, (diff >> 16) & 0xFF]
            row[104], row[105] = shard, clk
    return t


def div_rem_chip(log_n, seed=34, fill=0.75, name="DivRem"):
    ev, n = div_rem_events(log_n, seed, fill)
    t = div_rem_rows(ev, n)
    ch = Chip(name, "DivRem", M(t), local_only=True)
    ch.canon, ch.events = (None, t), ev
    return ch


# ------------------------------------------------------------------------------------------------------------------
# Recursion chip Poseidon2WideDeg3 / Deg9 (library.poseidon2_wide): random permutation inputs and memory addresses.
# The main trace is ALWAYS produced on the device from the 16-word inputs (zk_tracegen_poseidon2_wide); the
# preprocessed trace (addresses and multiplicities of the program's Poseidon2 instructions) is a plain scatter.
# ------------------------------------------------------------------------------------------------------------------
def poseidon2_wide_events(log_n, seed=31, fill=0.75):
    """(inputs [real, 16] Montgomery, instrs [real, 48] Montgomery: input addrs, output addrs, mults; padded height)"""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    real = max(1, int(n * fill)) if n > 1 else 1
    inputs = to_monty(rng.integers(0, P, (real, 16), dtype=np.uint64))
    addrs = rng.integers(0, 1 << 24, (real, 32), dtype=np.uint64)
    mults = rng.integers(0, 4, (real, 16), dtype=np.uint64)
    return inputs, to_monty(np.concatenate([addrs, mults], axis=1)), n


def poseidon2_wide_prep_rows(instrs, n):
    """Poseidon2WideChip::generate_preprocessed_trace (chips/poseidon2_wide/trace.rs:183-216), Montgomery in / out"""
    ins = np.asarray(instrs, np.uint32)
    t = np.zeros((n, 49), np.uint32)
    real = len(ins)
    t[:real, 0:16] = ins[:, 0:16]
    t[:real, 16:48:2] = ins[:, 16:32]
    t[:real, 17:48:2] = ins[:, 32:48]
    t[:real, 48] = to_monty(np.array([P - 1], np.uint64))[0]
    return t


def poseidon2_wide_chip(log_n, degree=3, seed=31, fill=0.75, name=None):
    inputs, instrs, n = poseidon2_wide_events(log_n, seed, fill)
    air = f"Poseidon2WideDeg{degree}"
    return Chip(name or air, air, None, preprocessed=poseidon2_wide_prep_rows(instrs, n), local_only=True,
                log_quotient_degree=1 if degree == 3 else 3, events=inputs, tracegen=air, rows=n)


# ------------------------------------------------------------------------------------------------------------------
# A toy recursion PROGRAM over three real RecursionAir chips whose memory bus balances, so that the shard's cumulative
# sum is zero and a verifier accepts the proof completely: MemoryConst writes constants c_j to addresses j; Poseidon2
# permutation k reads addresses k .. k + 15 and writes its 16 outputs to fresh addresses (never read: multiplicity 0);
# BaseAlu operation t reads two constants and writes its result to a fresh address.  Every constant's write
# multiplicity is the number of times it is read (crates/recursion/core/src/chips/mem/mod.rs:14-22: positive = write).
# ------------------------------------------------------------------------------------------------------------------
def _inv_mod(x):
    return np.array([pow(int(v), P - 2, P) for v in x], np.uint64)


def recursion_program_chips(log_p2=6, log_alu=5, log_mem=6, degree=3, seed=41, fill=0.75, names=None, log_var=None,
                            log_ext=None, log_sel=None, log_bf=None, log_exp=None, pv=False):
    """(MemoryConst, BaseAlu, Poseidon2WideDeg<degree>) chips of the toy program; the Poseidon2 rows are filled on the
    device from the permutation inputs, the other two chips' traces are a few columns and stay host arrays.
    With log_var / log_ext / log_sel also (MemoryVar, ExtAlu, Select): MemoryVar writes the two extension operands of
    every ExtAlu operation (read once each), Select reads a bit constant (addresses n_const - 2 / n_const - 1 hold 0 / 1)
    and two other constants.
    With log_bf / log_exp / pv also (BatchFRI, ExpReverseBitsLen, PublicValues) -- with them the program runs on all NINE
    chips of the reference's compress machine (recursion/core/src/machine.rs:112-128): a BatchFRI instruction of up to
    four (alpha_pow, p_at_z, p_at_x) triples reads its extension operands from MemoryVar and p_at_x from the constants,
    and its accumulator (written once, multiplicity 1) is the first operand of an ExtAlu operation; an
    ExpReverseBitsLen instruction of up to five bits reads its base and its bits (the two bit constants) and its result
    is the first operand of a BaseAlu operation; PublicValues reads eight constants, which are the digest of the
    shard's RecursionPublicValues (`pv_digest` on the chip, public_values_for puts it at 223..231)."""
    rng = np.random.default_rng(seed)
    n_perm = max(1, int((1 << log_p2) * fill))
    n_alu = max(1, int((4 << log_alu) * fill))
    n_const = n_perm + 15 + 2
    assert n_const <= (2 << log_mem), "MemoryConst too short for the constants of this program"
    consts = rng.integers(1, P, n_const, dtype=np.uint64)              # non-zero: BaseAlu divides by them
    consts[-2:] = (0, 1)                                               # the two bit constants (never ALU operands)
    reads = np.zeros(n_const, np.uint64)
    # Poseidon2 instructions
    win = np.arange(n_perm)[:, None] + np.arange(16)[None, :]          # addresses read by permutation k
    np.add.at(reads, win.ravel(), 1)
    out_base = 1 << 22
    p2_instrs = np.concatenate([win, out_base + 16 * np.arange(n_perm)[:, None] + np.arange(16)[None, :],
                                np.zeros((n_perm, 16), np.int64)], axis=1).astype(np.uint64)
    p2_inputs = to_monty(consts[win])
    # BaseAlu instructions
    a1 = rng.integers(0, n_const - 2, n_alu)
    a2 = rng.integers(0, n_const - 2, n_alu)
    exp_chip = None
    a1_addr = a1.astype(np.int64)
    x = consts[a1]
    from_const = np.ones(n_alu, bool)
    if log_exp is not None:
        # ExpReverseBitsLenChip (chips/exp_reverse_bits.rs:100-141 preprocessed, :216-246 main): one row per bit
        rows_x = 1 << log_exp
        n_xr = max(1, int(rows_x * fill))
        starts = np.arange(0, n_xr, 5)
        n_xi = len(starts)
        assert n_xi <= n_alu, "BaseAlu too short to consume the ExpReverseBitsLen results"
        instr_of = np.arange(n_xr) // 5
        it = np.arange(n_xr) % 5
        last = np.zeros(n_xr, bool)
        last[np.minimum(starts + 4, n_xr - 1)] = True
        base_addr = rng.integers(0, n_const - 2, n_xi)
        bits = rng.integers(0, 2, n_xr)
        np.add.at(reads, base_addr, 1)                                 # x is read once, on the first row
        np.add.at(reads, n_const - 2 + bits, 1)                        # every row reads its bit
        xb = consts[base_addr][instr_of]
        mult_col = np.where(bits == 1, xb, 1).astype(np.uint64)
        exp_main = np.zeros((rows_x, 7), np.uint64)
        accum = np.uint64(1)
        results = np.zeros(n_xi, np.uint64)
        for r in range(n_xr):
            prev = np.uint64(1) if it[r] == 0 else accum
            pas = prev * prev % P
            accum = pas * mult_col[r] % P
            exp_main[r] = (xb[r], bits[r], pas, accum, accum, accum * accum % P, mult_col[r])
            if last[r]:
                results[instr_of[r]] = accum
        exp_prep = np.zeros((rows_x, 10), np.uint64)
        res_addr = (1 << 29) + np.arange(n_xi)
        exp_prep[:n_xr, 0] = base_addr[instr_of]
        exp_prep[:n_xr, 1] = np.where(it == 0, P - 1, 0)               # -1: read (chips/mem/mod.rs:14-22)
        exp_prep[:n_xr, 2] = n_const - 2 + bits
        exp_prep[:n_xr, 3] = P - 1
        exp_prep[:n_xr, 4] = res_addr[instr_of]
        exp_prep[:n_xr, 5] = last                                      # the result is written once, read by BaseAlu
        exp_prep[:n_xr, 6], exp_prep[:n_xr, 7], exp_prep[:n_xr, 8], exp_prep[:n_xr, 9] = it, it == 0, last, 1
        exp_chip = Chip("ExpReverseBitsLen", "ExpReverseBitsLen", M(exp_main), preprocessed=M(exp_prep))
        exp_chip.canon = (exp_prep, exp_main)
        a1_addr[:n_xi], x[:n_xi], from_const[:n_xi] = res_addr, results, False
    np.add.at(reads, a1[from_const], 1)
    np.add.at(reads, a2, 1)
    op = rng.integers(0, 4, n_alu)                                     # 0 add, 1 sub, 2 mul, 3 div
    y = consts[a2]
    res = np.select([op == 0, op == 1, op == 2], [(x + y) % P, (x + P - y) % P, x * y % P], x * _inv_mod(y) % P)
    alu_rows = 1 << log_alu
    alu_main = np.zeros((alu_rows * 4, 3), np.uint64)
    alu_main[:n_alu] = np.stack([res, x, y], axis=1)                   # BaseAluIo {out, in1, in2}
    alu_prep = np.zeros((alu_rows * 4, 8), np.uint64)
    alu_prep[:n_alu, 0] = (1 << 23) + np.arange(n_alu)                 # fresh output addresses
    alu_prep[:n_alu, 1], alu_prep[:n_alu, 2] = a1_addr, a2
    alu_prep[np.arange(n_alu), 3 + op] = 1                             # is_add / is_sub / is_mul / is_div; mult = 0
    # MemoryConst: entries (value block, addr, mult), two per row
    mem_rows = 1 << log_mem
    mem_prep = np.zeros((mem_rows * 2, 6), np.uint64)
    extra = []
    if log_sel is not None:
        # SelectChip: out1 = bit ? in2 : in1, out2 = bit ? in1 : in2 (chips/select.rs:242-249); SelectIo order
        # {bit, out1, out2, in1, in2}; outputs go to fresh addresses nobody reads
        rows_s = 1 << log_sel
        n_sel = max(1, int(rows_s * fill))
        bit = rng.integers(0, 2, n_sel)
        s1, s2 = rng.integers(0, n_const - 2, n_sel), rng.integers(0, n_const - 2, n_sel)
        for a in (n_const - 2 + bit, s1, s2):
            np.add.at(reads, a, 1)
        v1, v2 = consts[s1], consts[s2]
        sel_main = np.zeros((rows_s, 5), np.uint64)
        sel_main[:n_sel] = np.stack([bit.astype(np.uint64), np.where(bit == 1, v2, v1), np.where(bit == 1, v1, v2), v1, v2], axis=1)
        sel_prep = np.zeros((rows_s, 8), np.uint64)
        sel_prep[:n_sel, 0] = 1
        sel_prep[:n_sel, 1], sel_prep[:n_sel, 4], sel_prep[:n_sel, 5] = n_const - 2 + bit, s1, s2
        sel_prep[:n_sel, 2] = (1 << 25) + 2 * np.arange(n_sel)
        sel_prep[:n_sel, 3] = (1 << 25) + 2 * np.arange(n_sel) + 1
        sel = Chip("Select", "Select", M(sel_main), preprocessed=M(sel_prep), local_only=True)
        sel.canon = (sel_prep, sel_main)
        extra.append(sel)
    if log_ext is not None:
        rows_e = 1 << log_ext
        n_ext = max(1, int(4 * rows_e * fill))
        assert 2 * n_ext <= (2 << log_var), "MemoryVar too short for the ExtAlu operands"
        eop = rng.integers(0, 4, n_ext)                                # 0 add, 1 sub, 2 mul, 3 div
        u = rng.integers(0, P, (n_ext, 4), dtype=np.uint64)
        v = rng.integers(0, P, (n_ext, 4), dtype=np.uint64)
        bf = None
        if log_bf is not None:
            # BatchFRIChip (chips/batch_fri.rs:90-118 preprocessed, :196-208 main): instructions of <= 4 triples
            rows_b = 1 << log_bf
            n_br = max(1, int(rows_b * fill))
            b_instr = np.arange(n_br) // 4
            n_bi = int(b_instr[-1]) + 1
            assert n_bi <= n_ext, "ExtAlu too short to consume the BatchFRI accumulators"
            b_end = np.zeros(n_br, bool)
            b_end[np.minimum(np.arange(0, n_br, 4) + 3, n_br - 1)] = True
            b_alpha = rng.integers(0, P, (n_br, 4), dtype=np.uint64)
            b_z = rng.integers(0, P, (n_br, 4), dtype=np.uint64)
            b_xaddr = rng.integers(0, n_const - 2, n_br)
            np.add.at(reads, b_xaddr, 1)
            bf = (rows_b, n_br, b_instr, n_bi, b_end, b_alpha, b_z, b_xaddr)
            eop[:n_bi] = 2 * rng.integers(0, 2, n_bi)                  # add or mul: out follows from in1 = acc

        def emul(a, c):                                                # F_p[X] / (X^4 - 3)
            r = np.zeros_like(a)
            for i in range(4):
                for j in range(4):
                    t = a[:, i] * c[:, j] % P
                    r[:, (i + j) % 4] = (r[:, (i + j) % 4] + (3 * t if i + j >= 4 else t)) % P
            return r
        # add: out = in1 + in2; sub: in1 = in2 + out; mul: out = in1 * in2; div: in1 = in2 * out (so out = in1 / in2)
        in2 = v
        out = np.where((eop == 0)[:, None], (u + v) % P, np.where((eop == 2)[:, None], emul(u, v), u))
        in1 = np.where((eop == 1)[:, None], (v + u) % P, np.where((eop == 3)[:, None], emul(v, u), u))
        ext_main = np.zeros((4 * rows_e, 12), np.uint64)
        ext_main[:n_ext] = np.concatenate([out, in1, in2], axis=1)
        ext_prep = np.zeros((4 * rows_e, 8), np.uint64)
        var_base = 1 << 24
        ext_prep[:n_ext, 0] = (1 << 26) + np.arange(n_ext)
        ext_prep[:n_ext, 1], ext_prep[:n_ext, 2] = var_base + 2 * np.arange(n_ext), var_base + 2 * np.arange(n_ext) + 1
        ext_prep[np.arange(n_ext), 3 + eop] = 1
        rows_v = 1 << log_var
        var_main = np.zeros((2 * rows_v, 4), np.uint64)
        var_main[0:2 * n_ext:2], var_main[1:2 * n_ext:2] = in1, in2
        var_prep = np.zeros((2 * rows_v, 2), np.uint64)
        var_prep[:2 * n_ext, 0], var_prep[:2 * n_ext, 1] = var_base + np.arange(2 * n_ext), 1
        if bf is not None:
            rows_b, n_br, b_instr, n_bi, b_end, b_alpha, b_z, b_xaddr = bf
            assert 2 * n_ext + 2 * n_br <= 2 * rows_v, "MemoryVar too short for the BatchFRI operands"
            b_x = consts[b_xaddr]
            zx = b_z.copy()
            zx[:, 0] = (zx[:, 0] + P - b_x) % P
            terms = emul(b_alpha, zx)
            b_acc = np.zeros((n_br, 4), np.uint64)
            for r in range(n_br):
                b_acc[r] = terms[r] if r % 4 == 0 else (b_acc[r - 1] + terms[r]) % P
            acc_addr = (1 << 27) + np.arange(n_bi)
            bf_main = np.zeros((rows_b, 13), np.uint64)
            bf_main[:n_br] = np.concatenate([b_acc, b_alpha, b_z, b_x[:, None]], axis=1)
            bf_prep = np.zeros((rows_b, 6), np.uint64)
            opnd = (1 << 28) + np.arange(2 * n_br)
            bf_prep[:n_br] = np.stack([np.ones(n_br, np.uint64), b_end.astype(np.uint64), acc_addr[b_instr].astype(np.uint64),
                                       opnd[0::2].astype(np.uint64), opnd[1::2].astype(np.uint64), b_xaddr.astype(np.uint64)], axis=1)
            o = 2 * n_ext
            var_main[o:o + 2 * n_br:2], var_main[o + 1:o + 2 * n_br:2] = b_alpha, b_z
            var_prep[o:o + 2 * n_br, 0], var_prep[o:o + 2 * n_br, 1] = opnd, 1
            # the accumulators are the first operands of the first n_bi ExtAlu operations: their MemoryVar entries
            # stay in the table with multiplicity 0 (written, never read)
            accs = b_acc[b_end]
            in1[:n_bi] = accs
            out[:n_bi] = np.where((eop[:n_bi] == 0)[:, None], (accs + in2[:n_bi]) % P, emul(accs, in2[:n_bi]))
            ext_main[:n_ext] = np.concatenate([out, in1, in2], axis=1)
            ext_prep[:n_bi, 1] = acc_addr
            var_prep[0:2 * n_bi:2, 1] = 0
            bfc = Chip("BatchFRI", "BatchFRI", M(bf_main), preprocessed=M(bf_prep))
            bfc.canon = (bf_prep, bf_main)
        var = Chip("MemoryVar", "MemoryVar", M(var_main.reshape(rows_v, 8)), preprocessed=M(var_prep.reshape(rows_v, 4)),
                   local_only=True)
        var.canon = (var_prep.reshape(rows_v, 4), var_main.reshape(rows_v, 8))
        ext = Chip("ExtAlu", "ExtAlu", M(ext_main.reshape(rows_e, 48)), preprocessed=M(ext_prep.reshape(rows_e, 32)),
                   local_only=True)
        ext.canon = (ext_prep.reshape(rows_e, 32), ext_main.reshape(rows_e, 48))
        extra += [var, ext]
        if bf is not None:
            extra.append(bfc)
    else:
        assert log_bf is None, "BatchFRI needs MemoryVar and ExtAlu (log_var, log_ext)"
    if exp_chip is not None:
        extra.append(exp_chip)
    if pv:
        # PublicValuesChip (chips/public_values.rs:82-125): 16 rows, row i < 8 reads digest element i
        d_addr = rng.integers(0, n_const - 2, 8)
        np.add.at(reads, d_addr, 1)
        pv_prep = np.zeros((16, 10), np.uint64)
        pv_prep[np.arange(8), np.arange(8)] = 1
        pv_prep[:8, 8], pv_prep[:8, 9] = d_addr, P - 1
        pv_main = np.zeros((16, 1), np.uint64)
        pv_main[:8, 0] = consts[d_addr]
        pvc = Chip("PublicValues", "PublicValues", M(pv_main), preprocessed=M(pv_prep))
        pvc.canon = (pv_prep, pv_main)
        pvc.pv_digest = consts[d_addr].copy()
        extra.append(pvc)
    mem_prep[:n_const, 0], mem_prep[:n_const, 4], mem_prep[:n_const, 5] = consts, np.arange(n_const), reads
    names = names or ("MemoryConst", "BaseAlu", f"Poseidon2WideDeg{degree}")
    mem = Chip(names[0], "MemoryConst", np.zeros((mem_rows, 1), np.uint32), preprocessed=M(mem_prep.reshape(mem_rows, 12)),
               local_only=True)
    alu = Chip(names[1], "BaseAlu", M(alu_main.reshape(alu_rows, 12)), preprocessed=M(alu_prep.reshape(alu_rows, 32)),
               local_only=True)
    mem.canon = (mem_prep.reshape(mem_rows, 12), np.zeros((mem_rows, 1), np.uint64))
    alu.canon = (alu_prep.reshape(alu_rows, 32), alu_main.reshape(alu_rows, 12))
    air = f"Poseidon2WideDeg{degree}"
    p2 = Chip(names[2], air, None, preprocessed=poseidon2_wide_prep_rows(to_monty(p2_instrs), 1 << log_p2),
              local_only=True, log_quotient_degree=1 if degree == 3 else 3, events=p2_inputs, tracegen=air,
              rows=1 << log_p2)
    return [mem, alu, p2] + extra


def fri_fold_program_chips(log_ff=5, log_var=8, log_mem=5, seed=43, fill=0.75):
    """(MemoryConst, MemoryVar, FriFold): a toy program of FriFold instructions (crates/recursion/core/src/chips/
    fri_fold.rs:117-184 preprocessed, :245-275 main) of up to three rows each with a balanced memory bus.  x comes from
    the constants and z, alpha (read on the instruction's first row) and the four vector inputs of every row from
    MemoryVar; the outputs are written with multiplicity 0.  ro_output = ro_input + alpha_pow_input * q with
    p_at_x = p_at_z + q * (x - z), so no extension inversion is needed to build a valid row."""
    rng = np.random.default_rng(seed)
    rows = 1 << log_ff
    n = max(1, int(rows * fill))
    instr = np.arange(n) // 3
    first = np.arange(n) % 3 == 0
    n_i = int(instr[-1]) + 1
    n_const = n_i
    consts = rng.integers(1, P, n_const, dtype=np.uint64)
    ext = lambda k: rng.integers(0, P, (k, 4), dtype=np.uint64)

    def emul(a, c):
        r = np.zeros_like(a)
        for i in range(4):
            for j in range(4):
                t = a[:, i] * c[:, j] % P
                r[:, (i + j) % 4] = (r[:, (i + j) % 4] + (3 * t if i + j >= 4 else t)) % P
        return r
    z_i, alpha_i = ext(n_i), ext(n_i)
    z, alpha, x = z_i[instr], alpha_i[instr], consts[instr]
    ap_in, ro_in, p_at_z, q = ext(n), ext(n), ext(n), ext(n)
    x_minus_z = (P - z) % P
    x_minus_z[:, 0] = (x + P - z[:, 0]) % P
    p_at_x = (p_at_z + emul(q, x_minus_z)) % P
    ap_out = emul(ap_in, alpha)
    ro_out = (ro_in + emul(ap_in, q)) % P
    main = np.zeros((rows, 33), np.uint64)
    main[:n] = np.concatenate([z, alpha, x[:, None], p_at_x, p_at_z, ap_in, ro_in, ap_out, ro_out], axis=1)
    # MemoryVar entries: z and alpha per instruction, then (alpha_pow_input, ro_input, p_at_x, p_at_z) per row
    var_vals = np.concatenate([z_i, alpha_i, ap_in, ro_in, p_at_x, p_at_z])
    n_var = len(var_vals)
    rows_v = 1 << log_var
    assert n_var <= 2 * rows_v, "MemoryVar too short for the FriFold operands"
    base = 1 << 24
    a_z, a_alpha = base + np.arange(n_i), base + n_i + np.arange(n_i)
    a_vec = base + 2 * n_i + np.arange(4 * n).reshape(4, n)
    neg1 = P - 1
    prep = np.zeros((rows, 20), np.uint64)
    prep[:n, 0] = first
    prep[:n, 1], prep[:n, 2] = a_z[instr], np.where(first, neg1, 0)
    prep[:n, 3], prep[:n, 4] = a_alpha[instr], np.where(first, neg1, 0)
    prep[:n, 5], prep[:n, 6] = instr, np.where(first, neg1, 0)          # x: constant number `instr`
    for k in range(4):                                                  # alpha_pow_input, ro_input, p_at_x, p_at_z
        prep[:n, 7 + 2 * k], prep[:n, 8 + 2 * k] = a_vec[k], neg1
    prep[:n, 15], prep[:n, 17] = (1 << 26) + np.arange(n), (1 << 27) + np.arange(n)   # outputs, multiplicity 0
    prep[:n, 19] = 1
    var_main = np.zeros((2 * rows_v, 4), np.uint64)
    var_main[:n_var] = var_vals
    var_prep = np.zeros((2 * rows_v, 2), np.uint64)
    var_prep[:n_var, 0], var_prep[:n_var, 1] = base + np.arange(n_var), 1
    mem_rows = 1 << log_mem
    assert n_const <= 2 * mem_rows
    mem_prep = np.zeros((2 * mem_rows, 6), np.uint64)
    mem_prep[:n_const, 0], mem_prep[:n_const, 4], mem_prep[:n_const, 5] = consts, np.arange(n_const), 1
    mem = Chip("MemoryConst", "MemoryConst", np.zeros((mem_rows, 1), np.uint32), preprocessed=M(mem_prep.reshape(mem_rows, 12)),
               local_only=True)
    mem.canon = (mem_prep.reshape(mem_rows, 12), np.zeros((mem_rows, 1), np.uint64))
    var = Chip("MemoryVar", "MemoryVar", M(var_main.reshape(rows_v, 8)), preprocessed=M(var_prep.reshape(rows_v, 4)),
               local_only=True)
    var.canon = (var_prep.reshape(rows_v, 4), var_main.reshape(rows_v, 8))
    ff = Chip("FriFold", "FriFold", M(main), preprocessed=M(prep))
    ff.canon = (prep, main)
    return [mem, var, ff]


class _Fp:
    """numpy uint64 vectors mod P with + and * only, so that library._external_linear_layer / _internal_linear_layer
    (written for constraint expressions) also run on values"""

    def __init__(self, v):
        self.v = np.asarray(v, np.uint64) % np.uint64(P)

    def __add__(self, o):
        return _Fp((self.v + (o.v if isinstance(o, _Fp) else np.uint64(int(o) % P))) % np.uint64(P))

    __radd__ = __add__

    def __mul__(self, o):
        return _Fp(self.v * (o.v if isinstance(o, _Fp) else np.uint64(int(o) % P)) % np.uint64(P))

    __rmul__ = __mul__


def poseidon2_skinny_rows(inputs, n):
    """Poseidon2SkinnyChip::generate_trace (chips/poseidon2_skinny/trace.rs:77-130, populate_* :326-400), canonical
    values: 11 rows per permutation -- input, 4 external rounds, the internal-rounds row, 4 external rounds, output."""
    from .air import library as L
    ext_rc, int_rc = L._poseidon2_round_constants()
    inputs = np.asarray(inputs, np.uint64)
    k = len(inputs)
    assert 11 * k <= n
    t = np.zeros((n, 28), np.uint64)
    rows = np.arange(k) * 11
    st = [_Fp(inputs[:, i]) for i in range(16)]
    put = lambda i, s: t.__setitem__((rows + i, slice(0, 16)), np.stack([x.v for x in s], axis=1))
    put(0, st)
    st = L._external_linear_layer(st)
    for i in range(1, 10):
        put(i, st)
        if i != 5:
            r = i - 1 if i < 5 else i - 2
            st = L._external_linear_layer([(x + ext_rc[r][j]) * (x + ext_rc[r][j]) * (x + ext_rc[r][j]) for j, x in enumerate(st)])
        else:
            for r in range(13):
                x = st[0] + int_rc[r]
                st[0] = x * x * x
                st = L._internal_linear_layer(st)
                if r < 12:
                    t[rows + 5, 16 + r] = st[0].v
    put(10, st)
    return t


def poseidon2_skinny_prep_rows(in_addrs, out_addrs, out_mults, n):
    """generate_preprocessed_trace (trace.rs:184-251), canonical: memory accesses on the input (mult -1) and output rows,
    round flags, and the round constants each row's constraints use (external: RC[round][0..16]; internal row:
    RC[4 + j][0] for j < 16, i.e. the 13 internal constants followed by column 0 of external rows 17..19)"""
    from .air import library as L
    ext_rc, int_rc = L._poseidon2_round_constants()
    k = len(in_addrs)
    t = np.zeros((n, 51), np.uint64)
    rows = np.arange(k) * 11
    t[rows, 0:32:2], t[rows, 1:32:2] = in_addrs, P - 1
    t[rows + 10, 0:32:2], t[rows + 10, 1:32:2] = out_addrs, out_mults
    t[rows, 32] = 1
    for i in (1, 2, 3, 4, 6, 7, 8, 9):
        t[rows + i, 33] = 1
        t[rows + i, 35:51] = ext_rc[i - 1 if i < 5 else i - 2]
    t[rows + 5, 34] = 1
    t[rows + 5, 35:51] = list(int_rc) + [ext_rc[4][0], ext_rc[5][0], ext_rc[6][0]]
    return t


def skinny_program_chips(log_sk=6, log_mem=4, seed=47, device=False):
    """(MemoryConst, Poseidon2SkinnyDeg9): permutation k reads constants k .. k + 15 and writes its outputs to fresh
    addresses nobody reads; as many permutations as fit (11 rows each).  `device=True`: the Skinny chip carries its
    permutation inputs (events) and the prover fills the rows on the GPU (zk_tracegen_poseidon2_skinny)."""
    rng = np.random.default_rng(seed)
    n = 1 << log_sk
    k = max(1, (n * 3 // 4) // 11)
    n_const = k + 15
    consts = rng.integers(0, P, n_const, dtype=np.uint64)
    win = np.arange(k)[:, None] + np.arange(16)[None, :]
    reads = np.zeros(n_const, np.uint64)
    np.add.at(reads, win.ravel(), 1)
    main = poseidon2_skinny_rows(consts[win], n)
    prep = poseidon2_skinny_prep_rows(win, (1 << 22) + 16 * np.arange(k)[:, None] + np.arange(16)[None, :], 0, n)
    mem_rows = 1 << log_mem
    assert n_const <= 2 * mem_rows
    mem_prep = np.zeros((2 * mem_rows, 6), np.uint64)
    mem_prep[:n_const, 0], mem_prep[:n_const, 4], mem_prep[:n_const, 5] = consts, np.arange(n_const), reads
    mem = Chip("MemoryConst", "MemoryConst", np.zeros((mem_rows, 1), np.uint32), preprocessed=M(mem_prep.reshape(mem_rows, 12)),
               local_only=True)
    mem.canon = (mem_prep.reshape(mem_rows, 12), np.zeros((mem_rows, 1), np.uint64))
    if device:
        sk = Chip("Poseidon2SkinnyDeg9", "Poseidon2SkinnyDeg9", None, preprocessed=M(prep), log_quotient_degree=3,
                  events=M(consts[win]), tracegen="Poseidon2SkinnyDeg9", rows=n)
    else:
        sk = Chip("Poseidon2SkinnyDeg9", "Poseidon2SkinnyDeg9", M(main), preprocessed=M(prep), log_quotient_degree=3)
    sk.canon = (prep, main)
    return [mem, sk]


# ------------------------------------------------------------------------------------------------------------------
# A toy core-machine PROGRAM: straight-line ALU instructions over registers 1..31, executed here, from which the rows of
# the CPU chip, the Program table, the seven ALU chips that implement the opcodes, MemoryLocal (initial / final register
# states) and the Byte table all follow.  The chips then interlock exactly as in the reference's machine: every
# instruction the CPU sends is received by one ALU chip, every instruction fetch is answered by Program, every register
# access chains on the memory bus from MemoryLocal's initial state to its final state, and every byte lookup is answered
# by Byte -- only the Global-kind lookups of MemoryLocal have no partner (the Global chip is not transcribed).
# ------------------------------------------------------------------------------------------------------------------
_CORE_OPS = {0: "AddSub", 1: "AddSub", 15: "Bitwise", 16: "Bitwise", 17: "Bitwise", 18: "Bitwise", 13: "Lt", 14: "Lt",
             9: "ShiftLeft", 10: "ShiftRight", 11: "ShiftRight", 12: "ShiftRight", 19: "CloClz", 20: "CloClz", 2: "Mul"}


def _core_op(opcode, b, c):
    M32 = 0xFFFFFFFF
    sb, sc = b - (1 << 32) if b >> 31 else b, c - (1 << 32) if c >> 31 else c
    sh = c & 31
    if opcode == 0: return (b + c) & M32
    if opcode == 1: return (b - c) & M32
    if opcode == 15: return b & c
    if opcode == 16: return b | c
    if opcode == 17: return b ^ c
    if opcode == 18: return ~(b | c) & M32
    if opcode == 13: return int(sb < sc)
    if opcode == 14: return int(b < c)
    if opcode == 9: return (b << sh) & M32
    if opcode == 10: return b >> sh
    if opcode == 11: return (sb >> sh) & M32
    if opcode == 12: return ((b >> sh) | (b << (32 - sh))) & M32
    if opcode == 19: return 32 - b.bit_length()
    if opcode == 20: return 32 - ((~b) & M32).bit_length()
    if opcode == 2: return (b * c) & M32
    raise ValueError(opcode)


def core_program_chips(log_cpu=7, seed=51, fill=0.9, shard=1, pc_start=0x1000, device=False):
    """Returns ([Cpu, Program, AddSub, Bitwise, Lt, ShiftLeft, ShiftRight, CloClz, Mul, DivRem, MovCond, Jump, Branch,
    MemoryLocal, Byte], public values (start_pc, next_pc, execution_shard)).  CpuChip::event_to_row (cpu/trace.rs:118-237) for the CPU
    rows.  About 60 % of the instructions are simple ALU operations, 8 % MULT / MULTU / DIV / DIVU / MOD / MODU (the HI
    register is written through the ALU chip's own memory access at clk + 4, and the CPU sends shard and clk along:
    is_check_memory), the rest conditional moves (MEQ / MNE / WSBH), branches
    (all six, with the delay slot: next_next_pc = target when taken) and jumps (Jumpi, JumpDirect); control flow always
    goes FORWARD to fresh addresses, so every executed pc is one row of the Program table.  The chips' own dependencies
    are generated as the executor's generate_dependencies does: CloClz's SRL on ShiftRight, Branch's two SLT on Lt and
    its target ADD on AddSub, JumpDirect's target ADD on AddSub, DivRem's MULT / MULTU on Mul, its ADDs on AddSub and its
    SLTU on Lt, all at UNUSED_PC."""
    rng = np.random.default_rng(seed)
    n = 1 << log_cpu
    real = max(2, int(n * fill))
    regs = {r: int(rng.integers(0, 1 << 32)) for r in list(range(1, 32)) + [33]}   # 33: the HI register
    initial = dict(regs)
    last = {r: (0, 0) for r in regs}                                   # (shard, clk) of the previous access
    touched = set()
    alu_ops = list(_CORE_OPS)
    cpu = np.zeros((n, 67), np.uint64)
    cpu[:, 19], cpu[:, 20], cpu[:, 22] = 1, 1, 1                       # padding rows: imm_b = imm_c = is_rw_a = 1
    cpu_ev = np.zeros((real, 22), np.uint32)                           # zk_cpu_event records (include/zkgpu.h)
    prog = np.zeros((n, 14), np.uint64)
    events = {name: [] for name in set(_CORE_OPS.values())}
    mov_events, jump_events, branch_events, div_events, mult_events = [], [], [], [], []
    M32 = 0xFFFFFFFF

    def word(v):
        return [(v >> (8 * k)) & 0xFF for k in range(4)]

    def access(row, base, reg, clk, value, prev_value=None):
        """MemoryAccessCols::populate_access (memory/consistency/trace.rs:69-105) at columns base.."""
        ps, pc_ = last[reg]
        touched.add(reg)
        compare = ps == shard
        diff = ((clk if compare else shard) - (pc_ if compare else ps) - 1) & M32
        assert diff < (1 << 24)
        if prev_value is not None:
            row[base:base + 4] = word(prev_value)
            base += 4
        row[base:base + 4] = word(value)
        row[base + 4:base + 9] = [ps, pc_, int(compare), diff & 0xFFFF, diff >> 16]
        last[reg] = (shard, clk)

    pc, next_pc, hi_pc = pc_start, pc_start + 4, pc_start + 4
    after_cf = False
    for i in range(real):
        what = int(rng.integers(0, 100))
        kind = ("alu" if after_cf or what < 62 else "muldiv" if what < 70 else "mov" if what < 82 else "branch" if what < 92
                else "jump")
        after_cf = kind in ("branch", "jump")
        ra, rb, rc = (int(x) for x in rng.integers(1, 32, 3))
        clk = 5 * i
        row = cpu[i]
        nnpc = next_pc + 4
        imm_b, is_rw_a, immutable, sequential, hi_slot, check_memory = False, 0, 0, 1, 0, 0
        if kind == "alu":
            opcode = alu_ops[int(rng.integers(0, len(alu_ops)))]
            imm_c = opcode in (19, 20) or int(rng.integers(0, 4)) == 0
            cval = (0 if opcode in (19, 20) else int(rng.integers(0, 1 << 16))) if imm_c else regs[rc]
        elif kind == "muldiv":
            opcode = (3, 4, 5, 6, 7, 8)[int(rng.integers(0, 6))]       # MULT MULTU DIV DIVU MOD MODU
            imm_c = int(rng.integers(0, 4)) == 0
            cval = int(rng.integers(0, 3)) if imm_c else regs[rc]      # small immediates: division by 0, 1, 2
        elif kind == "mov":
            opcode = (50, 51, 52)[int(rng.integers(0, 3))]             # MEQ, MNE, WSBH
            imm_c = opcode == 52 or int(rng.integers(0, 3)) == 0       # an immediate 0 makes c == 0 a common case
            cval = 0 if imm_c else regs[rc]
        elif kind == "branch":
            opcode = (21, 26, 25, 24, 23, 22)[int(rng.integers(0, 6))]  # BEQ BNE BLTZ BLEZ BGTZ BGEZ
            imm_b = opcode not in (21, 26)                             # the compare-with-zero forms: b = 0
            imm_c = True
            target = max(hi_pc, next_pc) + 4 * int(rng.integers(1, 9))
            cval = (target - next_pc) & M32
        else:
            opcode = (28, 29)[int(rng.integers(0, 2))]                 # Jumpi, JumpDirect
            imm_b, imm_c, cval = True, True, 0
            target = max(hi_pc, next_pc) + 4 * int(rng.integers(1, 9))
        c_prev = last[rc] if not imm_c else (0, 0)
        if not imm_c:
            access(row, 56, rc, clk + 1, cval)
        else:
            row[56:60] = word(cval)
        if kind == "jump":
            bval = target if opcode == 28 else (target - next_pc) & M32
        elif imm_b:
            bval = 0
        else:
            if kind == "branch" and int(rng.integers(0, 3)) == 0:
                rb = ra                                                # equal operands: BEQ taken / BNE not taken
            bval = regs[rb]
        b_prev = last[rb] if not imm_b else (0, 0)
        if not imm_b:
            access(row, 47, rb, clk + 2, bval)
        else:
            row[47:51] = word(bval)
        prev_a = regs[ra]
        if kind == "alu":
            aval = _core_op(opcode, bval, cval)
            events[_CORE_OPS[opcode]].append((pc, opcode, aval, bval, cval, next_pc))
            if opcode in (19, 20):
                # CloClz's dependency (alu/clo_clz eval: send_alu(SRL, sr1, bb, 31 - a) unless bb = 0) at UNUSED_PC
                bb = bval if opcode == 19 else M32 - bval
                if bb:
                    events["ShiftRight"].append((1, 10, bb >> (31 - aval), bb, 31 - aval, 5))
        elif kind == "muldiv":
            signed = opcode in (3, 5, 7)
            sx = lambda v: (v ^ 0x80000000) - 0x80000000 if signed else v
            if opcode in (3, 4):                                       # MULT / MULTU: low word to ra, high word to HI
                p64 = (sx(bval) * sx(cval)) & 0xFFFFFFFFFFFFFFFF
                aval, hi_val = p64 & M32, p64 >> 32
            else:
                q, rem = _quotient_remainder(bval, cval, signed)
                aval, hi_val = (q, rem) if opcode in (5, 6) else (rem, None)
                # DivRemChip's dependencies (alu/divrem eval): c * quotient on Mul, the absolute values on AddSub,
                # |remainder| < max(|c|, 1) on Lt -- all at UNUSED_PC
                ctq = (sx(q) * sx(cval)) & 0xFFFFFFFFFFFFFFFF
                mult_events.append((1, 5, 3 if signed else 4, ctq >> 32, ctq & M32, q, cval, 0, 0, 0, 0, 0, 0))
                abs_c, abs_rem = (abs(sx(cval)) & M32, abs(sx(rem)) & M32)
                if signed and cval >> 31:
                    events["AddSub"].append((1, 0, 0, cval, abs_c, 5))
                if signed and rem >> 31:
                    events["AddSub"].append((1, 0, 0, rem, abs_rem, 5))
                if cval:
                    events["Lt"].append((1, 14, 1, abs_rem, max(1, abs_c), 5))
            if hi_val is not None:                                     # the HI register write, checked by the ALU chip at clk + 4
                check_memory, hi_slot = 1, hi_val
                ps, pc_ = last[33]
                touched.add(33)
                rec = (shard, clk, regs[33], ps, pc_)
                last[33], regs[33] = (shard, clk + 4), hi_val
            if opcode in (3, 4):
                mult_events.append((pc, next_pc, opcode, hi_val, aval, bval, cval, 1) + rec)
            else:
                div_events.append((pc, next_pc, opcode, bval, cval) + (rec if hi_val is not None else (0, 0, 0, 0, 0)))
        elif kind == "mov":
            if opcode == 52:
                aval = ((bval & 0x00FF00FF) << 8) | ((bval & 0xFF00FF00) >> 8)
            else:
                take = (cval == 0) == (opcode == 50)
                aval = bval if take else prev_a
                is_rw_a, hi_slot = 1, prev_a
            mov_events.append((pc, next_pc, opcode, aval, bval, cval, hi_slot))
        elif kind == "branch":
            aval, immutable, sequential = prev_a, 1, 0
            sa, sb = (aval ^ 0x80000000) - 0x80000000, (bval ^ 0x80000000) - 0x80000000
            lt, gt, eq = sa < sb, sa > sb, sa == sb
            taken = {21: eq, 26: not eq, 25: lt, 24: lt or eq, 23: gt, 22: eq or gt}[opcode]
            if taken:
                nnpc = target
                events["AddSub"].append((1, 0, target, next_pc, cval, 5))  # send_alu(ADD, target_pc, next_pc, c)
            events["Lt"].append((1, 13, int(lt), aval, bval, 5))        # send_alu(SLT, a_lt_b, a, b)
            events["Lt"].append((1, 13, int(gt), bval, aval, 5))        # send_alu(SLT, a_gt_b, b, a)
            branch_events.append((pc, next_pc, nnpc, opcode, aval, bval, cval))
        else:
            aval, sequential, nnpc = (next_pc + 4) & M32, 0, target    # the link address goes to register ra
            if opcode == 29:
                events["AddSub"].append((1, 0, target, next_pc, bval, 5))  # send_alu(ADD, next_next_pc, next_pc, op_b)
            jump_events.append((pc, next_pc, nnpc, opcode, aval, bval, cval))
        a_prev = last[ra]
        access(row, 34, ra, clk + 3, aval, prev_value=prev_a)
        regs[ra] = aval
        op_b_word = bval if imm_b else rb
        op_c_word = cval if imm_c else rc
        row[0], row[1], row[2] = shard, clk & 0xFFFF, clk >> 16
        row[3], row[4], row[23] = shard * check_memory, clk * check_memory, check_memory
        row[5], row[6], row[7] = pc, next_pc, nnpc
        row[8], row[9] = opcode, ra
        row[10:14], row[14:18] = word(op_b_word), word(op_c_word)
        row[18], row[19], row[20] = 0, int(imm_b), int(imm_c)
        row[22], row[25] = is_rw_a, sequential
        row[26:30], row[30:34] = word(aval), word(hi_slot)
        row[65], row[66] = 1, immutable
        prog[i] = [pc, opcode, ra] + word(op_b_word) + word(op_c_word) + [0, int(imm_b), int(imm_c)]
        flags = int(imm_b) << 1 | int(imm_c) << 2 | is_rw_a << 3 | check_memory << 4 | sequential << 6 | immutable << 7
        cpu_ev[i] = [pc, next_pc, nnpc, clk, shard, opcode, ra, op_b_word, op_c_word, flags, 0, aval, bval, cval, hi_slot,
                     prev_a, a_prev[0], a_prev[1], b_prev[0], b_prev[1], c_prev[0], c_prev[1]]
        hi_pc = max(hi_pc, pc, next_pc, nnpc)
        pc, next_pc = next_pc, nnpc
    final_next_pc = pc                                                 # the last executed row's next_pc

    def pow2(k):
        return max(2, (max(k, 1) - 1).bit_length())
    chips = []
    c = Chip("Cpu", "Cpu", None, events=cpu_ev, tracegen="Cpu", rows=n) if device else Chip("Cpu", "Cpu", M(cpu))
    c.canon, c.cpu_events = (None, cpu), cpu_ev                        # device=True: the rows are filled on the GPU (zk_tracegen_cpu)
    chips.append(c)
    mult = np.zeros((n, 1), np.uint64)
    mult[:real] = 1
    c = Chip("Program", "Program", M(mult), preprocessed=M(prog))
    c.canon = (prog, mult)
    chips.append(c)
    fillers = {"AddSub": add_sub_rows, "Bitwise": bitwise_rows, "Lt": lt_rows, "ShiftLeft": shift_left_rows,
               "ShiftRight": shift_right_rows, "CloClz": clo_clz_rows}
    for name in ("AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Mul"):
        e = np.array(events[name], np.uint64).reshape(-1, 6)          # pc, opcode, a, b, c, next_pc
        h = 1 << pow2(len(e))
        if name == "Mul":                                              # MUL: no HI write, shard = clk = 0
            ev = np.zeros((len(e), 13), np.uint64)
            ev[:, 0], ev[:, 1], ev[:, 2], ev[:, 4], ev[:, 5], ev[:, 6] = e[:, 0], e[:, 5], e[:, 1], e[:, 2], e[:, 3], e[:, 4]
            # MULT / MULTU instructions (HI written) and DivRem's c * quotient products (at UNUSED_PC, HI not written)
            ev = np.concatenate([ev, np.array(mult_events, np.uint64).reshape(-1, 13)])
            h = 1 << pow2(len(ev))
            t = mul_rows(ev, h)
        else:
            ev = _alu_event_array(e[:, 0], e[:, 1], e[:, 2], e[:, 3], e[:, 4])
            ev[:, 1] = e[:, 5]                                         # the delay slot of a taken branch: next_pc = target
            t = fillers[name](ev, h)
        c = Chip(name, name, M(t), local_only=_air(name).local_only)
        c.canon = (None, t)
        chips.append(c)
    e = np.array(div_events, np.uint64).reshape(-1, 10)
    t = div_rem_rows(e, 1 << pow2(len(e)))
    c = Chip("DivRem", "DivRem", M(t), local_only=True)
    c.canon = (None, t)
    chips.append(c)
    for name, evs, rows_of in (("MovCond", mov_events, mov_cond_rows), ("Jump", jump_events, jump_rows),
                               ("Branch", branch_events, branch_rows)):
        e = np.array(evs, np.uint64).reshape(-1, 7)
        t = rows_of(e, 1 << pow2(len(e)))
        c = Chip(name, name, M(t), local_only=True)
        c.canon = (None, t)
        chips.append(c)
    regs_used = sorted(touched)
    rows_m = 1 << pow2(-(-len(regs_used) // 4))
    ml = np.zeros((4 * rows_m, 14), np.uint64)
    for k, r in enumerate(regs_used):
        ml[k] = [r, 0, last[r][0], 0, last[r][1]] + word(initial[r]) + word(regs[r]) + [1]
    ml = ml.reshape(rows_m, 56)
    c = Chip("MemoryLocal", "MemoryLocal", M(ml))
    c.canon = (None, ml)
    chips.append(c)
    chips.append(byte_chip_for([ch for ch in chips if ch.air != "Program"]))
    pvs = {40: pc_start, 41: final_next_pc, 44: shard}
    chips[0].core_pvs = pvs
    return chips, pvs
