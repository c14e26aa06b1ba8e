"""AIRs compiled into libzkgpu: the synthetic ones BASELINE.md section 4 names (fibonacci, wide_bitwise_*, lookup_*, ...)
and 28 REAL Ziren chips transcribed by hand from their `Air::eval`, one function per chip citing the lines it follows --
seventeen of the core machine (Cpu, AddSub, Mul, DivRem, Lt, Bitwise, ShiftLeft, ShiftRight, CloClz, MovCond, Jump,
Branch, MemoryLocal, SyscallCore, SyscallPrecompile, Byte, Program) and all eleven RecursionAir variants.  The other
MipsAir chips need the Rust-side exporter (rust/air-export), whose JSON `Air.from_exported_json` loads."""
from .ir import Air, AirBuilder


def fibonacci():
    """FibonacciAir of crates/stark/src/stark_testing.rs:25-62: 2 columns, 3 public values, 5 constraints."""
    air = Air("fibonacci", main_width=2, num_public_values=3)
    b = AirBuilder(air)
    pis = b.public_values()
    local, nxt = b.main().local(), b.main().next()
    first = b.when_first_row()
    first.assert_eq(local[0], pis[0])
    first.assert_eq(local[1], pis[1])
    tr = b.when_transition()
    tr.assert_eq(local[1], nxt[0])
    tr.assert_eq(local[0] + local[1], nxt[1])
    b.when_last_row().assert_eq(local[1], pis[2])
    b.eval_permutation_constraints()
    return air


def wide_bitwise(width=256, name=None):
    """Keccak-like wide synthetic chip: groups of 4 columns (a, b, x, c) with
        a, b, c boolean;  x = a xor b = a + b - 2ab;  next.a = c on transitions (degree-3 with the selector),
        c * (x - a) * b = 0 ... is replaced by the cubic  c*x = c*(a + b - 2ab)  to load the multiplier.
    width/4 * 6 constraints, max degree 3."""
    assert width % 4 == 0
    air = Air(name or f"wide_bitwise_{width}", main_width=width)
    b = AirBuilder(air)
    local, nxt = b.main().local(), b.main().next()
    for g in range(width // 4):
        a, bb, x, c = local[4 * g: 4 * g + 4]
        b.assert_bool(a)
        b.assert_bool(bb)
        b.assert_bool(c)
        xor = a + bb - 2 * (a * bb)
        b.assert_eq(x, xor)
        b.assert_eq(c * x, c * xor)
        b.when_transition().assert_eq(nxt[4 * g], c)
    b.eval_permutation_constraints()
    return air


def lookup_pair():
    """Small chip with preprocessed columns and LogUp lookups, in the shape of Ziren's chips: its own
    constraints, then three lookups (two sends, one receive; batch size 2 -> two batch columns + the running
    sum) whose permutation constraints are appended by eval_permutation_constraints.
    Columns: main (a, b, c, s, m), preprocessed (p0, p1)."""
    air = Air("lookup_pair", main_width=5, prep_width=2, perm_width=0, num_public_values=4)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p = b.preprocessed().local()
    b.assert_eq(m[2], m[0] * m[1] + p[0])
    b.assert_bool(m[3])
    # public value 3: the shard's public values are ONE vector shared by all chips (prover.rs:322, quotient.rs:30);
    # 0..2 belong to the Fibonacci chip
    b.when_transition().assert_eq(mn[0], m[0] + b.public_values()[3])
    b.send(4, [m[0], m[1] + 2 * p[1], 7], m[3])            # Byte-kind lookup with a linear combination and a constant
    b.send(1, [m[2] - p[0]], m[4])                         # Memory-kind
    b.receive(4, [p[0], p[1], m[0] + 1], m[3] + m[4])      # Byte-kind receive
    b.eval_permutation_constraints(batch_size=2)
    return air


def quintic():
    """Degree-5 chip: log_quotient_degree = 2 (four quotient chunks; needs log_blowup >= 2, the shrink configuration
    crates/stark/src/kb31_poseidon2.rs:217-227).  Columns (a, b, d): d = a^4 * b, next.a = a + 1 on transitions,
    b boolean on the first row."""
    air = Air("quintic", main_width=3)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    a2 = m[0] * m[0]
    b.assert_eq(m[2], a2 * a2 * m[1])
    b.when_transition().assert_eq(mn[0], m[0] + 1)
    b.when_first_row().assert_bool(m[1])
    b.eval_permutation_constraints()
    return air


def lookup_side(send, name=None):
    """One side of a balanced LogUp pair: `lookup_send` sends (kind 5, [x, y + 3]) with multiplicity m, `lookup_recv`
    receives the same tuples with the same multiplicities, so over a shard holding both chips with the same (x, y, m)
    columns the local cumulative sums cancel -- which is what the verifier demands of a shard
    (crates/stark/src/verifier.rs:236-244).  Columns (x, y, m, pad); x * (x - y) * pad = 0 keeps a cubic constraint."""
    air = Air(name or ("lookup_send" if send else "lookup_recv"), main_width=4, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    b.assert_zero(m[0] * (m[0] - m[1]) * m[3])
    (b.send if send else b.receive)(5, [m[0], m[1] + 3], m[2])
    b.eval_permutation_constraints(batch_size=2)
    return air


def global_tail():
    """Global-scope chip (the shape of Ziren's `Global` chip as far as this path sees it): its last 14 main columns
    carry the running septic digest, and on the last row they ARE the chip's global cumulative sum
    (prover.rs:353-361, permutation.rs:333-346).  Columns (b, s, d0..d13): b boolean, next.s = s + b."""
    air = Air("global_tail", main_width=16, commit_scope="global")
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    b.assert_bool(m[0])
    b.when_transition().assert_eq(mn[1], m[1] + m[0])
    b.eval_permutation_constraints()
    return air


def local_bool(width=8):
    """`local_only` chip: no constraint touches the next row, so it is opened at zeta only (prover.rs:503-531)."""
    air = Air("local_bool", main_width=width, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    for g in range(width // 4):
        a, bb, x, c = m[4 * g: 4 * g + 4]
        b.assert_bool(a)
        b.assert_bool(bb)
        b.assert_eq(x, a + bb - 2 * (a * bb))
        b.assert_eq(c * x, c * a)
    b.eval_permutation_constraints()
    return air


# ------------------------------------------------------------------------------------------------------------------
# Real Ziren chips, transcribed by hand from their `Air::eval` (the Rust-side exporter of SURVEY f2 would emit exactly
# these programs as JSON: `Air.to_json`).  Column order = the `#[repr(C)]` column structs; constraint and lookup
# emission order = the order of the builder calls in `eval`, which fixes the powers of alpha and the LogUp batches.
# Committed columns per row (main + 4 * permutation + 4 * 2 quotient) match `mips_costs.json`: AddSub 47, Lt 56,
# Bitwise 42.
# ------------------------------------------------------------------------------------------------------------------
LOOKUP_MEMORY, LOOKUP_PROGRAM, LOOKUP_INSTRUCTION, LOOKUP_BYTE = 1, 2, 3, 4   # LookupKind, stark/src/lookup/lookup.rs:23-44
BYTE_AND, BYTE_OR, BYTE_XOR, BYTE_U8RANGE, BYTE_LTU, BYTE_NOR = 0, 1, 2, 4, 6, 9  # ByteOpcode, executor/src/opcode.rs:184-205
OP_ADD, OP_SUB, OP_SLT, OP_SLTU, OP_AND, OP_OR, OP_XOR, OP_NOR = 0, 1, 13, 14, 15, 16, 17, 18  # Opcode, opcode.rs:15-35


def _send_byte(b, opcode, a, bb, c, mult):
    """ZKMAirBuilder::send_byte -> send_byte_pair(opcode, a1, 0, b, c) (stark/src/air/builder.rs:120-150)"""
    b.send(LOOKUP_BYTE, [opcode, a, 0, bb, c], mult)


def _slice_range_check_u8(b, xs, mult):
    """WordAirBuilder::slice_range_check_u8 (core/machine/src/air/word.rs:55-80): bytes are checked in pairs"""
    i = 0
    while i + 1 < len(xs):
        _send_byte(b, BYTE_U8RANGE, 0, xs[i], xs[i + 1], mult)
        i += 2
    if i < len(xs):
        _send_byte(b, BYTE_U8RANGE, 0, xs[i], 0, mult)


def _receive_instruction(b, pc, next_pc, opcode, a, bb, c, mult):
    """ZKMAirBuilder::receive_instruction as the ALU chips call it (stark/src/air/builder.rs:237-279): shard = clk = 0,
    next_next_pc = next_pc + 4, num_extra_cycles = 0, hi = 0, op_a_immutable = is_rw_a = is_check_memory = is_halt = 0,
    is_sequential = 1."""
    vals = [0, 0, pc, next_pc, next_pc + 4, 0, opcode] + list(a) + list(bb) + list(c) + [0, 0, 0, 0] + [0, 0, 0, 0, 1]
    b.receive(LOOKUP_INSTRUCTION, vals, mult)


def add_sub():
    """AddSubChip (crates/core/machine/src/alu/add_sub/mod.rs:42-62 columns, :180-249 eval) with AddOperation
    (crates/core/machine/src/operations/add.rs:14-99).  19 main columns, 14 constraints + 6 byte sends + 2 instruction
    receives; `local_only`."""
    air = Air("AddSub", main_width=19, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    value, carry = m[2:6], m[6:9]
    op1, op2 = m[9:13], m[13:17]
    is_add, is_sub = m[17], m[18]
    # AddOperation::eval(builder, operand_1, operand_2, add_operation, is_add + is_sub)
    is_real = is_add + is_sub
    r = b.when(is_real)
    base = 256
    o0 = op1[0] + op2[0] - value[0]
    o1 = op1[1] + op2[1] - value[1] + carry[0]
    o2 = op1[2] + op2[2] - value[2] + carry[1]
    o3 = op1[3] + op2[3] - value[3] + carry[2]
    r.assert_zero(o3 * (o3 - base))
    r.assert_zero(carry[0] * (o0 - base))
    r.assert_zero(carry[1] * (o1 - base))
    r.assert_zero(carry[2] * (o2 - base))
    r.assert_zero((carry[0] - 1) * o0)
    r.assert_zero((carry[1] - 1) * o1)
    r.assert_zero((carry[2] - 1) * o2)
    r.assert_bool(carry[0])
    r.assert_bool(carry[1])
    r.assert_bool(carry[2])
    r.assert_bool(is_real)
    _slice_range_check_u8(b, op1, is_real)
    _slice_range_check_u8(b, op2, is_real)
    _slice_range_check_u8(b, value, is_real)
    _receive_instruction(b, pc, next_pc, OP_ADD, value, op1, op2, is_add)
    _receive_instruction(b, pc, next_pc, OP_SUB, op1, value, op2, is_sub)
    b.assert_bool(is_add)
    b.assert_bool(is_sub)
    b.assert_bool(is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


def bitwise():
    """BitwiseChip (crates/core/machine/src/alu/bitwise/mod.rs:33-58 columns, :179-234 eval): 18 main columns, 4 byte
    sends + 1 instruction receive, 5 constraints (the reference asserts `is_xor` boolean twice and never `is_nor`;
    transcribed as written)."""
    air = Air("Bitwise", main_width=18, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    a, bb, c = m[2:6], m[6:10], m[10:14]
    is_nor, is_xor, is_or, is_and = m[14], m[15], m[16], m[17]
    opcode = is_xor * BYTE_XOR + is_or * BYTE_OR + is_and * BYTE_AND + is_nor * BYTE_NOR
    mult = is_xor + is_or + is_and + is_nor
    for i in range(4):
        _send_byte(b, opcode, a[i], bb[i], c[i], mult)
    cpu_opcode = is_xor * OP_XOR + is_or * OP_OR + is_and * OP_AND + is_nor * OP_NOR
    _receive_instruction(b, pc, next_pc, cpu_opcode, a, bb, c, mult)
    b.assert_bool(is_xor)
    b.assert_bool(is_or)
    b.assert_bool(is_and)
    b.assert_bool(is_xor)
    b.assert_bool(mult)
    b.eval_permutation_constraints(batch_size=2)
    return air


def lt():
    """LtChip (crates/core/machine/src/alu/lt/mod.rs:36-85 columns, :274-468 eval): 36 main columns, 3 byte sends + 1
    instruction receive."""
    air = Air("Lt", main_width=36, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc, is_slt, is_sltu = m[0], m[1], m[2], m[3]
    a, bw, cw = m[4:8], m[8:12], m[12:16]
    byte_flags = m[16:20]
    b_masked, c_masked, not_eq_inv = m[20], m[21], m[22]
    msb_b, msb_c, bit_b, bit_c = m[23], m[24], m[25], m[26]
    sltu, is_comp_eq, is_sign_eq = m[27], m[28], m[29]
    comparison_bytes = m[30:32]
    # m[32:36] = byte_equality_check: allocated by the reference, never constrained
    is_real = is_slt + is_sltu
    b_comp = list(bw)
    c_comp = list(cw)
    b_comp[3] = bw[3] * is_sltu + b_masked * is_slt
    c_comp[3] = cw[3] * is_sltu + c_masked * is_slt
    _send_byte(b, BYTE_AND, b_masked, bw[3], 0x7F, is_real)
    _send_byte(b, BYTE_AND, c_masked, cw[3], 0x7F, is_real)
    b.assert_eq(bit_b, msb_b * is_slt)
    b.assert_eq(bit_c, msb_c * is_slt)
    inv_128 = pow(128, -1, 0x7F000001)
    b.assert_eq(msb_b, (bw[3] - b_masked) * inv_128)
    b.assert_eq(msb_c, (cw[3] - c_masked) * inv_128)
    b.assert_bool(is_sign_eq)
    b.when(is_sign_eq).assert_eq(bit_b, bit_c)
    b.when(is_real).when_not(is_sign_eq).assert_one(bit_b + bit_c)
    b.assert_eq(a[0], bit_b * (1 - bit_c) + is_sign_eq * sltu)
    b.assert_zero(a[1])
    b.assert_zero(a[2])
    b.assert_zero(a[3])
    sum_flags = byte_flags[0] + byte_flags[1] + byte_flags[2] + byte_flags[3]
    for f in byte_flags:
        b.assert_bool(f)
    b.assert_bool(sum_flags)
    b.when(is_real).assert_eq(1 - is_comp_eq, sum_flags)
    b.assert_bool(is_comp_eq)
    visited = None
    b_byte_sel = c_byte_sel = None
    for k in (3, 2, 1, 0):
        flag = byte_flags[k]
        visited = flag if visited is None else visited + flag
        b_byte_sel = b_comp[k] * flag if b_byte_sel is None else b_byte_sel + b_comp[k] * flag
        c_byte_sel = c_comp[k] * flag if c_byte_sel is None else c_byte_sel + c_comp[k] * flag
        b.when_not(visited).assert_eq(b_comp[k], c_comp[k])
        b.when(is_comp_eq).assert_zero(visited)
    b.assert_eq(comparison_bytes[0], b_byte_sel)
    b.assert_eq(comparison_bytes[1], c_byte_sel)
    b.when_not(is_comp_eq).assert_eq(not_eq_inv * (comparison_bytes[0] - comparison_bytes[1]), is_real)
    _send_byte(b, BYTE_LTU, sltu, comparison_bytes[0], comparison_bytes[1], is_real)
    b.assert_bool(is_slt)
    b.assert_bool(is_sltu)
    b.assert_bool(is_slt + is_sltu)
    _receive_instruction(b, pc, next_pc, is_slt * OP_SLT + is_sltu * OP_SLTU, a, bw, cw, is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


# ------------------------------------------------------------------------------------------------------------------
# Recursion chip: Poseidon2WideChip<DEGREE> (crates/recursion/core/src/chips/poseidon2_wide/air.rs:34-178), the widest
# chip of the compress / shrink provers (RecursionAir<F, 3> / <F, 9>, crates/recursion/core/src/machine.rs).  Columns:
# chips/poseidon2_wide/columns/permutation.rs:20-35 (main), columns/preprocessed.rs:8-14 (preprocessed).
# ------------------------------------------------------------------------------------------------------------------
def _poseidon2_round_constants():
    """canonical RC_16_30_U32 rows as the chip indexes them (crates/primitives/src/lib.rs:563-1104): 8 external rounds
    (round r < 4 -> row r, else row r + 13) and column 0 of the 13 internal rows, read from the generated header the
    CUDA kernels and the oracle use"""
    import os
    import re
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "include",
                             "zk_poseidon2_rc.h")).read().replace("\\\n", " ")
    def table(name):
        body = re.search(r"#define " + name + r"\s+(.*)", text).group(1)
        return [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)u", body)]
    ext = table("ZK_P2_EXT_RC_CANON")
    return [ext[16 * r:16 * r + 16] for r in range(8)], table("ZK_P2_INT_RC_CANON")


P = 0x7F000001
P2_DIAG = [P - 2, 1, 2, (P + 1) >> 1, 3, 4, (P - 1) >> 1, P - 3, P - 4, P - ((P - 1) >> 8), P - ((P - 1) >> 3), P - 127,
           (P - 1) >> 8, (P - 1) >> 3, (P - 1) >> 4, 127]  # chips/poseidon2_wide/mod.rs:85-102


def _apply_m_4(x):
    """chips/poseidon2_wide/mod.rs:46-60"""
    t01 = x[0] + x[1]
    t23 = x[2] + x[3]
    t0123 = t01 + t23
    t01123 = t0123 + x[1]
    t01233 = t0123 + x[3]
    return [t01123 + t01, t01123 + (x[2] + x[2]), t01233 + t23, t01233 + (x[0] + x[0])]


def _external_linear_layer(state):
    """chips/poseidon2_wide/mod.rs:63-74"""
    st = []
    for j in range(0, 16, 4):
        st += _apply_m_4(state[j:j + 4])
    sums = [st[k] + st[4 + k] + st[8 + k] + st[12 + k] for k in range(4)]
    return [st[j] + sums[j % 4] for j in range(16)]


def _internal_linear_layer(state):
    """chips/poseidon2_wide/mod.rs:104-114 -> p3_poseidon2::matmul_internal: state[i] * diag[i] + sum"""
    total = state[0]
    for x in state[1:]:
        total = total + x
    return [state[i] * P2_DIAG[i] + total for i in range(16)]


def poseidon2_wide(degree=3):
    """Poseidon2WideChip<DEGREE>::eval, statement by statement.  DEGREE 3 carries the S-box columns (313 main columns),
    DEGREE 9 does not (172); 49 preprocessed columns; 32 memory sends; 1 + 8 * 16 (+ 8 * 16) + 12 + 16 (+ 13)
    constraints before the permutation argument; `local_only`."""
    assert degree in (3, 9)
    sbox_cols = degree == 3
    ext_rc, int_rc = _poseidon2_round_constants()
    air = Air(f"Poseidon2WideDeg{degree}", main_width=313 if sbox_cols else 172, prep_width=49, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    p = b.preprocessed().local()
    ext_state = [m[16 * r:16 * r + 16] for r in range(8)]
    int_state, s0, out = m[128:144], m[144:156], m[156:172]
    ext_sbox = [m[172 + 16 * r:188 + 16 * r] for r in range(8)] if sbox_cols else None
    int_sbox = m[300:313] if sbox_cols else None
    prep_input = p[0:16]
    prep_output = [(p[16 + 2 * i], p[17 + 2 * i]) for i in range(16)]   # MemoryAccessColsChips {addr, mult}
    is_real_neg = p[48]
    # air.rs:45-52: dummy constraint that normalises the chip to DEGREE
    x00 = ext_state[0][0]
    lhs = x00
    for _ in range(degree - 1):
        lhs = lhs * x00
    b.assert_eq(lhs, lhs)
    # air.rs:55-69: memory lookups, send_single(addr, val, mult) = send(Memory, [addr, val, 0, 0, 0], mult)
    # (crates/recursion/core/src/builder.rs:19-44)
    for i in range(16):
        b.send(LOOKUP_MEMORY, [prep_input[i], ext_state[0][i], 0, 0, 0], is_real_neg)
    for i in range(16):
        b.send(LOOKUP_MEMORY, [prep_output[i][0], out[i], 0, 0, 0], prep_output[i][1])
    # air.rs:72-74, eval_external_round :83-139
    for r in range(8):
        state = list(ext_state[r])
        if r == 0:
            state = _external_linear_layer(state)
        add_rc = [state[i] + ext_rc[r][i] for i in range(16)]
        sb = []
        for i in range(16):
            cube = add_rc[i] * add_rc[i] * add_rc[i]
            if sbox_cols:
                b.assert_eq(ext_sbox[r][i], cube)
                sb.append(ext_sbox[r][i])
            else:
                sb.append(cube)
        state = _external_linear_layer(sb)
        nxt = int_state if r == 3 else out if r == 7 else ext_state[r + 1]
        for i in range(16):
            b.assert_eq(nxt[i], state[i])
    # eval_internal_rounds, air.rs:142-177
    state = list(int_state)
    for r in range(13):
        add_rc = (state[0] if r == 0 else s0[r - 1]) + int_rc[r]
        cube = add_rc * add_rc * add_rc
        if sbox_cols:
            b.assert_eq(int_sbox[r], cube)
            cube = int_sbox[r]
        state[0] = cube
        state = _internal_linear_layer(state)
        if r < 12:
            b.assert_eq(s0[r], state[0])
    for i in range(16):
        b.assert_eq(ext_state[4][i], state[i])
    b.eval_permutation_constraints(batch_size=2 if degree == 3 else 8)
    return air


def memory_const():
    """MemoryConstChip (crates/recursion/core/src/chips/mem/constant.rs:16-41 columns, :141-154 eval): one unused main
    column, 12 preprocessed columns = 2 x (Block value[4], addr, mult); each entry WRITES its block:
    send_block(addr, value, mult) = send(Memory, [addr, v0, v1, v2, v3], mult) (builder.rs:30-44)."""
    air = Air("MemoryConst", main_width=1, prep_width=12, local_only=True)
    b = AirBuilder(air)
    p = b.preprocessed().local()
    for e in range(2):
        value, addr, mult = p[6 * e:6 * e + 4], p[6 * e + 4], p[6 * e + 5]
        b.send(LOOKUP_MEMORY, [addr] + list(value), mult)
    b.eval_permutation_constraints(batch_size=2)
    return air


def base_alu():
    """BaseAluChip (crates/recursion/core/src/chips/alu_base.rs:27-62 columns, :266-297 eval): 4 operations per row.
    Main: 4 x BaseAluIo {out, in1, in2}; preprocessed: 4 x {addrs {out, in1, in2}, is_add, is_sub, is_mul, is_div, mult}.
    Per operation 5 constraints, two memory reads (receive_single, multiplicity is_real) and one write (send_single,
    multiplicity mult).  Lookups are recorded in call order: receive in1, receive in2, send out."""
    air = Air("BaseAlu", main_width=12, prep_width=32, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    p = b.preprocessed().local()
    for e in range(4):
        out, in1, in2 = m[3 * e], m[3 * e + 1], m[3 * e + 2]
        a_out, a_in1, a_in2 = p[8 * e], p[8 * e + 1], p[8 * e + 2]
        is_add, is_sub, is_mul, is_div, mult = p[8 * e + 3:8 * e + 8]
        is_real = is_add + is_sub + is_mul + is_div
        b.assert_bool(is_real)
        b.when(is_add).assert_eq(in1 + in2, out)
        b.when(is_sub).assert_eq(in1, in2 + out)
        b.when(is_mul).assert_eq(out, in1 * in2)
        b.when(is_div).assert_eq(in2 * out, in1)
        b.receive(LOOKUP_MEMORY, [a_in1, in1, 0, 0, 0], is_real)
        b.receive(LOOKUP_MEMORY, [a_in2, in2, 0, 0, 0], is_real)
        b.send(LOOKUP_MEMORY, [a_out, out, 0, 0, 0], mult)
    b.eval_permutation_constraints(batch_size=2)
    return air


def memory_var():
    """MemoryVarChip (crates/recursion/core/src/chips/mem/variable.rs:15-38 columns, :148-159 eval): hint / witness
    values written to memory: 2 entries per row, main = 2 x Block value[4], preprocessed = 2 x {addr, mult};
    send_block(addr, value, mult) per entry."""
    air = Air("MemoryVar", main_width=8, prep_width=4, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    p = b.preprocessed().local()
    for e in range(2):
        b.send(LOOKUP_MEMORY, [p[2 * e]] + list(m[4 * e:4 * e + 4]), p[2 * e + 1])
    b.eval_permutation_constraints(batch_size=2)
    return air


def _ext_mul(a, c):
    """BinomialExtension<Expr> * BinomialExtension<Expr>, X^4 = 3 (crates/stark/src/air/extension.rs:55-75)"""
    res = [None] * 4
    for i in range(4):
        for j in range(4):
            k = i + j
            term = 3 * a[i] * c[j] if k >= 4 else a[i] * c[j]
            k %= 4
            res[k] = term if res[k] is None else res[k] + term
    return res


def ext_alu():
    """ExtAluChip (crates/recursion/core/src/chips/alu_ext.rs:18-57 columns, :268-302 eval): 4 extension-field operations
    per row.  Main: 4 x ExtAluIo<Block> {out[4], in1[4], in2[4]}; preprocessed: 4 x {addrs {out, in1, in2}, is_add,
    is_sub, is_mul, is_div, mult}.  assert_ext_eq = one assert_eq per coefficient (stark/src/air/builder.rs:400-408)."""
    air = Air("ExtAlu", main_width=48, prep_width=32, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    p = b.preprocessed().local()
    for e in range(4):
        out, in1, in2 = m[12 * e:12 * e + 4], m[12 * e + 4:12 * e + 8], m[12 * e + 8:12 * e + 12]
        a_out, a_in1, a_in2 = p[8 * e], p[8 * e + 1], p[8 * e + 2]
        is_add, is_sub, is_mul, is_div, mult = p[8 * e + 3:8 * e + 8]
        is_real = is_add + is_sub + is_mul + is_div
        b.assert_bool(is_real)
        for flag, lhs, rhs in ((is_add, [x + y for x, y in zip(in1, in2)], out),
                               (is_sub, in1, [y + z for y, z in zip(in2, out)]),
                               (is_mul, _ext_mul(in1, in2), out),
                               (is_div, in1, _ext_mul(in2, out))):
            for l, r in zip(lhs, rhs):
                b.when(flag).assert_eq(l, r)
        b.receive(LOOKUP_MEMORY, [a_in1] + list(in1), is_real)
        b.receive(LOOKUP_MEMORY, [a_in2] + list(in2), is_real)
        b.send(LOOKUP_MEMORY, [a_out] + list(out), mult)
    b.eval_permutation_constraints(batch_size=2)
    return air


def select():
    """SelectChip (crates/recursion/core/src/chips/select.rs:16-35 columns, :229-250 eval): main = SelectIo {bit, out1,
    out2, in1, in2}; preprocessed = {is_real, addrs {bit, out1, out2, in1, in2}, mult1, mult2}."""
    air = Air("Select", main_width=5, prep_width=8, local_only=True)
    b = AirBuilder(air)
    bit, out1, out2, in1, in2 = b.main().local()
    p = b.preprocessed().local()
    is_real, a_bit, a_out1, a_out2, a_in1, a_in2, mult1, mult2 = p
    b.receive(LOOKUP_MEMORY, [a_bit, bit, 0, 0, 0], is_real)
    b.receive(LOOKUP_MEMORY, [a_in1, in1, 0, 0, 0], is_real)
    b.receive(LOOKUP_MEMORY, [a_in2, in2, 0, 0, 0], is_real)
    b.send(LOOKUP_MEMORY, [a_out1, out1, 0, 0, 0], mult1)
    b.send(LOOKUP_MEMORY, [a_out2, out2, 0, 0, 0], mult2)
    b.assert_eq(out1, bit * in2 + (1 - bit) * in1)
    b.assert_eq(out2, bit * in1 + (1 - bit) * in2)
    b.eval_permutation_constraints(batch_size=2)
    return air


def batch_fri(degree=3):
    """BatchFRIChip<DEGREE> (crates/recursion/core/src/chips/batch_fri.rs:36-55 columns, :291-358 eval): one row per
    (alpha_pow, p_at_z, p_at_x) triple of a BatchFRI instruction, acc accumulated down the rows of the instruction and
    written to memory on its `is_end` row.  Main: acc[4], alpha_pow[4], p_at_z[4], p_at_x; preprocessed: is_real, is_end,
    acc_addr, alpha_pow_addr, p_at_z_addr, p_at_x_addr.  Uses the next row (not `local_only`)."""
    air = Air("BatchFRI", main_width=13, prep_width=6)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p = b.preprocessed().local()
    is_real, is_end, acc_addr, alpha_pow_addr, p_at_z_addr, p_at_x_addr = p

    def cols(r):
        return r[0:4], r[4:8], r[8:12], r[12]
    acc, alpha_pow, p_at_z, p_at_x = cols(m)
    n_acc, n_alpha_pow, n_p_at_z, n_p_at_x = cols(mn)
    # :347-350 dummy constraint normalising the chip to DEGREE
    lhs = is_real
    for _ in range(degree - 1):
        lhs = lhs * is_real
    b.assert_eq(lhs, lhs)
    # :300-307 memory reads of alpha_pow, p_at_z (blocks) and p_at_x (single); acc written with multiplicity is_end
    b.receive(LOOKUP_MEMORY, [alpha_pow_addr] + list(alpha_pow), is_real)
    b.receive(LOOKUP_MEMORY, [p_at_z_addr] + list(p_at_z), is_real)
    b.receive(LOOKUP_MEMORY, [p_at_x_addr, p_at_x, 0, 0, 0], is_real)
    b.send(LOOKUP_MEMORY, [acc_addr] + list(acc), is_end)

    def term(ap, z, x):                                 # alpha_pow * (p_at_z - from_base(p_at_x))
        return _ext_mul(ap, [z[0] - x, z[1], z[2], z[3]])
    # :310-315 first row
    for l, r in zip(acc, term(alpha_pow, p_at_z, p_at_x)):
        b.when_first_row().assert_eq(l, r)
    # :318-323 the row after an `is_end` row starts a new accumulator
    for l, r in zip(n_acc, term(n_alpha_pow, n_p_at_z, n_p_at_x)):
        b.when_transition().when(is_end).assert_eq(l, r)
    # :326-332 otherwise it continues the running one
    for l, a, r in zip(n_acc, acc, term(n_alpha_pow, n_p_at_z, n_p_at_x)):
        b.when_transition().when_not(is_end).assert_eq(l, a + r)
    b.eval_permutation_constraints(batch_size=2 if degree == 3 else 8)
    return air


def exp_reverse_bits_len(degree=3):
    """ExpReverseBitsLenChip<DEGREE> (crates/recursion/core/src/chips/exp_reverse_bits.rs:32-67 columns, :355-426 eval):
    one row per exponent bit, accum <- accum^2 * (bit ? x : 1).  Main: x, current_bit, prev_accum_squared,
    prev_accum_squared_times_multiplier, accum, accum_squared, multiplier; preprocessed: x_mem, exponent_mem,
    result_mem (MemoryAccessColsChips {addr, mult}: negative multiplicity = read, chips/mem/mod.rs:14-22),
    iteration_num, is_first, is_last, is_real."""
    air = Air("ExpReverseBitsLen", main_width=7, prep_width=10)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p, pn = b.preprocessed().local(), b.preprocessed().next()
    x, current_bit, prev_accum_squared, pasm, accum, accum_squared, multiplier = m
    x_addr, x_mult, e_addr, e_mult, r_addr, r_mult, _iteration_num, is_first, is_last, is_real = p
    n_is_real = pn[9]
    if degree > 3:                                      # :366-370
        lhs = is_real
        for _ in range(degree - 1):
            lhs = lhs * is_real
        b.assert_eq(lhs, lhs)
    b.send(LOOKUP_MEMORY, [x_addr, x, 0, 0, 0], x_mult)                                             # :374
    b.when_transition().when(n_is_real).when_not(is_last).assert_eq(x, mn[0])                        # :377-381
    b.send(LOOKUP_MEMORY, [e_addr, current_bit, 0, 0, 0], e_mult)                                    # :384-388
    b.when(is_first).assert_eq(accum, multiplier)                                                    # :391
    b.when(is_real).when(current_bit).assert_eq(multiplier, x)                                       # :394-397
    b.when(is_real).when_not(current_bit).assert_eq(multiplier, 1)                                   # :398-401
    b.when(is_real).assert_eq(pasm, prev_accum_squared * multiplier)                                 # :405-408
    b.when(is_real).when_not(is_first).assert_eq(accum, pasm)                                        # :410-413
    b.when(is_real).assert_eq(accum_squared, accum * accum)                                          # :416
    b.when_transition().when(n_is_real).when_not(is_last).assert_eq(mn[2], accum_squared)            # :418-422
    b.send(LOOKUP_MEMORY, [r_addr, accum, 0, 0, 0], r_mult)                                          # :425
    b.eval_permutation_constraints(batch_size=2 if degree == 3 else 8)
    return air


def fri_fold(degree=3):
    """FriFoldChip<DEGREE> (crates/recursion/core/src/chips/fri_fold.rs:45-83 columns, :371-462 eval; a chip of
    machine_wide_with_all_chips, machine.rs:68-87): one row per opened polynomial of a FRI-fold instruction.
    Main (33): z[4], alpha[4], x, p_at_x[4], p_at_z[4], alpha_pow_input[4], ro_input[4], alpha_pow_output[4],
    ro_output[4]; preprocessed (20): is_first, then {addr, mult} of z, alpha, x, alpha_pow_input, ro_input, p_at_x,
    p_at_z, ro_output, alpha_pow_output, then is_real.  Every access is a send with a signed multiplicity."""
    air = Air("FriFold", main_width=33, prep_width=20)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p, pn = b.preprocessed().local(), b.preprocessed().next()
    z, alpha, x = m[0:4], m[4:8], m[8]
    p_at_x, p_at_z, ap_in, ro_in, ap_out, ro_out = m[9:13], m[13:17], m[17:21], m[21:25], m[25:29], m[29:33]
    (z_a, z_m), (al_a, al_m), (x_a, x_m), (api_a, api_m), (roi_a, roi_m), (px_a, px_m), (pz_a, pz_m), (roo_a, roo_m), \
        (apo_a, apo_m) = [(p[1 + 2 * k], p[2 + 2 * k]) for k in range(9)]
    is_real = p[19]
    n_is_first, n_is_real = pn[0], pn[19]
    lhs = is_real                                                      # :484-486 dummy constraint of degree DEGREE
    for _ in range(degree - 1):
        lhs = lhs * is_real
    b.assert_eq(lhs, lhs)
    same = lambda: b.when_transition().when(n_is_real).when_not(n_is_first)
    b.send(LOOKUP_MEMORY, [x_a, x, 0, 0, 0], x_m)                      # :380
    same().assert_eq(x, mn[8])                                         # :383-387
    b.send(LOOKUP_MEMORY, [z_a] + list(z), z_m)                        # :390
    for l, r in zip(z, mn[0:4]):                                       # :393-397
        same().assert_eq(l, r)
    b.send(LOOKUP_MEMORY, [al_a] + list(alpha), al_m)                  # :400
    for l, r in zip(alpha, mn[4:8]):                                   # :403-407
        same().assert_eq(l, r)
    b.send(LOOKUP_MEMORY, [api_a] + list(ap_in), api_m)                # :410-441 vector inputs, then outputs
    b.send(LOOKUP_MEMORY, [roi_a] + list(ro_in), roi_m)
    b.send(LOOKUP_MEMORY, [pz_a] + list(p_at_z), pz_m)
    b.send(LOOKUP_MEMORY, [px_a] + list(p_at_x), px_m)
    b.send(LOOKUP_MEMORY, [apo_a] + list(ap_out), apo_m)
    b.send(LOOKUP_MEMORY, [roo_a] + list(ro_out), roo_m)
    for l, r in zip(_ext_mul(ap_in, alpha), ap_out):                   # :447 new_alpha_pow = old_alpha_pow * alpha
        b.assert_eq(l, r)
    # :458-461 (new_ro - old_ro) * (x - z) = (p_at_x - p_at_z) * old_alpha_pow
    d_ro = [n - o for n, o in zip(ro_out, ro_in)]
    x_minus_z = [x - z[0], 0 - z[1], 0 - z[2], 0 - z[3]]
    d_p = [a - c for a, c in zip(p_at_x, p_at_z)]
    for l, r in zip(_ext_mul(d_ro, x_minus_z), _ext_mul(d_p, ap_in)):
        b.assert_eq(l, r)
    b.eval_permutation_constraints(batch_size=2 if degree == 3 else 8)
    return air


def poseidon2_skinny(degree=9):
    """Poseidon2SkinnyChip<DEGREE> (crates/recursion/core/src/chips/poseidon2_skinny/air.rs:25-163; columns/mod.rs:20-26,
    columns/preprocessed.rs:5-19; the Poseidon2 chip of the wrap machine, machine.rs:138-153): ELEVEN rows per
    permutation -- input row (memory reads), four external rounds, one row holding all 13 internal rounds, four
    external rounds, output row (memory writes) -- tied together by next-row constraints selected by preprocessed round
    flags.  Main (28): state_var[16], internal_rounds_s0[12]; preprocessed (51): 16 x {addr, mult}, is_input_round,
    is_external_round, is_internal_round, round_constants[16].  DEGREE >= 9 only (S-boxes are not materialised)."""
    assert degree >= 9
    air = Air(f"Poseidon2SkinnyDeg{degree}", main_width=28, prep_width=51)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p = b.preprocessed().local()
    state, s0, nxt = list(m[0:16]), m[16:28], mn[0:16]
    mem = [(p[2 * i], p[2 * i + 1]) for i in range(16)]
    is_input, is_external, is_internal, rc = p[32], p[33], p[34], p[35:51]
    lhs = state[0]                                                     # air.rs:43-45
    for _ in range(degree - 1):
        lhs = lhs * state[0]
    b.assert_eq(lhs, lhs)
    for i in range(16):                                                # air.rs:48-54
        b.send(LOOKUP_MEMORY, [mem[i][0], state[i], 0, 0, 0], mem[i][1])
    lin = _external_linear_layer(state)                                # eval_input_round, air.rs:71-91
    for i in range(16):
        b.when_transition().when(is_input).assert_eq(nxt[i], lin[i])
    add_rc = [state[i] + rc[i] for i in range(16)]                     # eval_external_round, air.rs:93-127
    lin = _external_linear_layer([x * x * x for x in add_rc])
    for i in range(16):
        b.when_transition().when(is_external).assert_eq(nxt[i], lin[i])
    st = list(state)                                                   # eval_internal_rounds, air.rs:129-162
    for r in range(13):
        x = (st[0] if r == 0 else s0[r - 1]) + rc[r]
        st[0] = x * x * x
        st = _internal_linear_layer(st)
        if r < 12:
            b.when(is_internal).assert_eq(s0[r], st[0])
    for i in range(16):
        b.when(is_internal).assert_eq(nxt[i], st[i])
    b.eval_permutation_constraints(batch_size=8)
    return air


RECURSIVE_PROOF_NUM_PV_ELTS = 231    # size_of::<RecursionPublicValues<u8>>() = PROOF_MAX_NUM_PVS (stark/src/types.rs:73)
RECURSION_PV_DIGEST = 223            # RECURSION_PUBLIC_VALUES_COL_MAP.digest[0] (recursion/core/src/air/public_values.rs:79-145)


def public_values_chip():
    """PublicValuesChip (crates/recursion/core/src/chips/public_values.rs:37-50 columns, :289-309 eval): 16 rows, row i < 8
    reads digest element i from memory (send_single with multiplicity -1) and ties it to the shard's public values:
    pv_idx[i] * (public_values.digest[i] - pv_element) = 0.  Main: pv_element; preprocessed: pv_idx[8], pv_mem {addr,
    mult}.  RecursionPublicValues is 231 elements, digest at 223..231."""
    air = Air("PublicValues", main_width=1, prep_width=10, num_public_values=RECURSIVE_PROOF_NUM_PV_ELTS)
    b = AirBuilder(air)
    pv_element = b.main().local()[0]
    p = b.preprocessed().local()
    pv = b.public_values()
    b.send(LOOKUP_MEMORY, [p[8], pv_element, 0, 0, 0], p[9])
    for i in range(8):
        b.when(p[i]).assert_eq(pv[RECURSION_PV_DIGEST + i], pv_element)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_MEQ, OP_MNE, OP_WSBH = 50, 51, 52                                    # Opcode, executor/src/opcode.rs:70-72


def _is_zero_operation(b, a, inverse, result, is_real):
    """IsZeroOperation::eval (core/machine/src/operations/is_zero.rs:47-68)"""
    b.when(is_real).assert_eq(1 - inverse * a, result)
    b.when(is_real).assert_bool(result)
    b.when(is_real).when(result).assert_zero(a)


def _is_zero_word_operation(b, word, cols, is_real):
    """IsZeroWordOperation::eval (core/machine/src/operations/is_zero_word.rs:44-78); cols = is_zero_byte[4] x {inverse,
    result}, is_lower_half_zero, is_upper_half_zero, result (11 columns)"""
    byte = [(cols[2 * i], cols[2 * i + 1]) for i in range(4)]
    lower, upper, result = cols[8], cols[9], cols[10]
    for i in range(4):
        _is_zero_operation(b, word[i], byte[i][0], byte[i][1], is_real)
    b.assert_bool(is_real)
    r = b.when(is_real)
    r.assert_bool(lower)
    r.assert_bool(upper)
    r.assert_bool(result)
    r.assert_eq(lower, byte[0][1] * byte[1][1])
    r.assert_eq(upper, byte[2][1] * byte[3][1])
    r.assert_eq(result, lower * upper)


def mov_cond():
    """MovCondChip (crates/core/machine/src/misc/mov_cond/mod.rs:34-55 columns, :152-240 eval): MEQ / MNE (conditional
    move on c == 0 / c != 0; a keeps its previous value otherwise) and WSBH (byte swap inside each half word).  32 main
    columns: pc, next_pc, op_a_value, prev_a_value, op_b_value, op_c_value (words), IsZeroWordOperation c_eq_0 (11),
    is_mne, is_meq, is_wsbh; one instruction receive whose `hi` slot carries prev_a_value and whose is_rw_a is
    is_mne + is_meq; 43 constraints; `local_only`.  mips_costs.json: 32 + 4 * 2 + 8 = 48."""
    air = Air("MovCond", main_width=32, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    a, prev_a, bb, c = m[2:6], m[6:10], m[10:14], m[14:18]
    c_eq_0 = m[18:29]
    is_mne, is_meq, is_wsbh = m[29], m[30], m[31]
    is_real = is_mne + is_meq + is_wsbh
    opcode = is_wsbh * OP_WSBH + is_meq * OP_MEQ + is_mne * OP_MNE
    # receive_instruction(shard 0, clk 0, pc, next_pc, next_pc + 4, 0, opcode, a, b, c, hi = prev_a, op_a_immutable 0,
    #                     is_rw_a = is_mne + is_meq, is_check_memory 0, is_halt 0, is_sequential 1; is_real)
    b.receive(LOOKUP_INSTRUCTION, [0, 0, pc, next_pc, next_pc + 4, 0, opcode] + list(a) + list(bb) + list(c) + list(prev_a)
              + [0, is_mne + is_meq, 0, 0, 1], is_real)
    _is_zero_word_operation(b, c, c_eq_0, is_real)
    c_zero = c_eq_0[10]
    for flag, when_zero, other in ((is_meq, True, bb), (is_meq, False, prev_a), (is_mne, False, bb), (is_mne, True, prev_a)):
        f = b.when(flag).when(c_zero) if when_zero else b.when(flag).when_not(c_zero)
        for l, r in zip(a, other):
            f.assert_eq(l, r)
    for i, j in ((0, 1), (1, 0), (2, 3), (3, 2)):                      # eval_wsbh
        b.when(is_wsbh).assert_eq(a[i], bb[j])
    b.assert_bool(is_mne)
    b.assert_bool(is_meq)
    b.assert_bool(is_wsbh)
    b.assert_bool(is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_JUMP, OP_JUMPI, OP_JUMPDIRECT = 27, 28, 29                          # Opcode, executor/src/opcode.rs:45-47
UNUSED_PC, DEFAULT_PC_INC = 1, 4                                        # stark/src/air/builder.rs:19-22


def _reduce(word):
    """Word::reduce (stark/src/word.rs:60-63)"""
    return word[0] + word[1] * (1 << 8) + word[2] * (1 << 16) + word[3] * (1 << 24)


def _koalabear_word_range_check(b, value, cols, is_real):
    """KoalaBearWordRangeChecker::range_check (core/machine/src/operations/koala_bear_word.rs:45-100): the word is < p =
    0x7F000001.  cols = most_sig_byte_decomp[8], and_most_sig_byte_decomp_0_to_{2..7} (14 columns); 17 constraints."""
    bits, ands = cols[0:8], cols[8:14]
    recomposed = 0
    for i in range(8):
        b.when(is_real).assert_bool(bits[i])
        recomposed = recomposed + bits[i] * (1 << i)
    b.when(is_real).assert_eq(recomposed, value[3])
    b.when(is_real).assert_zero(bits[7])
    b.when(is_real).assert_eq(ands[0], bits[0] * bits[1])
    for k in range(1, 6):
        b.when(is_real).assert_eq(ands[k], ands[k - 1] * bits[k + 1])
    b.when(is_real).when(ands[5]).assert_zero(value[0] + value[1] + value[2])


def _send_alu(b, opcode, a, bb, c, mult):
    """ZKMAirBuilder::send_alu -> send_alu_with_hi -> send_instruction (stark/src/air/builder.rs:282-324): pc = UNUSED_PC,
    hi = 0, is_sequential = 1"""
    b.send(LOOKUP_INSTRUCTION, [0, 0, UNUSED_PC, UNUSED_PC + DEFAULT_PC_INC, UNUSED_PC + 2 * DEFAULT_PC_INC, 0, opcode]
           + list(a) + list(bb) + list(c) + [0, 0, 0, 0] + [0, 0, 0, 0, 1], mult)


def jump():
    """JumpChip (crates/core/machine/src/control_flow/jump/columns.rs:11-29, air.rs:21-112): Jump / Jumpi (target in
    op_b) and JumpDirect (target = next_pc + op_b, delegated to the ADD chip through send_alu); op_a is the link address
    next_pc + 4.  66 main columns: pc, next_pc (word + range checker), next_next_pc (word + range checker), op_a, op_b,
    op_c, three opcode flags, op_a's range checker; 60 constraints, one instruction receive (is_sequential = 0: the CPU
    does not derive next_next_pc itself), one ALU send; `local_only`.  mips_costs.json: 66 + 4 * 2 + 8 = 82."""
    air = Air("Jump", main_width=66, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc, next_pc_rc = m[0], m[1:5], m[5:19]
    nnpc, nnpc_rc = m[19:23], m[23:37]
    a, bb, c = m[37:41], m[41:45], m[45:49]
    is_jump, is_jumpi, is_jumpdirect = m[49], m[50], m[51]
    a_rc = m[52:66]
    b.assert_bool(is_jump)
    b.assert_bool(is_jumpi)
    b.assert_bool(is_jumpdirect)
    is_real = is_jump + is_jumpi + is_jumpdirect
    b.assert_bool(is_real)
    opcode = is_jump * OP_JUMP + is_jumpi * OP_JUMPI + is_jumpdirect * OP_JUMPDIRECT
    b.receive(LOOKUP_INSTRUCTION, [0, 0, pc, _reduce(next_pc), _reduce(nnpc), 0, opcode] + list(a) + list(bb) + list(c)
              + [0, 0, 0, 0] + [0, 0, 0, 0, 0], is_real)
    b.when(is_real).assert_eq(_reduce(a), _reduce(next_pc) + 4)
    _koalabear_word_range_check(b, a, a_rc, is_real)
    _koalabear_word_range_check(b, next_pc, next_pc_rc, is_real)
    _koalabear_word_range_check(b, nnpc, nnpc_rc, is_real)
    for l, r in zip(nnpc, bb):
        b.when(is_jump + is_jumpi).assert_eq(l, r)
    _send_alu(b, OP_ADD, nnpc, next_pc, bb, is_jumpdirect)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_BEQ, OP_BGEZ, OP_BGTZ, OP_BLEZ, OP_BLTZ, OP_BNE = 21, 22, 23, 24, 25, 26   # Opcode, executor/src/opcode.rs:39-44


def branch():
    """BranchChip (crates/core/machine/src/control_flow/branch/columns.rs:10-49, air.rs:26-208): BEQ / BNE / BLTZ / BLEZ /
    BGTZ / BGEZ.  The comparison bits a_lt_b / a_gt_b come from the LT chip (two SLT sends), the branch target
    next_pc + c from the ADD chip (send with multiplicity is_branching); next_next_pc = target if branching else
    next_pc + 4.  62 main columns; 60 constraints; one instruction receive (op_a_immutable = 1, is_sequential = 0), three
    ALU sends, four byte range-check sends (multiplicity is_real - is_branching); `local_only`.
    mips_costs.json: 62 + 4 * 5 + 8 = 90."""
    air = Air("Branch", main_width=62, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc, next_pc_rc = m[0], m[1:5], m[5:19]
    target_pc, nnpc, nnpc_rc = m[19:23], m[23:27], m[27:41]
    a, bb, c = m[41:45], m[45:49], m[49:53]
    is_beq, is_bne, is_bltz, is_blez, is_bgtz, is_bgez = m[53:59]
    is_branching, a_gt_b, a_lt_b = m[59], m[60], m[61]
    for f in (is_beq, is_bne, is_bltz, is_bgez, is_blez, is_bgtz):
        b.assert_bool(f)
    is_real = is_beq + is_bne + is_bltz + is_bgez + is_blez + is_bgtz
    b.assert_bool(is_real)
    opcode = (is_beq * OP_BEQ + is_bne * OP_BNE + is_bltz * OP_BLTZ + is_bgez * OP_BGEZ + is_blez * OP_BLEZ
              + is_bgtz * OP_BGTZ)
    b.receive(LOOKUP_INSTRUCTION, [0, 0, pc, _reduce(next_pc), _reduce(nnpc), 0, opcode] + list(a) + list(bb) + list(c)
              + [0, 0, 0, 0] + [1, 0, 0, 0, 0], is_real)
    _koalabear_word_range_check(b, next_pc, next_pc_rc, is_real)
    _koalabear_word_range_check(b, nnpc, nnpc_rc, is_real)
    _send_alu(b, OP_ADD, target_pc, next_pc, c, is_branching)
    b.when(is_real).when_not(is_branching).assert_eq(_reduce(next_pc) + 4, _reduce(nnpc))
    _slice_range_check_u8(b, next_pc, is_real - is_branching)
    _slice_range_check_u8(b, nnpc, is_real - is_branching)
    for l, r in zip(target_pc, nnpc):
        b.when(is_real).when(is_branching).assert_eq(l, r)
    b.when_not(is_real).assert_zero(is_branching)
    b.when(is_real).assert_bool(is_branching)
    ne = a_gt_b + a_lt_b
    b.when(is_beq * is_branching).assert_zero(ne)
    b.when(is_beq).when_not(is_branching).assert_one(ne)
    b.when(is_bne * is_branching).assert_one(ne)
    b.when(is_bne).when_not(is_branching).assert_zero(ne)
    b.when(is_bltz * is_branching).assert_one(a_lt_b)
    b.when(is_bltz).when_not(is_branching).assert_zero(a_lt_b)
    b.when(is_blez * is_branching).assert_zero(a_gt_b)
    b.when(is_blez).when_not(is_branching).assert_one(a_gt_b)
    b.when(is_bgtz * is_branching).assert_one(a_gt_b)
    b.when(is_bgtz).when_not(is_branching).assert_zero(a_gt_b)
    b.when(is_bgez * is_branching).assert_zero(a_lt_b)
    b.when(is_bgez).when_not(is_branching).assert_one(a_lt_b)
    _send_alu(b, OP_SLT, [a_lt_b, 0, 0, 0], a, bb, is_real)
    _send_alu(b, OP_SLT, [a_gt_b, 0, 0, 0], bb, a, is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_SLL = 9                                                             # Opcode, executor/src/opcode.rs:26


def shift_left():
    """ShiftLeft (crates/core/machine/src/alu/sll/mod.rs:31-55 columns, :228-345 eval): a = b << (c mod 32) as a bit shift
    (multiplication of the bytes by 2^(c mod 8) with carries) followed by a byte shift.  44 main columns: pc, next_pc,
    a, b, c, c_least_sig_byte[8], shift_by_n_bits[8], bit_shift_multiplier, bit_shift_result[4],
    bit_shift_result_carry[4], shift_by_n_bytes[4], is_real; 64 constraints (the one-hot sums hold on padding rows too,
    which is why the padding row is not zero), four byte range-check sends, one instruction receive; `local_only`.
    mips_costs.json: 44 + 4 * 4 + 8 = 68."""
    air = Air("ShiftLeft", main_width=44, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    a, bb, c = m[2:6], m[6:10], m[10:14]
    c_bits, by_bits, mult = m[14:22], m[22:30], m[30]
    res, carry, by_bytes, is_real = m[31:35], m[35:39], m[39:43], m[43]
    c_byte_sum = 0
    for i in range(8):
        c_byte_sum = c_byte_sum + c_bits[i] * (1 << i)
    b.assert_eq(c_byte_sum, c[0])
    num_bits = 0
    for i in range(3):
        num_bits = num_bits + c_bits[i] * (1 << i)
    for i in range(8):
        b.when(by_bits[i]).assert_eq(num_bits, i)
    for i in range(8):
        b.when(by_bits[i]).assert_eq(mult, 1 << i)
    for i in range(4):
        v = bb[i] * mult - carry[i] * 256
        if i > 0:
            v = v + carry[i - 1]
        b.assert_eq(res[i], v)
    num_bytes = c_bits[3] + c_bits[4] * 2
    for i in range(4):
        b.when(by_bytes[i]).assert_eq(num_bytes, i)
    for k in range(4):
        for i in range(4):
            b.when(by_bytes[k]).assert_eq(a[i], 0 if i < k else res[i - k])
    for x in c_bits:
        b.assert_bool(x)
    for x in by_bits:
        b.assert_bool(x)
    total = 0
    for x in by_bits:
        total = total + x
    b.assert_eq(total, 1)
    _slice_range_check_u8(b, res, is_real)
    _slice_range_check_u8(b, carry, is_real)
    for x in by_bytes:
        b.assert_bool(x)
    total = 0
    for x in by_bytes:
        total = total + x
    b.assert_eq(total, 1)
    b.assert_bool(is_real)
    _receive_instruction(b, pc, next_pc, OP_SLL, a, bb, c, is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_SRL, OP_CLZ, OP_CLO = 10, 19, 20                                    # Opcode, executor/src/opcode.rs:27,36-37


def clo_clz():
    """CloClzChip (crates/core/machine/src/alu/clo_clz/mod.rs:28-49 columns, :137-232 eval): a = number of leading zeros
    (CLZ) or ones (CLO) of b.  bb = b or its complement; a <= 32 by a byte LTU lookup; bb = 0 gives 32, otherwise
    bb >> (31 - a) must equal 1, which is delegated to the shift-right chip (send_alu SRL).  22 main columns: pc, next_pc,
    a, b, bb, is_bb_zero, sr1, is_clz, is_clo, is_real; 20 constraints; three byte sends, one instruction receive, one ALU
    send; NOT `local_only`; the padding row is CLZ of 0 (a = 32).  mips_costs.json: 22 + 4 * 4 + 8 = 46."""
    air = Air("CloClz", main_width=22)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    a, bv, bb, is_bb_zero, sr1 = m[2:6], m[6:10], m[10:14], m[14], m[15:19]
    is_clz, is_clo, is_real = m[19], m[20], m[21]
    for x, y in zip(bv, bb):
        b.when(is_clo).assert_eq(x + y, 255)
        b.when(is_clz).assert_eq(x, y)
    _slice_range_check_u8(b, bb, is_real)
    _send_byte(b, BYTE_LTU, 1, a[0], 33, is_real)
    b.when(is_real).assert_zero(a[1])
    b.when(is_real).assert_zero(a[2])
    b.when(is_real).assert_zero(a[3])
    opcode = is_clo * OP_CLO + is_clz * OP_CLZ
    b.receive(LOOKUP_INSTRUCTION, [0, 0, pc, next_pc, next_pc + 4, 0, opcode] + list(a) + list(bv) + [0, 0, 0, 0]
              + [0, 0, 0, 0] + [0, 0, 0, 0, 1], is_real)
    b.assert_bool(is_bb_zero)
    b.when(is_bb_zero).assert_zero(_reduce(bb))
    b.when(is_bb_zero).assert_zero(bb[3])
    b.when(is_bb_zero).assert_eq(a[0], 32)
    _send_alu(b, OP_SRL, sr1, bb, [31 - a[0], 0, 0, 0], 1 - is_bb_zero)
    b.when_not(is_bb_zero).assert_one(_reduce(sr1))
    b.when_not(is_bb_zero).assert_zero(sr1[3])
    b.assert_bool(is_clo)
    b.assert_bool(is_clz)
    b.assert_one(is_clo + is_clz)
    b.eval_permutation_constraints(batch_size=2)
    return air


def byte_chip():
    """ByteChip (crates/core/machine/src/bytes/columns.rs:10-44, air.rs:22-74): the 2^16-row table that ANSWERS every byte
    lookup of the machine.  Preprocessed (12): b, c, and, or, xor, nor, sll, shr, shr_carry, ltu, msb, value_u16 for every
    byte pair (row index b * 256 + c); main (10): one multiplicity per ByteOpcode.  No constraints of its own, ten byte
    receives in opcode order.  Chip::cost counts the preprocessed columns: 12 + 10 + 4 * 6 + 8 = 54 (mips_costs.json)."""
    air = Air("Byte", main_width=10, prep_width=12)
    b = AirBuilder(air)
    mult = b.main().local()
    p = b.preprocessed().local()
    bv, cv = p[0], p[1]
    for opcode, (a1, a2, x, y) in enumerate(((p[2], 0, bv, cv), (p[3], 0, bv, cv), (p[4], 0, bv, cv), (p[6], 0, bv, cv),
                                             (0, 0, bv, cv), (p[7], p[8], bv, cv), (p[9], 0, bv, cv), (p[10], 0, bv, 0),
                                             (p[11], 0, 0, 0), (p[5], 0, bv, cv))):
        # AND, OR, XOR, SLL, U8Range, ShrCarry, LTU, MSB, U16Range, NOR (ByteOpcode::all, executor/src/events/byte.rs:165-177)
        b.receive(LOOKUP_BYTE, [opcode, a1, a2, x, y], mult[opcode])
    b.eval_permutation_constraints(batch_size=2)
    return air


def program_chip():
    """ProgramChip (crates/core/machine/src/program/mod.rs:25-37 columns, :84-94 eval): the program as a preprocessed
    table -- pc, InstructionCols {opcode, op_a, op_b[4], op_c[4], op_a_0, imm_b, imm_c} (cpu/columns/instruction.rs:12-33) --
    and one multiplicity column (how often the shard executed each instruction); a single receive_program lookup
    (air/program.rs:29-42) that answers the CPU's instruction fetches.  Cost 14 + 1 + 4 * 2 + 8 = 31 (mips_costs.json)."""
    air = Air("Program", main_width=1, prep_width=14)
    b = AirBuilder(air)
    b.receive(LOOKUP_PROGRAM, list(b.preprocessed().local()), b.main().local()[0])
    b.eval_permutation_constraints(batch_size=2)
    return air


LOOKUP_SYSCALL, LOOKUP_GLOBAL = 6, 7                                   # LookupKind, stark/src/lookup/lookup.rs:38-43


def syscall_chip(kind="Core"):
    """SyscallChip (crates/core/machine/src/syscall/chip.rs:55-70 columns, :221-293 eval), `SyscallCore` in core shards and
    `SyscallPrecompile` in precompile shards: shard, clk, syscall_id, arg1, arg2, is_real.  The core flavour RECEIVES the
    syscall from the CPU and forwards it to the Global table as a send (…, 1, 0, Syscall); the precompile flavour SENDS it
    to the precompile chip of its shard and forwards a receive (…, 0, 1, Syscall).  One boolean constraint, two lookups;
    cost 6 + 4 * 2 + 8 = 22 (mips_costs.json)."""
    assert kind in ("Core", "Precompile")
    air = Air("Syscall" + kind, main_width=6)
    b = AirBuilder(air)
    shard, clk, syscall_id, arg1, arg2, is_real = b.main().local()
    b.assert_bool(is_real)
    core = kind == "Core"
    (b.receive if core else b.send)(LOOKUP_SYSCALL, [shard, clk, syscall_id, arg1, arg2], is_real)
    b.send(LOOKUP_GLOBAL, [shard, clk, syscall_id, arg1, arg2, 0, 0, is_real * (1 if core else 0), is_real * (0 if core else 1),
                           LOOKUP_SYSCALL], is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


def memory_local():
    """MemoryLocalChip (crates/core/machine/src/memory/local.rs:28-61 columns, :204-270 eval): four memory cells per row,
    each {addr, initial_shard, final_shard, initial_clk, final_clk, initial_value[4], final_value[4], is_real}.  Per cell
    the initial state is received from the shard's memory bus and forwarded to the Global table as a receive
    (…, 0, 1, Memory), the final state is forwarded as a send (…, 1, 0, Memory) and put on the memory bus: one boolean
    constraint and four lookups per cell.  Cost 56 + 4 * 9 + 8 = 100 (mips_costs.json)."""
    air = Air("MemoryLocal", main_width=56)
    b = AirBuilder(air)
    m = b.main().local()
    for e in range(4):
        addr, ishard, fshard, iclk, fclk = m[14 * e:14 * e + 5]
        ival, fval, is_real = m[14 * e + 5:14 * e + 9], m[14 * e + 9:14 * e + 13], m[14 * e + 13]
        b.assert_bool(is_real)
        b.receive(LOOKUP_MEMORY, [ishard, iclk, addr] + list(ival), is_real)
        b.send(LOOKUP_GLOBAL, [ishard, iclk, addr] + list(ival) + [is_real * 0, is_real * 1, LOOKUP_MEMORY], is_real)
        b.send(LOOKUP_GLOBAL, [fshard, fclk, addr] + list(fval) + [is_real * 1, is_real * 0, LOOKUP_MEMORY], is_real)
        b.send(LOOKUP_MEMORY, [fshard, fclk, addr] + list(fval), is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_SRA, OP_ROR = 11, 12                                                # Opcode, executor/src/opcode.rs:28-29
BYTE_SHR_CARRY, BYTE_MSB = 5, 7                                         # ByteOpcode, executor/src/opcode.rs:195-199


def shift_right():
    """ShiftRightChip (crates/core/machine/src/alu/sr/mod.rs:40-78 columns, :266-430 eval): SRL / SRA / ROR as a byte shift of
    the 8-byte extension of b (zeros, sign bytes, or b again for the rotation) followed by a bit shift whose per-byte
    (shifted, carry) pairs come from the Byte table's ShrCarry.  71 main columns; 83 constraints; one MSB lookup, eight
    ShrCarry lookups (bytes 7..0), sixteen range-check pairs, one instruction receive; NOT `local_only`; padding rows have
    shift_by_n_bits[0] = shift_by_n_bytes[0] = 1.  mips_costs.json: 71 + 4 * 14 + 8 = 135."""
    air = Air("ShiftRight", main_width=71)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    a, bv, c = m[2:6], m[6:10], m[10:14]
    by_bits, by_bytes = m[14:22], m[22:26]
    byte_res, bit_res, carry, shifted = m[26:34], m[34:42], m[42:50], m[50:58]
    b_msb, c_bits = m[58], m[59:67]
    is_srl, is_ror, is_sra, is_real = m[67], m[68], m[69], m[70]
    _send_byte(b, BYTE_MSB, b_msb, bv[3], 0, is_real)
    c_byte_sum = 0
    for i in range(8):
        c_byte_sum = c_byte_sum + c_bits[i] * (1 << i)
    b.assert_eq(c_byte_sum, c[0])
    num_bits = 0
    for i in range(3):
        num_bits = num_bits + c_bits[i] * (1 << i)
    for i in range(8):
        b.when(by_bits[i]).assert_eq(num_bits, i)
    total = 0
    for x in by_bits:
        total = total + x
    b.assert_eq(total, 1)
    num_bytes = c_bits[3] + c_bits[4] * 2
    for i in range(4):
        b.when(by_bytes[i]).assert_eq(num_bytes, i)
    total = 0
    for x in by_bytes:
        total = total + x
    b.assert_eq(total, 1)
    ext = list(bv) + [is_sra * b_msb * 0xFF + is_ror * bv[i] for i in range(4)]
    for k in range(4):
        for i in range(8 - k):
            b.when(by_bytes[k]).assert_eq(byte_res[i], ext[i + k])
    carry_multiplier = 0
    for i in range(8):
        carry_multiplier = carry_multiplier + by_bits[i] * (1 << (8 - i))
    for i in reversed(range(8)):
        b.send(LOOKUP_BYTE, [BYTE_SHR_CARRY, shifted[i], carry[i], byte_res[i], num_bits], is_real)
    for i in reversed(range(8)):
        v = shifted[i]
        if i + 1 < 8:
            v = v + carry[i + 1] * carry_multiplier
        b.assert_eq(v, bit_res[i])
    for i in range(4):
        b.assert_eq(a[i], bit_res[i])
    for f in (is_srl, is_sra, is_ror, is_real, b_msb):
        b.assert_bool(f)
    for x in list(by_bytes) + list(by_bits) + list(c_bits):
        b.assert_bool(x)
    for long_word in (byte_res, bit_res, carry, shifted):
        _slice_range_check_u8(b, long_word, is_real)
    b.assert_bool(is_srl)
    b.assert_bool(is_sra)
    b.assert_bool(is_ror)
    b.assert_bool(is_real)
    b.assert_eq(is_srl + is_sra + is_ror, is_real)
    _receive_instruction(b, pc, next_pc, is_srl * OP_SRL + is_sra * OP_SRA + is_ror * OP_ROR, a, bv, c, is_real)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_MUL, OP_MULT, OP_MULTU = 2, 3, 4                                    # Opcode, executor/src/opcode.rs:19-21
BYTE_U16RANGE = 8
MEM_POS_HI, REG_HI = 4, 33                                             # MemoryAccessPosition::HI, the HI register


def _eval_memory_access(b, shard, clk, addr, cols, do_check):
    """MemoryAirBuilder::eval_memory_access for MemoryReadWriteCols (core/machine/src/air/memory.rs:14-137;
    memory/consistency/columns.rs:20-51): cols = prev_value[4], value[4], prev_shard, prev_clk, compare_clk,
    diff_16bit_limb, diff_8bit_limb.  The access happens after the previous one (same shard: by clk, else by shard; the
    difference minus one fits 24 bits), the previous state is consumed from the memory bus and the new one put on it."""
    prev_value, value = cols[0:4], cols[4:8]
    prev_shard, prev_clk, compare_clk, d16, d8 = cols[8:13]
    b.assert_bool(do_check)
    b.when(do_check).assert_bool(compare_clk)
    b.when(do_check).when(compare_clk).assert_eq(shard, prev_shard)
    prev_comp = compare_clk * prev_clk + (1 - compare_clk) * prev_shard
    cur_comp = compare_clk * clk + (1 - compare_clk) * shard
    b.when(do_check).assert_eq(cur_comp - prev_comp - 1, d16 + d8 * (1 << 16))
    _send_byte(b, BYTE_U16RANGE, d16, 0, 0, do_check)
    _send_byte(b, BYTE_U8RANGE, 0, 0, d8, do_check)
    b.send(LOOKUP_MEMORY, [prev_shard, prev_clk, addr] + list(prev_value), do_check)
    b.receive(LOOKUP_MEMORY, [shard, clk, addr] + list(value), do_check)


def mul():
    """MulChip (crates/core/machine/src/alu/mul/mod.rs:41-85 columns, :335-498 eval): MUL / MULT / MULTU as an 8-byte
    schoolbook product of the (sign-extended) operands with byte carries; the low word is a, the high word goes to HI,
    whose register write is checked through eval_memory_access when the event carries it.  58 main columns; 41
    constraints; two MSB lookups, eight U16 and four paired U8 range checks, the instruction receive (is_check_memory =
    hi_record_is_real) and the memory access's two range checks, bus send and bus receive; `local_only`.
    mips_costs.json: 58 + 4 * 11 + 8 = 110."""
    air = Air("Mul", main_width=58, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    hi, a, bv, cv = m[2:6], m[6:10], m[10:14], m[14:18]
    carry, product = m[18:26], m[26:34]
    b_msb, c_msb, b_sx, c_sx = m[34], m[35], m[36], m[37]
    is_mul, is_mult, is_multu, is_real = m[38], m[39], m[40], m[41]
    hi_access, hi_real, shard, clk = m[42:55], m[55], m[56], m[57]
    _send_byte(b, BYTE_MSB, b_msb, bv[3], 0, is_real)
    _send_byte(b, BYTE_MSB, c_msb, cv[3], 0, is_real)
    b.assert_eq(b_sx, is_mult * b_msb)
    b.assert_eq(c_sx, is_mult * c_msb)
    be = list(bv) + [b_sx * 0xFF] * 4
    ce = list(cv) + [c_sx * 0xFF] * 4
    mm = [0] * 8
    for i in range(8):
        for j in range(8):
            if i + j < 8:
                mm[i + j] = mm[i + j] + be[i] * ce[j]
    for i in range(8):
        b.assert_eq(product[i], mm[i] - carry[i] * 256 if i == 0 else mm[i] + carry[i - 1] - carry[i] * 256)
    has_hi = is_mult + is_multu
    for i in range(4):
        b.assert_eq(product[i], a[i])
        b.when(has_hi).assert_eq(product[i + 4], hi[i])
    for f in (b_msb, c_msb, b_sx, c_sx, is_mul, is_mult, is_multu, is_real, hi_real):
        b.assert_bool(f)
    b.when(b_sx).assert_eq(b_msb, 1)
    b.when(c_sx).assert_eq(c_msb, 1)
    b.when(is_real).assert_one(is_mul + is_mult + is_multu)
    opcode = is_mul * OP_MUL + is_mult * OP_MULT + is_multu * OP_MULTU
    for x in carry:
        _send_byte(b, BYTE_U16RANGE, x, 0, 0, is_real)
    _slice_range_check_u8(b, product, is_real)
    b.receive(LOOKUP_INSTRUCTION, [shard, clk, pc, next_pc, next_pc + 4, 0, opcode] + list(a) + list(bv) + list(cv) + list(hi)
              + [0, 0, hi_real, 0, 1], is_real)
    _eval_memory_access(b, shard, clk + MEM_POS_HI, REG_HI, hi_access, hi_real)
    b.when(hi_real).assert_one(is_mult + is_multu)
    for l, r in zip(hi, hi_access[4:8]):
        b.when(hi_real).assert_eq(l, r)
    b.when_not(hi_real).assert_zero(clk)
    b.when_not(hi_real).assert_zero(shard)
    b.eval_permutation_constraints(batch_size=2)
    return air


OP_DIV, OP_DIVU, OP_MOD, OP_MODU = 5, 6, 7, 8                          # Opcode, executor/src/opcode.rs:22-25


def _send_alu_with_hi(b, opcode, a, bb, c, hi, mult):
    """ZKMAirBuilder::send_alu_with_hi (stark/src/air/builder.rs:295-324)"""
    b.send(LOOKUP_INSTRUCTION, [0, 0, UNUSED_PC, UNUSED_PC + DEFAULT_PC_INC, UNUSED_PC + 2 * DEFAULT_PC_INC, 0, opcode]
           + list(a) + list(bb) + list(c) + list(hi) + [0, 0, 0, 0, 1], mult)


def div_rem():
    """DivRemChip (crates/core/machine/src/alu/divrem/mod.rs:38-93 columns, :375-750 eval): DIV / DIVU (quotient to op_a,
    remainder to HI through eval_memory_access) and MOD / MODU (remainder to op_a).  b = c * quotient + remainder is checked
    on 8 bytes with carries, c * quotient comes from the Mul chip (send_alu_with_hi MULT / MULTU), |remainder| < max(|c|, 1)
    from the Lt chip (SLTU), the absolute values from the AddSub chip (0 = x + |x| for negative x); division by zero gives
    quotient 0xFFFFFFFF, i32::MIN / -1 is the overflow case.  106 main columns; 126 constraints; 21 lookups; `local_only`.
    mips_costs.json: 106 + 4 * 12 + 8 = 162."""
    air = Air("DivRem", main_width=106, local_only=True)
    b = AirBuilder(air)
    m = b.main().local()
    pc, next_pc = m[0], m[1]
    bv, cv, quotient, remainder = m[2:6], m[6:10], m[10:14], m[14:18]
    abs_rem, abs_c, max_abs = m[18:22], m[22:26], m[26:30]
    ctq, carry = m[30:38], m[38:46]
    is_c_0 = m[46:57]
    is_div, is_divu, is_mod, is_modu, is_overflow = m[57:62]
    ovf_b, ovf_c = m[62:73], m[73:84]
    b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg = m[84:90]
    rcm, hi_access, shard, clk = m[90], m[91:104], m[104], m[105]
    is_real = is_div + is_divu + is_mod + is_modu
    signed = is_div + is_mod
    for msb, neg in ((b_msb, b_neg), (rem_msb, rem_neg), (c_msb, c_neg)):
        b.assert_eq(msb * signed, neg)
    _send_alu_with_hi(b, signed * OP_MULT + (is_divu + is_modu) * OP_MULTU, ctq[0:4], quotient, cv, ctq[4:8], is_real)
    # IsEqualWordOperation::eval (operations/is_equal_word.rs:32-50) against i32::MIN and -1
    for word, const, cols in ((bv, (0, 0, 0, 0x80), ovf_b), (cv, (0xFF, 0xFF, 0xFF, 0xFF), ovf_c)):
        b.assert_bool(is_real)
        _is_zero_word_operation(b, [x - k for x, k in zip(word, const)], cols, is_real)
    b.assert_eq(is_overflow, ovf_b[10] * ovf_c[10] * signed)
    sign_extension = rem_neg * 0xFF
    cqr = []
    for i in range(8):
        v = ctq[i] + (remainder[i] if i < 4 else sign_extension) - carry[i] * 256
        if i > 0:
            v = v + carry[i - 1]
        cqr.append(v)
    not_overflow = 1 - is_overflow
    for i in range(8):
        if i < 4:
            b.assert_eq(bv[i], cqr[i])
        else:
            b.when(not_overflow).when(b_neg).assert_eq(cqr[i], 0xFF)
            b.when(not_overflow).when(1 - b_neg).assert_zero(cqr[i])
            b.when(is_overflow).assert_zero(cqr[i])
    rem_byte_sum = remainder[0] + remainder[1] + remainder[2] + remainder[3]
    b.when(rem_neg).assert_one(b_neg)
    b.when(rem_byte_sum).when(1 - rem_neg).assert_zero(b_neg)
    _is_zero_word_operation(b, cv, is_c_0, is_real)
    c_zero = is_c_0[10]
    for i in range(4):
        b.when(c_zero).assert_eq(quotient[i], 0xFF)
    for i in range(4):
        b.when_not(c_neg).assert_eq(cv[i], abs_c[i])
        b.when_not(rem_neg).assert_eq(remainder[i], abs_rem[i])
    _send_alu(b, OP_ADD, [0, 0, 0, 0], cv, abs_c, c_neg)
    _send_alu(b, OP_ADD, [0, 0, 0, 0], remainder, abs_rem, rem_neg)
    want = [c_zero * 1 + (1 - c_zero) * abs_c[0]] + [(1 - c_zero) * abs_c[i] for i in range(1, 4)]
    for i in range(4):
        b.when(is_real).assert_eq(max_abs[i], want[i])
    b.assert_eq((1 - c_zero) * is_real, rcm)
    _send_alu(b, OP_SLTU, [1, 0, 0, 0], abs_rem, max_abs, rcm)
    for msb, byte in ((b_msb, bv[3]), (c_msb, cv[3]), (rem_msb, remainder[3])):
        _send_byte(b, BYTE_MSB, msb, byte, 0, is_real)
    _slice_range_check_u8(b, quotient, is_real)
    _slice_range_check_u8(b, remainder, is_real)
    for x in carry:
        b.assert_bool(x)
    _slice_range_check_u8(b, ctq, is_real)
    for f in (is_div, is_divu, is_mod, is_modu, is_overflow, b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg):
        b.assert_bool(f)
    b.when(is_real).assert_eq(1, is_divu + is_div + is_mod + is_modu)
    opcode = is_divu * OP_DIVU + is_div * OP_DIV + is_mod * OP_MOD + is_modu * OP_MODU
    b.receive(LOOKUP_INSTRUCTION, [shard, clk, pc, next_pc, next_pc + 4, 0, opcode] + list(quotient) + list(bv) + list(cv)
              + list(remainder) + [0, 0, 1, 0, 1], is_div + is_divu)
    b.receive(LOOKUP_INSTRUCTION, [0, 0, pc, next_pc, next_pc + 4, 0, opcode] + list(remainder) + list(bv) + list(cv)
              + [0, 0, 0, 0] + [0, 0, 0, 0, 1], is_mod + is_modu)
    _eval_memory_access(b, shard, clk + MEM_POS_HI, REG_HI, hi_access, is_div + is_divu)
    for l, r in zip(remainder, hi_access[4:8]):
        b.when(is_div + is_divu).assert_eq(l, r)
    b.eval_permutation_constraints(batch_size=2)
    return air


MEM_POS_C, MEM_POS_B, MEM_POS_A = 1, 2, 3                              # MemoryAccessPosition, executor/src/events/memory.rs:29-40
PV_START_PC, PV_NEXT_PC, PV_EXECUTION_SHARD = 40, 41, 44               # PublicValues<Word<T>, T>, stark/src/air/public_values.rs:17-46
CORE_NUM_PV_ELTS = 231                                                  # PROOF_MAX_NUM_PVS


def _read_cols_as_rw(cols):
    """MemoryReadCols {access} seen through MemoryCols: prev_value() = value() (memory/consistency/columns.rs:60-90)"""
    return list(cols[0:4]) + list(cols)


def cpu():
    """CpuChip (crates/core/machine/src/cpu/columns/mod.rs:16-62, air/mod.rs:20-211, air/register.rs:11-76): one row per
    executed instruction.  It fetches the instruction from the Program table, reads op_b / op_c and writes op_a through
    three eval_memory_access (registers are memory cells 0..33; clk + 2, + 1, + 3), hands the instruction with its operand
    VALUES to whichever chip implements the opcode (send_instruction), and chains shard / clk / pc from row to row and to
    the shard's public values (start_pc, next_pc, execution_shard).  67 main columns, 61 constraints, 19 lookups.
    mips_costs.json: 67 + 4 * 11 + 8 = 119."""
    air = Air("Cpu", main_width=67, num_public_values=CORE_NUM_PV_ELTS)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    pv = b.public_values()

    def cols(r):
        return dict(shard=r[0], clk16=r[1], clk8=r[2], shard_to_send=r[3], clk_to_send=r[4], pc=r[5], next_pc=r[6],
                    next_next_pc=r[7], opcode=r[8], op_a=r[9], op_b=r[10:14], op_c=r[14:18], op_a_0=r[18], imm_b=r[19],
                    imm_c=r[20], num_extra_cycles=r[21], is_rw_a=r[22], is_check_memory=r[23], is_halt=r[24],
                    is_sequential=r[25], op_a_value=r[26:30], hi_or_prev_a=r[30:34], a_access=r[34:47], b_access=r[47:56],
                    c_access=r[56:65], is_real=r[65], op_a_immutable=r[66])
    l, n = cols(m), cols(mn)
    is_real = l["is_real"]
    clk = l["clk8"] * (1 << 16) + l["clk16"]
    instruction = [l["opcode"], l["op_a"]] + list(l["op_b"]) + list(l["op_c"]) + [l["op_a_0"], l["imm_b"], l["imm_c"]]
    b.send(LOOKUP_PROGRAM, [l["pc"]] + instruction, is_real)                                   # send_program
    # eval_registers (air/register.rs)
    a_val, a_prev = l["a_access"][4:8], l["a_access"][0:4]
    b_val, c_val = l["b_access"][0:4], l["c_access"][0:4]
    for x, y in zip(b_val, l["op_b"]):
        b.when(l["imm_b"]).assert_eq(x, y)
    for x, y in zip(c_val, l["op_c"]):
        b.when(l["imm_c"]).assert_eq(x, y)
    _eval_memory_access(b, l["shard"], clk + MEM_POS_B, l["op_b"][0], _read_cols_as_rw(l["b_access"]), 1 - l["imm_b"])
    _eval_memory_access(b, l["shard"], clk + MEM_POS_C, l["op_c"][0], _read_cols_as_rw(l["c_access"]), 1 - l["imm_c"])
    for x in a_val:
        b.when(l["op_a_0"]).assert_zero(x)
    for x, y in zip(l["op_a_value"], a_val):
        b.when_not(l["op_a_0"]).assert_eq(x, y)
    for x, y in zip(l["hi_or_prev_a"], a_prev):
        b.when(l["is_rw_a"]).assert_eq(x, y)
    _eval_memory_access(b, l["shard"], clk + MEM_POS_A, l["op_a"], l["a_access"], is_real)
    _slice_range_check_u8(b, a_val, is_real)
    for x, y in zip(a_val, a_prev):
        b.when(l["op_a_immutable"]).assert_eq(x, y)
    # air/mod.rs:49-76: what is sent along with the instruction
    b.when(is_real).assert_eq(l["shard_to_send"], l["is_check_memory"] * l["shard"] + (1 - l["is_check_memory"]) * 0)
    b.when(is_real).assert_eq(l["clk_to_send"], l["is_check_memory"] * clk + (1 - l["is_check_memory"]) * 0)
    b.send(LOOKUP_INSTRUCTION, [l["shard_to_send"], l["clk_to_send"], l["pc"], l["next_pc"], l["next_next_pc"],
                                l["num_extra_cycles"], l["opcode"]] + list(l["op_a_value"]) + list(b_val) + list(c_val)
           + list(l["hi_or_prev_a"]) + [l["op_a_immutable"], l["is_rw_a"], l["is_check_memory"], l["is_halt"],
                                        l["is_sequential"]], is_real)
    # eval_shard_clk
    b.when_transition().when(n["is_real"]).assert_eq(l["shard"], n["shard"])
    _send_byte(b, BYTE_U16RANGE, l["shard"], 0, 0, is_real)
    b.when_first_row().assert_zero(clk)
    next_clk = n["clk8"] * (1 << 16) + n["clk16"]
    b.when_transition().when(n["is_real"]).assert_eq(clk + 5 + l["num_extra_cycles"], next_clk)
    b.when(is_real).assert_eq(clk, l["clk16"] + l["clk8"] * (1 << 16))                         # eval_range_check_24bits
    _send_byte(b, BYTE_U16RANGE, l["clk16"], 0, 0, is_real)
    _send_byte(b, BYTE_U8RANGE, 0, 0, l["clk8"], is_real)
    # eval_pc
    b.when(is_real).assert_eq(pv[PV_EXECUTION_SHARD], l["shard"])
    b.when_first_row().assert_eq(pv[PV_START_PC], l["pc"])
    b.when_first_row().when_not(l["is_halt"]).assert_eq(l["pc"] + 4, l["next_pc"])
    b.when_transition().when(n["is_real"]).assert_eq(l["next_pc"], n["pc"])
    b.when_transition().when(n["is_real"]).when_not(n["is_halt"]).assert_eq(l["next_next_pc"], n["next_pc"])
    b.when_transition().when(is_real).when(l["is_sequential"]).assert_eq(l["next_next_pc"], l["next_pc"] + 4)
    b.when_transition().when(is_real - n["is_real"]).assert_eq(pv[PV_NEXT_PC], l["next_pc"])
    b.when_last_row().when(is_real).assert_eq(pv[PV_NEXT_PC], l["next_pc"])
    # eval_is_real
    b.assert_bool(is_real)
    b.when_first_row().assert_one(is_real)
    b.when_transition().when_not(is_real).assert_zero(n["is_real"])
    b.when_transition().when(l["is_halt"]).assert_zero(n["is_real"])
    # padding rows carry imm_b = imm_c = is_rw_a = 1
    not_real = 1 - is_real
    b.when(not_real).assert_zero(1 - l["imm_b"])
    b.when(not_real).assert_zero(1 - l["imm_c"])
    b.when(not_real).assert_zero(1 - l["is_rw_a"])
    b.eval_permutation_constraints(batch_size=2)
    return air


def all_airs():
    return [fibonacci(), lookup_pair(), wide_bitwise(64, "wide_bitwise_64"), wide_bitwise(256, "wide_bitwise_256"),
            wide_bitwise(1024, "wide_bitwise_1024"),
            wide_bitwise(4096, "wide_bitwise_4096"), quintic(), lookup_side(True), lookup_side(False), global_tail(),
            local_bool(), add_sub(), lt(), bitwise(), poseidon2_wide(3), poseidon2_wide(9), memory_const(), base_alu(), memory_var(), ext_alu(), select(),
            batch_fri(3), exp_reverse_bits_len(3), public_values_chip(), fri_fold(3), poseidon2_skinny(9), mov_cond(), jump(), branch(), shift_left(), clo_clz(), byte_chip(), program_chip(), syscall_chip("Core"),
            syscall_chip("Precompile"), memory_local(), shift_right(), mul(), cpu(), div_rem()]
