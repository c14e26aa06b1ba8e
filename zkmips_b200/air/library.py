"""AIRs compiled into libzkgpu.  Real Ziren chips need the Rust-side exporter; these are
the synthetic ones BASELINE.md section 4 names."""
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


def all_airs():
    return [fibonacci(), lookup_pair(), wide_bitwise(64, "wide_bitwise_64"), wide_bitwise(1024, "wide_bitwise_1024"),
            wide_bitwise(4096, "wide_bitwise_4096"), quintic(), lookup_side(True), lookup_side(False), global_tail(),
            local_bool()]
