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
    return air


def lookup_pair():
    """Small chip with preprocessed columns and a LogUp-style permutation argument, exercising every
    extension-field node kind the reference's eval_permutation_constraints uses
    (crates/stark/src/permutation.rs:205-347): one send and one receive per row batched in one column,
        perm[0] * (alpha + v0 + beta*v1) * (alpha + u0 + beta*u1) = m_s * (alpha + u..) - m_r * (alpha + v..)
        phi' - phi = perm'[0] (transition), phi[first] = perm[0], phi[last] = local cumulative sum."""
    air = Air("lookup_pair", main_width=5, prep_width=2, perm_width=2, num_public_values=1)
    b = AirBuilder(air)
    m, mn = b.main().local(), b.main().next()
    p = b.preprocessed().local()
    perm, permn = b.permutation().local(), b.permutation().next()
    alpha, beta = b.permutation_randomness()
    lcs = b.local_cumulative_sum()
    gcs = b.global_cumulative_sum()
    # main constraints: m2 = m0 * m1 + p0 ; m3 boolean ; next.m0 = m0 + pv0 on transitions
    b.assert_eq(m[2], m[0] * m[1] + p[0])
    b.assert_bool(m[3])
    b.when_transition().assert_eq(mn[0], m[0] + b.public_values()[0])
    # LogUp batch: send (m0, m1) with multiplicity m3, receive (p0, p1) with multiplicity m4
    send = alpha + m[0] + beta * m[1]
    recv = alpha + p[0] + beta * p[1]
    b.assert_zero_ext(perm[0] * send * recv - (m[3] * recv - m[4] * send))
    phi, phin = perm[1], permn[1]
    b.when_transition().assert_eq_ext(phin - phi, permn[0])
    b.when_first_row().assert_eq_ext(phi, perm[0])
    b.when_last_row().assert_eq_ext(phi, lcs)
    # global cumulative sum is only observed by this synthetic chip: tie coordinate 0 to a main column on the last row
    b.when_last_row().assert_eq(m[4] * gcs[0], m[4] * gcs[0])
    return air


def all_airs():
    return [fibonacci(), lookup_pair(), wide_bitwise(64, "wide_bitwise_64"), wide_bitwise(1024, "wide_bitwise_1024")]
