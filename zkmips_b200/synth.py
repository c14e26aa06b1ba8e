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
    return M(pv)


