"""Shared helpers for the parity tests: seeded inputs and a pure-Python field for tiny cases."""
import numpy as np

P = 0x7F000001
R = 1 << 32


from zkmips_b200.synth import config2_trace, splitmix64  # noqa: E402,F401


def canon_matrix(h, w, kind="rand", seed=0x5A4B4D49):
    """Canonical (non-Montgomery) test matrix: 'index' = (r*w + c) mod p (mirrors the deterministic
    inputs of recursion/circuit/src/fri.rs:829-836), 'rand' = splitmix64."""
    if kind == "index":
        return (np.arange(h * w, dtype=np.uint64) % P).astype(np.uint32).reshape(h, w)
    return splitmix64(seed + 7919 * h + w, h * w).reshape(h, w)


def monty(a):
    """canonical -> Montgomery, numpy only (no oracle involved)."""
    a = np.asarray(a, dtype=np.uint64)
    return ((a << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def unmonty(a):
    rinv = pow(R, -1, P)
    a = np.asarray(a, dtype=np.uint64)
    return ((a * np.uint64(rinv % P)) % np.uint64(P)).astype(np.uint32) if False else \
        np.array([(int(x) * rinv) % P for x in a.reshape(-1)], dtype=np.uint32).reshape(a.shape)


def two_adic_generator(bits):
    g = pow(3, 127, P)
    for _ in range(bits, 24):
        g = g * g % P
    return g


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


# ---- degree-4 extension over canonical ints (X^4 = 3), for tiny independent checks
def ext_mul(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] = (t[i + j] + a[i] * b[j]) % P
    return [(t[0] + 3 * t[4]) % P, (t[1] + 3 * t[5]) % P, (t[2] + 3 * t[6]) % P, t[3]]


def ext_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def ext_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def ext_pow(a, e):
    r = [1, 0, 0, 0]
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


def ext_inv(a):
    # a^(p^4 - 2)
    return ext_pow(a, P ** 4 - 2)
