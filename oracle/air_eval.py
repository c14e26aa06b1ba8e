"""oracle/air_eval.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

Independent evaluation of an AIR constraint program (zkmips_b200.air.ir.Air -- the chip DEFINITION, shared
with the code generator the same way the reference shares `Air::eval` between prover and verifier):
  * quotient_values(): numpy restatement of crates/stark/src/quotient.rs:19-171 (+ folder.rs:79-102,
    selectors per crates/recursion/circuit/src/domain.rs:46-64) over canonical integers;
  * verify_constraints(): pure-Python restatement of the reference VERIFIER's identity
    crates/stark/src/verifier.rs:316-435 (eval_constraints with Horner folding folder.rs:245-249,
    recompute_quotient, selectors_at_point), which pins the prover side.
All values here are canonical (non-Montgomery) integers."""
import numpy as np

P = 0x7F000001
W = 3


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


# ---------------------------------------------------------------------------------- vectorised field
def _u(a):
    return np.asarray(a, dtype=np.uint64)


def vmul(a, b):
    return (_u(a) * _u(b)) % np.uint64(P)


def vadd(a, b):
    return (_u(a) + _u(b)) % np.uint64(P)


def vsub(a, b):
    return (_u(a) + np.uint64(P) - _u(b)) % np.uint64(P)


def vpow(a, e):
    r = np.ones_like(_u(a))
    b = _u(a).copy()
    while e:
        if e & 1:
            r = vmul(r, b)
        b = vmul(b, b)
        e >>= 1
    return r


def vinv(a):
    return vpow(a, P - 2)


class VExt:
    """vector of extension elements: 4 arrays"""

    def __init__(self, c):
        self.c = [_u(x) for x in c]

    @staticmethod
    def from_base(b):
        z = np.zeros_like(_u(b))
        return VExt([_u(b), z, z, z])

    def add(self, o):
        return VExt([vadd(x, y) for x, y in zip(self.c, o.c)])

    def sub(self, o):
        return VExt([vsub(x, y) for x, y in zip(self.c, o.c)])

    def neg(self):
        return VExt([vsub(0, x) for x in self.c])

    def mul_base(self, b):
        return VExt([vmul(x, b) for x in self.c])

    def mul(self, o):
        t = [np.zeros_like(self.c[0]) for _ in range(7)]
        for i in range(4):
            for j in range(4):
                t[i + j] = vadd(t[i + j], vmul(self.c[i], o.c[j]))
        return VExt([vadd(t[0], vmul(W, t[4])), vadd(t[1], vmul(W, t[5])), vadd(t[2], vmul(W, t[6])), t[3]])


def _lift(x, like):
    if isinstance(x, VExt):
        return x
    return VExt.from_base(np.broadcast_to(_u(x), like.shape))


def eval_rows(air, rows, sel, chal, lcs, gcs, pvs):
    """rows: dict kind -> (local, next) matrices (canonical, (n_rows, width)); perm matrices hold 4 base
    columns per extension column.  Returns the list of constraint values (arrays or VExt)."""
    n_rows = rows["main"][0].shape[0]
    like = np.zeros(n_rows, np.uint64)
    val = [None] * len(air.nodes)
    for i, node in enumerate(air.nodes):
        k = node[0]
        if k == "const":
            v = np.full(n_rows, node[1], np.uint64)
        elif k in ("main", "prep"):
            v = _u(rows[k][node[1]][:, node[2]])
        elif k == "perm":
            m = rows["perm"][node[1]]
            v = VExt([m[:, 4 * node[2] + e] for e in range(4)])
        elif k == "pv":
            v = np.full(n_rows, int(pvs[node[1]]), np.uint64)
        elif k == "gcs":
            v = np.full(n_rows, int(gcs[node[1]]), np.uint64)
        elif k == "lcs":
            v = VExt([np.full(n_rows, int(x), np.uint64) for x in lcs])
        elif k == "chal":
            v = VExt([np.full(n_rows, int(x), np.uint64) for x in chal[node[1]]])
        elif k in ("first", "last", "trans"):
            v = _u(sel[k])
        elif k == "neg":
            a = val[node[1]]
            v = a.neg() if isinstance(a, VExt) else vsub(0, a)
        else:
            a, b = val[node[1]], val[node[2]]
            if not isinstance(a, VExt) and not isinstance(b, VExt):
                v = {"add": vadd, "sub": vsub, "mul": vmul}[k](a, b)
            elif k == "mul" and not isinstance(b, VExt):
                v = a.mul_base(b)
            elif k == "mul" and not isinstance(a, VExt):
                v = b.mul_base(a)
            else:
                a, b = _lift(a, like), _lift(b, like)
                v = {"add": a.add, "sub": a.sub, "mul": a.mul}[k](b)
        val[i] = v
    return [val[c] for c in air.constraints]


def selectors_on_coset(log_n, lqd):
    """Unnormalised Lagrange selectors of the trace domain (size 2^log_n, shift 1) on the quotient domain
    GENERATOR * <g_{n+lqd}> in natural order."""
    size = 1 << (log_n + lqd)
    g = two_adic_generator(log_n + lqd)
    xs = np.empty(size, np.uint64)
    x = 3
    for i in range(size):
        xs[i] = x
        x = x * g % P
    N = 1 << log_n
    zh = vsub(vpow(xs, N), 1)
    ginv = pow(two_adic_generator(log_n), P - 2, P)
    return {"first": vmul(zh, vinv(vsub(xs, 1))), "last": vmul(zh, vinv(vsub(xs, ginv))), "trans": vsub(xs, ginv),
            "inv_zh": vinv(zh)}


def quotient_values(air, log_n, lqd, main_q, alpha, prep_q=None, perm_q=None, chal=(), lcs=(0, 0, 0, 0),
                    gcs=(0,) * 14, pvs=()):
    """main_q / prep_q / perm_q: traces on the quotient domain in NATURAL order (canonical).  Returns the
    (2^(n+lqd), 4) canonical quotient values q(x_i) = inv_zeroifier * sum_k alpha^(n-1-k) C_k(x_i)."""
    alpha = [int(x) for x in alpha]
    size = 1 << (log_n + lqd)
    step = 1 << lqd
    nxt = (np.arange(size) + step) % size
    rows = {"main": (main_q, main_q[nxt])}
    if prep_q is not None:
        rows["prep"] = (prep_q, prep_q[nxt])
    if perm_q is not None:
        rows["perm"] = (perm_q, perm_q[nxt])
    sel = selectors_on_coset(log_n, lqd)
    cons = eval_rows(air, rows, sel, chal, lcs, gcs, pvs)
    n = len(cons)
    acc = VExt([np.zeros(size, np.uint64)] * 4)
    ap = [1, 0, 0, 0]
    pows = []
    for _ in range(n):
        pows.append(ap)
        ap = ext_mul(ap, list(alpha))
    for k, c in enumerate(cons):
        a = pows[n - 1 - k]
        av = VExt([np.full(size, x, np.uint64) for x in a])
        acc = acc.add(av.mul(c) if isinstance(c, VExt) else av.mul_base(c))
    q = acc.mul_base(sel["inv_zh"])
    return np.stack(q.c, axis=1).astype(np.uint32)


# ---------------------------------------------------------------------------------- scalar extension
def ext_mul(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] = (t[i + j] + a[i] * b[j]) % P
    return [(t[0] + W * t[4]) % P, (t[1] + W * t[5]) % P, (t[2] + W * t[6]) % P, t[3]]


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
    return ext_pow(a, P ** 4 - 2)


def _e(x):
    return x if isinstance(x, list) else [x % P, 0, 0, 0]


def eval_point(air, opened, sel, chal, lcs, gcs, pvs, alpha):
    """VerifierConstraintFolder (folder.rs:151-260): every input is an extension element; Horner folding."""
    alpha = [int(x) for x in alpha]
    val = []
    for node in air.nodes:
        k = node[0]
        if k == "const":
            v = _e(node[1])
        elif k in ("main", "prep", "perm"):
            v = opened[k][node[1]][node[2]]
        elif k == "pv":
            v = _e(int(pvs[node[1]]))
        elif k == "gcs":
            v = _e(int(gcs[node[1]]))
        elif k == "lcs":
            v = [int(x) for x in lcs]
        elif k == "chal":
            v = [int(x) for x in chal[node[1]]]
        elif k in ("first", "last", "trans"):
            v = sel[k]
        elif k == "neg":
            v = ext_sub([0, 0, 0, 0], val[node[1]])
        else:
            v = {"add": ext_add, "sub": ext_sub, "mul": ext_mul}[k](val[node[1]], val[node[2]])
        val.append(v)
    acc = [0, 0, 0, 0]
    for c in air.constraints:
        acc = ext_add(ext_mul(acc, list(alpha)), val[c])
    return acc


def selectors_at_point(log_n, zeta):
    """domain.rs:46-64 with shift 1"""
    zh = ext_sub(ext_pow(zeta, 1 << log_n), [1, 0, 0, 0])
    ginv = pow(two_adic_generator(log_n), P - 2, P)
    return {"first": ext_mul(zh, ext_inv(ext_sub(zeta, [1, 0, 0, 0]))),
            "last": ext_mul(zh, ext_inv(ext_sub(zeta, [ginv, 0, 0, 0]))),
            "trans": ext_sub(zeta, [ginv, 0, 0, 0]), "inv_zh": ext_inv(zh)}


def recompute_quotient(chunk_openings, log_n, lqd, zeta):
    """verifier.rs:400-435.  chunk_openings[c] = 4 extension values (the 4 base columns of chunk c at zeta);
    chunk domain c: size N, shift GENERATOR * g_{n+lqd}^c."""
    nchunks = 1 << lqd
    g = two_adic_generator(log_n + lqd)
    shifts = [3 * pow(g, c, P) % P for c in range(nchunks)]
    N = 1 << log_n

    def zp(shift, x):  # (x / shift)^N - 1
        return ext_sub(ext_pow(ext_mul(x, _e(pow(shift, P - 2, P))), N), [1, 0, 0, 0])

    total = [0, 0, 0, 0]
    for i in range(nchunks):
        zps = [1, 0, 0, 0]
        for j in range(nchunks):
            if j != i:
                zps = ext_mul(zps, ext_mul(zp(shifts[j], zeta), ext_inv(zp(shifts[j], _e(shifts[i])))))
        for e in range(4):
            mono = [0, 0, 0, 0]
            mono[e] = 1
            total = ext_add(total, ext_mul(ext_mul(zps, mono), chunk_openings[i][e]))
    return total


def verify_constraints(air, opened, chunk_openings, log_n, lqd, zeta, alpha, chal=(), lcs=(0, 0, 0, 0), gcs=(0,) * 14,
                       pvs=()):
    """verifier.rs:316-350: folded(zeta) * inv_zeroifier(zeta) == quotient(zeta)."""
    zeta = [int(x) for x in zeta]
    sel = selectors_at_point(log_n, list(zeta))
    folded = eval_point(air, opened, sel, chal, lcs, gcs, pvs, alpha)
    return ext_mul(folded, sel["inv_zh"]) == recompute_quotient(chunk_openings, log_n, lqd, list(zeta))
