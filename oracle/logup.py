"""oracle/logup.py -- TEST INFRASTRUCTURE (CPU oracle), not product code.

numpy restatement of the reference's LogUp permutation-trace generation
(crates/stark/src/permutation.rs:29-69 populate_local_permutation_row, :102-196 generate_permutation_trace):
  entry[b] = sum over the lookups of batch b of  (+-multiplicity) / (alpha + kind + sum_k beta^(k+1) value_k)
  last column = inclusive prefix sum over the rows of sum_b entry[b];  local cumulative sum = its last value.
All values canonical."""
import numpy as np

from .air_eval import P, VExt, vadd, vmul, vsub, vinv, _u

W = 3


def vext_inv(a: VExt) -> VExt:
    """inverse through the norm to F_p[Y]/(Y^2 - 3), Y = X^2 (any method gives the same field element)"""
    a0, a1, a2, a3 = a.c
    t3 = lambda x: vmul(W, x)
    n0 = vsub(vadd(vmul(a0, a0), t3(vmul(a2, a2))), t3(vmul(2, vmul(a1, a3))))
    n1 = vsub(vmul(2, vmul(a0, a2)), vadd(vmul(a1, a1), t3(vmul(a3, a3))))
    d = vinv(vsub(vmul(n0, n0), t3(vmul(n1, n1))))
    i0, i1 = vmul(n0, d), vsub(0, vmul(n1, d))
    A0 = vadd(vmul(a0, i0), t3(vmul(a2, i1)))
    A1 = vadd(vmul(a0, i1), vmul(a2, i0))
    B0 = vadd(vmul(a1, i0), t3(vmul(a3, i1)))
    B1 = vadd(vmul(a1, i1), vmul(a3, i0))
    return VExt([A0, vsub(0, B0), A1, vsub(0, B1)])


def _apply(lf, prep, main, h):
    c, terms = lf
    acc = np.full(h, c, np.uint64)
    for (t, col, w) in terms:
        src = main if t == "main" else prep
        acc = vadd(acc, vmul(_u(src[:, col]), w))
    return acc


def generate_permutation_trace(air, prep, main, alpha, beta):
    h = main.shape[0]
    lookups = [(l, True) for l in air.sends] + [(l, False) for l in air.receives]
    width = air.permutation_width
    if width == 0:
        return np.zeros((h, 0), np.uint32), [0, 0, 0, 0]
    full = lambda e: VExt([np.full(h, int(x), np.uint64) for x in e])
    A, B = full(alpha), full(beta)
    cols = []
    total = VExt([np.zeros(h, np.uint64)] * 4)
    bs = air.batch_size
    for b in range(width - 1):
        entry = VExt([np.zeros(h, np.uint64)] * 4)
        for l, is_send in lookups[b * bs:(b + 1) * bs]:
            den = A.add(VExt.from_base(np.full(h, l["kind"], np.uint64)))
            bp = B
            for lf in l["values"]:
                den = den.add(bp.mul_base(_apply(lf, prep, main, h)))
                bp = bp.mul(B)
            m = _apply(l["mult"], prep, main, h)
            if not is_send:
                m = vsub(0, m)
            entry = entry.add(vext_inv(den).mul_base(m))
        cols.append(entry)
        total = total.add(entry)
    # inclusive scan over rows
    phi = [np.cumsum(c.astype(object)) % P for c in total.c]
    phi = VExt([np.array(x, dtype=np.uint64) for x in phi])
    cols.append(phi)
    out = np.zeros((h, 4 * width), np.uint32)
    for j, e in enumerate(cols):
        for k in range(4):
            out[:, 4 * j + k] = e.c[k].astype(np.uint32)
    return out, [int(phi.c[k][-1]) for k in range(4)]
