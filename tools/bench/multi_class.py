"""One-off measurement: leaf hashing of a height class made of several chips of arbitrary widths
(hash_rows_multi) versus a single matrix of the same total width rounded to a multiple of 8 (hash_rows_w8)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from zkmips_b200 import native
from tests import util

ctx = native.load().ctx_create(0)
log_n = 19
ws = [119, 47, 115]
mats = [util.monty(util.splitmix64(7 + i, (1 << log_n) * w).reshape(1 << log_n, w)) for i, w in enumerate(ws)]
one = 0x01FFFFFE
for name, ms_, sh in (("3 chips 119+47+115", mats, [one] * 3),
                      ("1 matrix 280", [util.monty(util.splitmix64(3, (1 << log_n) * 280).reshape(1 << log_n, 280))], [one])):
    for it in range(3):
        ctx.prof_reset(); ctx.prof_enable(True)
        t = time.perf_counter()
        root, pd = ctx.commit(ms_, sh, 1)
        dt = time.perf_counter() - t
        ctx.prof_enable(False)
        st = {}
        for n, ms, _ in ctx.prof_records():
            st[n] = st.get(n, 0) + ms
        pd.free()
    print(name, f"wall {dt*1e3:.2f} ms", {k: round(v, 3) for k, v in st.items()})
