import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from zkmips_b200 import native
from tests import shard_util as su
lib = native.load()
ctx = lib.ctx_create(0)
chips = [su.wide_chip(16, 1024, seed=11), su.wide_chip(18, 64, seed=12), su.fibonacci_chip(20, 1, 1)]
chips = sorted(chips, key=lambda c: (-c.main.shape[0], c.name))
mats = [torch.from_numpy(c.main.view(np.int32)).pin_memory().numpy().view(np.uint32) for c in chips]
one = 0x01FFFFFE
def run(ms, label, n=6):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r, pd = ctx.commit(ms, [one]*len(ms), 1)
        t1 = time.perf_counter()
        pd.free(); ctx.sync()
        t2 = time.perf_counter()
        ts.append("%.2f(+%.2f free)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    print(label, " ".join(ts), flush=True)
run(mats, "off all    ")
ctx.keep_traces(True)
run(mats, "on  all    ")
run(mats[:2], "on fib+w64 ")
run(mats[1:], "on w64+w1k ")
run([mats[0], mats[2]], "on fib+w1k ")
run(mats[2:], "on w1k     ")
run(mats, "on  all    ")
ctx.prof_reset(); ctx.prof_enable(True)
r, pd = ctx.commit(mats, [one]*3, 1); pd.free(); ctx.sync()
print({n: round(ms, 3) for n, ms, _ in ctx.prof_records()})
