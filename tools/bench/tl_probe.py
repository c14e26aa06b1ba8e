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
def run(ms, label):
    for _ in range(2):
        r, pd = ctx.commit(ms, [one]*len(ms), 1); pd.free(); ctx.sync()
    ctx.prof_reset(); ctx.prof_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r, pd = ctx.commit(ms, [one]*len(ms), 1)
    t1 = time.perf_counter()
    pd.free(); ctx.sync(); ctx.prof_enable(False)
    recs = ctx.prof_records()
    print(label, "wall %.2f ms; device records: %s ; sum %.2f" % ((t1 - t0) * 1e3, [(n, round(m, 3), l) for n, m, l in recs], sum(m for _, m, _ in recs)), flush=True)
run(mats[:1], "fib")
run(mats[1:2], "w64")
run(mats[2:], "w1k")
run(mats, "all")
# device-resident
devs = [torch.from_numpy(m.view(np.int32)).cuda() for m in mats]
for sel, label in (([0], "dev fib"), ([0, 1, 2], "dev all")):
    for _ in range(2):
        r, pd = ctx.commit_dev([devs[i].data_ptr() for i in sel], [mats[i].shape for i in sel], [one]*len(sel), 1); pd.free(); ctx.sync()
    ctx.prof_reset(); ctx.prof_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r, pd = ctx.commit_dev([devs[i].data_ptr() for i in sel], [mats[i].shape for i in sel], [one]*len(sel), 1)
    t1 = time.perf_counter()
    pd.free(); ctx.sync(); ctx.prof_enable(False)
    recs = ctx.prof_records()
    print(label, "wall %.2f ms; %s ; sum %.2f" % ((t1 - t0) * 1e3, [(n, round(m, 3), l) for n, m, l in recs], sum(m for _, m, _ in recs)), flush=True)
