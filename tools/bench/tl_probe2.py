"""Stage records of a host commit of the keccak-like chip (2^16 x 4096), with and without trace retention."""
import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from zkmips_b200 import native
from tests import shard_util as su
lib = native.load()
ctx = lib.ctx_create(0)
chips = [su.wide_chip(16, 4096, seed=11), su.fibonacci_chip(16, 1, 1)]
chips = sorted(chips, key=lambda c: (-c.main.shape[0], c.name))
mats = [torch.from_numpy(c.main.view(np.int32)).pin_memory().numpy().view(np.uint32) for c in chips]
print([ (c.name, m.shape) for c, m in zip(chips, mats)])
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
    agg = {}
    for n, m, l in recs: agg[n] = agg.get(n, 0) + m
    print(label, "wall %.2f ms; %d records; %s ; sum %.2f" % ((t1 - t0) * 1e3, len(recs), {k: round(v, 2) for k, v in agg.items()}, sum(agg.values())), flush=True)
    print("   leaf_hash records:", [round(m, 2) for n, m, l in recs if n == "leaf_hash"])
run(mats, "keep off")
ctx.keep_traces(True)
run(mats, "keep on ")
for mb in (64, 128, 512):
    os.environ["ZK_SLAB_MB"] = str(mb)
    c2 = lib.ctx_create(0)
    ctx, old = c2, ctx
    run(mats, "slab %d MB, keep off" % mb)
    ctx = old
