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
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def commit_all():
    r, pd = ctx.commit(mats, [one]*3, 1); pd.free()
print("shapes", [m.shape for m in mats])
print("commit all (keep_traces off): %.2f ms" % t(commit_all))
for i, m in enumerate(mats):
    def one_m():
        r, pd = ctx.commit([m], [one], 1); pd.free()
    print("  commit only %s: %.2f ms" % (m.shape, t(one_m)))
devs = [torch.from_numpy(m.view(np.int32)).cuda() for m in mats]
def commit_dev():
    r, pd = ctx.commit_dev([d.data_ptr() for d in devs], [m.shape for m in mats], [one]*3, 1); pd.free()
print("commit_dev all: %.2f ms" % t(commit_dev))
ctx.keep_traces(True)
print("commit all (keep_traces on): %.2f ms" % t(commit_all))
for i, m in enumerate(mats):
    def one_m():
        r, pd = ctx.commit([m], [one], 1); pd.free()
    print("  keep_traces on, commit only %s: %.2f ms" % (m.shape, t(one_m)))
def h2d():
    for m, d in zip(mats, devs): d.copy_(torch.from_numpy(m.view(np.int32)), non_blocking=True)
print("plain H2D of the three traces: %.2f ms" % t(h2d))
ctx.prof_reset(); ctx.prof_enable(True)
commit_all(); torch.cuda.synchronize()
print({n: round(ms, 3) for n, ms, _ in ctx.prof_records()})
