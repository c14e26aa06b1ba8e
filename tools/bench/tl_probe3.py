"""Execution-shard commit: wall vs device records, host vs device-resident, per matrix."""
import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from zkmips_b200 import native
lib = native.load()
ctx = lib.ctx_create(0)
order = sorted(bench.EXEC21_SHAPE.items(), key=lambda kv: (-kv[1][0], kv[0]))
mats = [torch.from_numpy(bench.synth_trace(lg, w, 100 + k).view(np.int32)).pin_memory().numpy().view(np.uint32) for k, (n, (lg, w)) in enumerate(order)]
one = 0x01FFFFFE
def run(fn, label, n=2):
    fn(); ctx.sync()
    ctx.prof_reset(); ctx.prof_enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.prof_enable(False)
    agg = {}
    for nme, m, l in ctx.prof_records(): agg[nme] = agg.get(nme, 0) + m / n
    print(label, "wall %.2f ms; %s ; sum %.2f" % ((t1 - t0) / n * 1e3, {k: round(v, 2) for k, v in agg.items()}, sum(agg.values())), flush=True)
def host_all():
    r, pd = ctx.commit(mats, [one]*len(mats), 1); pd.free()
run(host_all, "host all      ")
for i in (0, 1, 3, 4):
    def one_m():
        r, pd = ctx.commit([mats[i]], [one], 1); pd.free()
    run(one_m, "host only %-14s" % (order[i][0] + str(mats[i].shape)))
devs = [torch.from_numpy(m.view(np.int32)).cuda() for m in mats]
def dev_all():
    r, pd = ctx.commit_dev([d.data_ptr() for d in devs], [m.shape for m in mats], [one]*len(mats), 1); pd.free()
run(dev_all, "device all    ")
def h2d():
    for m, d in zip(mats, devs): d.copy_(torch.from_numpy(m.view(np.int32)), non_blocking=True)
run(h2d, "plain H2D     ")
