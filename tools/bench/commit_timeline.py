"""Timeline (start, duration, gap) of the stage records of one commit from pinned host memory.
usage: python tools/bench/commit_timeline.py [exec|keccak|headline] [keep]   (exec: bench.EXEC21_SHAPE; keccak: 2^16 x 2 + 2^16 x 4096; headline: 2^20 x 256)"""
import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from zkmips_b200 import native
lib = native.load()
ctx = lib.ctx_create(0)
which = sys.argv[1] if len(sys.argv) > 1 else "exec"
if "keep" in sys.argv[1:]: ctx.keep_traces(True)
shape = {"exec": bench.EXEC21_SHAPE, "keccak": {"Fibonacci": (16, 2), "Wide4096": (16, 4096)}, "headline": {"Trace": (20, 256)}}[which if which in ("exec", "keccak", "headline") else "exec"]
order = sorted(shape.items(), key=lambda kv: (-kv[1][0], kv[0]))
mats = [torch.from_numpy(bench.synth_trace(lg, w, 100 + k).view(np.int32)).pin_memory().numpy().view(np.uint32) for k, (n, (lg, w)) in enumerate(order)]
one = 0x01FFFFFE
for _ in range(2):
    r, pd = ctx.commit(mats, [one]*len(mats), 1); pd.free(); ctx.sync()
ctx.prof_reset(); ctx.prof_enable(True)
torch.cuda.synchronize(); t0 = time.perf_counter()
r, pd = ctx.commit(mats, [one]*len(mats), 1)
t1 = time.perf_counter()
ctx.prof_enable(False)
print("wall %.2f ms" % ((t1 - t0) * 1e3))
end = 0.0
for name, st, ms in ctx.prof_timeline():
    gap = st - end
    print("%-10s start %8.2f  dur %7.2f  gap before %6.2f" % (name, st, ms, gap))
    end = st + ms
pd.free()
