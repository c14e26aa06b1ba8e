"""A/B: execution-shard commit wall time per call, trace retention off/on, ctx on torch's stream or its own."""
import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from zkmips_b200 import native
lib = native.load()
order = sorted(bench.EXEC21_SHAPE.items(), key=lambda kv: (-kv[1][0], kv[0]))
mats = [torch.from_numpy(bench.synth_trace(lg, w, 100 + k).view(np.int32)).pin_memory().numpy().view(np.uint32) for k, (n, (lg, w)) in enumerate(order)]
one = 0x01FFFFFE
big = torch.empty(1 << 28, dtype=torch.int32, device="cuda")
def run(ctx, label, n=4):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r, pd = ctx.commit(mats, [one]*len(mats), 1)
        t1 = time.perf_counter(); pd.free(); ctx.sync()
        ts.append(round((t1 - t0) * 1e3, 1))
    print(label, ts, flush=True)
for own in (True, False):
    ctx = lib.ctx_create(0) if own else lib.ctx_create(0, stream=torch.cuda.current_stream().cuda_stream)
    tag = "own stream " if own else "torch stream"
    run(ctx, tag + " keep off")
    ctx.keep_traces(True)
    run(ctx, tag + " keep on ")
    ctx.keep_traces(False)
    run(ctx, tag + " keep off")
    ctx.destroy()
