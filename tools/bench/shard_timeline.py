"""Timeline of the stage records of ONE shard proof (commit + quotient + open) and the host phase clocks beside it.
usage: python tools/bench/shard_timeline.py [mixed|recursion|keccak|large]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from zkmips_b200 import native

which = sys.argv[1] if len(sys.argv) > 1 else "mixed"
lib = native.load()
ctx = lib.ctx_create(0)
chips = bench._pin(torch, bench.shard_chips(which, 0))
fri = bench.FRI_PARAMS.get(which, (1, 84, 16))
w = bench.ShardWorker(ctx, chips, fri, bench.num_pv(which))
for _ in range(2):
    w.prove(chips)
w.prover.phase_ms = {}
ctx.prof_reset(); ctx.prof_enable(True)
torch.cuda.synchronize(); t0 = time.perf_counter()
w.prove(chips)
torch.cuda.synchronize(); t1 = time.perf_counter()
ctx.prof_enable(False)
print("wall %.2f ms; host phases %s" % ((t1 - t0) * 1e3, {k: round(v, 2) for k, v in w.prover.phase_ms.items()}))
recs = sorted(ctx.prof_timeline(), key=lambda r: r[1])
end = 0.0
for name, st, ms in recs:
    if ms < 0.03 and st - end < 0.05:
        end = max(end, st + ms)
        continue
    print("%-18s start %8.2f  dur %7.3f  gap before %6.2f" % (name, st, ms, st - end))
    end = max(end, st + ms)
