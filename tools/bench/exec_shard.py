"""Commit of the maximal log-21 execution shard alone (bench.py's `exec_shard_commit` leg), for A/B runs:
   ZK_EVEN_PITCH=0 python tools/bench/exec_shard.py      # dense odd-pitch LDEs
   python tools/bench/exec_shard.py --timeline           # plus the stage timeline of the last commit"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from zkmips_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--timeline", action="store_true")
args = ap.parse_args()
ctx = native.load().ctx_create(0)
out = bench.exec_shard_leg(ctx, torch, args)
print(json.dumps(out))
if args.timeline:
    recs = ctx.prof_timeline()
    per = len(recs) // 6
    t0 = recs[-per][1]
    for name, start, ms in recs[-per:]:
        print(f"{start - t0:9.3f} ms  +{ms:7.3f}  {name}")
