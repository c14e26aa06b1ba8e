#!/usr/bin/env python3
"""Profiling aid: runs the Poseidon2WideDeg3 device trace filler (2^18 rows, inputs resident in HBM) a few times and
prints the library's own event timing.  `ncu -k regex:poseidon2_wide_rows --set full python tools/bench/tracegen_probe.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from zkmips_b200 import native, synth  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
ctx = native.load().ctx_create(0)
inputs, _, n = synth.poseidon2_wide_events(log_n, seed=77, fill=0.9)
d_in = ctx.upload(inputs)
ctx.prof_enable(True)
for _ in range(4):
    ptr, w = ctx.tracegen_poseidon2_wide((d_in, len(inputs)), n, True)
    ctx.dev_free(ptr)
ctx.sync()
for name, ms, _ in ctx.prof_records():
    print(name, round(ms, 4), "ms", round((len(inputs) * 64 + n * w * 4) / ms / 1e6, 1), "GB/s")
