"""What caps the aggregate host->device rate on the multi-GPU box (task 6b/6c)?

1. Concurrent raw H2D of 1 GiB pinned buffers on GPU subsets (one host thread + one stream per GPU): 0 alone, the pairs
   0+1, 0+2, 0+4, the quads and all GPUs -- pairs that share a PCIe switch uplink drop together, pairs that do not keep
   their single rate; if EVERY pair drops the same way the cap is host memory / the root complex, not a switch.
2. Staging through an idle GPU: half of GPU 0's 1 GiB goes up directly, the other half goes to GPU k over ITS PCIe link
   and is forwarded to GPU 0 over NVLink (peer copy), in 64 MiB chunks pipelined on a second stream.  If GPU 0 receives
   its GiB faster than alone, N < 8 ranks can borrow idle links.
Prints `lspci -tv` and `nvidia-smi topo -m` for the record.  Run: python tools/bench/h2d_multi_probe.py"""
import subprocess
import threading
import time

import torch

GIB = 1 << 30
n = torch.cuda.device_count()
print(f"{n} GPUs: {torch.cuda.get_device_name(0)}")
hosts = [torch.empty(GIB // 4, dtype=torch.int32).pin_memory() for _ in range(n)]
for h in hosts:
    h.random_(0, 1 << 30)
devs = [torch.empty(GIB // 4, dtype=torch.int32, device=f"cuda:{g}") for g in range(n)]
streams = [torch.cuda.Stream(device=g) for g in range(n)]


def run(gpus, reps=4):
    """aggregate GB/s of `reps` back-to-back 1 GiB H2D copies on every GPU of `gpus`, started together"""
    bar = threading.Barrier(len(gpus) + 1)
    times = {}

    def work(g):
        torch.cuda.set_device(g)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(streams[g]):
            devs[g].copy_(hosts[g], non_blocking=True)  # warm up
            streams[g].synchronize()
            bar.wait()
            a.record(streams[g])
            for _ in range(reps):
                devs[g].copy_(hosts[g], non_blocking=True)
            b.record(streams[g])
            streams[g].synchronize()
        times[g] = a.elapsed_time(b) / reps

    th = [threading.Thread(target=work, args=(g,)) for g in gpus]
    for t in th:
        t.start()
    bar.wait()
    for t in th:
        t.join()
    per = {g: GIB / (ms * 1e6) for g, ms in times.items()}
    agg = len(gpus) * GIB / (max(times.values()) * 1e6)
    print(f"H2D on GPUs {gpus}: aggregate {agg:6.1f} GB/s; per GPU " + ", ".join(f"{g}:{v:.1f}" for g, v in sorted(per.items())))
    return agg


run([0])
if n >= 2:
    for other in [g for g in (1, 2, 3, 4, 7) if g < n]:
        run([0, other])
if n >= 4:
    run([0, 1, 2, 3])
    if n >= 8:
        run([0, 2, 4, 6])
        run([4, 5, 6, 7])
        run(list(range(8)))


def staged(helper, chunk=64 << 20, reps=3):
    """GPU 0 receives 1 GiB: first half direct, second half through GPU `helper` (H2D there, then peer copy)."""
    half = GIB // 8  # elements
    per_chunk = chunk // 4
    s0 = streams[0]
    sh = streams[helper]
    sp = torch.cuda.Stream(device=helper)  # peer-copy stream
    stage = devs[helper]
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(helper)
        t = time.perf_counter()
        with torch.cuda.stream(s0):
            devs[0][:half].copy_(hosts[0][:half], non_blocking=True)
        evs = []
        for off in range(0, half, per_chunk):
            with torch.cuda.stream(sh):
                stage[off:off + per_chunk].copy_(hosts[0][half + off:half + off + per_chunk], non_blocking=True)
                e = torch.cuda.Event()
                e.record(sh)
            with torch.cuda.stream(sp):
                sp.wait_event(e)
                devs[0][half + off:half + off + per_chunk].copy_(stage[off:off + per_chunk], non_blocking=True)
        s0.synchronize()
        sp.synchronize()
        best = min(best, time.perf_counter() - t)
    print(f"GPU 0 <- 1 GiB, half staged through GPU {helper} + NVLink: {best * 1e3:6.2f} ms = {GIB / best / 1e9:6.1f} GB/s into GPU 0")


if n >= 2:
    for helper in [g for g in (1, 2, 4) if g < n]:
        try:
            staged(helper)
        except Exception as e:  # peer access not available
            print(f"staging through GPU {helper} failed: {e}")

for cmd in (["nvidia-smi", "topo", "-m"], ["lspci", "-tv"], ["nproc"], ["numactl", "-H"]):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
        print("$ " + " ".join(cmd))
        print(out[:6000])
    except Exception as e:
        print("$ " + " ".join(cmd), "->", e)
