"""PCIe probe for the streaming commit: contiguous vs strided (cudaMemcpy2DAsync) host-to-device copies of a
2^20 x 256 u32 trace from pinned memory, by slab width.  Run on the GPU box: python tools/bench/h2d_probe.py"""
import ctypes
import glob
import os
import time

import torch

so = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(so[0])
H, W = 1 << 20, 256
host = torch.empty((H, W), dtype=torch.int32).pin_memory()
host.random_(0, 1 << 30)
dev = torch.empty((H, W), dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
s = ctypes.c_void_p(stream.cuda_stream)
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def contiguous():
    rc = rt.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), H * W * 4, 1, s)
    assert rc == 0


def slabs(nc):
    def fn():
        for c0 in range(0, W, nc):
            rc = rt.cudaMemcpy2DAsync(dev.data_ptr() + c0 * H * 4, nc * 4, host.data_ptr() + c0 * 4, W * 4, nc * 4, H, 1, s)
            assert rc == 0
    return fn


ms = timed(contiguous)
print(f"contiguous 1 GiB            : {ms:7.2f} ms  {H * W * 4 / ms / 1e6:6.1f} GB/s")
for nc in (16, 32, 64, 128, 256):
    ms = timed(slabs(nc))
    print(f"2D slabs of {nc:3d} cols ({nc * 4:4d} B rows): {ms:7.2f} ms  {H * W * 4 / ms / 1e6:6.1f} GB/s")
