"""Backends the parity tests run against.

* `gpu()`  -- the product: zkmips_b200/libzkgpu.so on cuda:0, through the C ABI.  Used by `-m gpu` tests.
* `emu()`  -- TEST-ONLY: the same sources compiled by g++ against tests/emu/include/cuda_runtime.h, which
              executes kernel bodies on the CPU.  It exists so kernel index math can be debugged in a
              container without a GPU; it is never loaded by the product and no result is claimed from it.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libzkgpu_emu.so")
CSRC = os.path.join(ROOT, "zkmips_b200", "csrc")

_emu = None
_gpu = None


def build_emu(force=False):
    from zkmips_b200.air import codegen
    codegen.write()
    gen = os.path.join(CSRC, "gen")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "gen"] + [os.path.join(gen, f) for f in os.listdir(gen)] + [
        os.path.join(EMU_DIR, "emu_runtime.cpp"), os.path.join(EMU_DIR, "include", "cuda_runtime.h"),
        os.path.join(ROOT, "include", "zkgpu.h")]
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_SO) for s in srcs):
        return EMU_SO
    # the TMA-fed NTT pass (inline PTX: cp.async.bulk.tensor, mbarrier) has no CPU rendering; the emulator's runtime
    # answers "not handled" for it and the plain shared-memory pass runs instead (tests/emu/emu_runtime.cpp)
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and not f.startswith("ntt_tma")) + sorted(
        os.path.join("gen", f) for f in os.listdir(gen) if f.endswith(".cu"))
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(EMU_DIR, "include")]
    for f in cu:
        cmd += ["-x", "c++", os.path.join(CSRC, f)]
    cmd += ["-x", "c++", os.path.join(EMU_DIR, "emu_runtime.cpp"), "-o", EMU_SO]
    subprocess.check_call(cmd)
    return EMU_SO


def emu():
    global _emu
    if _emu is None:
        from zkmips_b200 import native
        _emu = native.load(build_emu()).ctx_create(0)
    return _emu


def gpu():
    global _gpu
    if _gpu is None:
        from zkmips_b200 import native
        _gpu = native.load().ctx_create(0)
    return _gpu
