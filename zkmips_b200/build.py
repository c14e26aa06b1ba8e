"""Build libzkgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libzkgpu.so")
SOURCES = ["zkgpu.cu", "ntt_fwd.cu", "ntt_inv.cu", "fri.cu", "quotient.cu"]  # + csrc/gen/airs_kernels_*.cu (generated)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC", "-Xptxas", "-v",
]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not os.path.isdir(os.path.join(CSRC, f))]
    deps += [os.path.join(CSRC, "gen", f) for f in os.listdir(os.path.join(CSRC, "gen"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "zkgpu.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "zk_poseidon2_rc.h"))
    return deps


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    """One nvcc per translation unit, in parallel, then one link.  Objects live in csrc/build/ (ignored)."""
    from concurrent.futures import ThreadPoolExecutor

    from .air import codegen
    codegen.write()  # csrc/gen/airs_gen.cuh (only rewritten when its text changes)
    if not force and not _stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    sources = SOURCES + sorted(os.path.join("gen", f) for f in os.listdir(os.path.join(CSRC, "gen")) if f.endswith(".cu"))

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src).replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources))
    log = "".join(r[3] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + [r[1] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(log)
    if verbose or rc:
        print(log)
    if rc:
        raise RuntimeError("nvcc failed building libzkgpu.so (see zkmips_b200/build.log)")
    return SO


if __name__ == "__main__":
    build(force=True, verbose=True)
