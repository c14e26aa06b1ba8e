"""Build libzkgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libzkgpu.so")
SOURCES = ["zkgpu.cu", "ntt_fwd.cu", "ntt_inv.cu", "ntt_tma_fwd.cu", "ntt_tma_inv.cu", "fri.cu", "quotient.cu", "tracegen.cu"]  # + csrc/gen/airs_kernels_*.cu (generated)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC", "-Xptxas", "-v",
]


def nvcc_with_sass_pass(nvcc, args, keep_dir, balance_args=()):
    """Runs the steps `nvcc <args>` would run (taken from `nvcc -dryrun --keep`), with tools/sass_balance.py applied to
    the cubin between ptxas and fatbinary.  Returns (returncode, log)."""
    import re
    import shlex
    import sys
    os.makedirs(keep_dir, exist_ok=True)
    dry = subprocess.run([nvcc, "-dryrun", "--keep", "--keep-dir", keep_dir] + list(args), capture_output=True, text=True)
    if dry.returncode:
        return dry.returncode, dry.stdout + dry.stderr
    env = dict(os.environ)
    log = ""
    balance = os.path.join(os.path.dirname(HERE), "tools", "sass_balance.py")
    for line in dry.stderr.split("\n"):
        if not line.startswith("#$ "):
            continue
        cmd = line[3:].strip()
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", cmd)
        if m and " " not in m.group(1) and not cmd.startswith(("gcc", "g++")):
            val = m.group(2).strip()
            val = re.sub(r"\$([A-Za-z_][A-Za-z0-9_]*)", lambda mm: env.get(mm.group(1), ""), val)
            env[m.group(1)] = " ".join(shlex.split(val)) if val else ""
            continue
        if cmd.startswith("fatbinary") and "-link" not in shlex.split(cmd):
            cub = re.search(r"--image3=kind=elf,sm=100a,file=([^\s\"]+)", cmd)
            if cub:
                res = subprocess.run([sys.executable, balance, cub.group(1), cub.group(1)] + list(balance_args),
                                     capture_output=True, text=True)
                log += res.stdout + res.stderr
                if res.returncode:
                    return res.returncode, log
        res = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True, env=env)
        log += res.stdout + res.stderr
        if res.returncode and not cmd.startswith("rm "):
            return res.returncode, log + "\nFAILED: " + cmd
    return 0, log


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not os.path.isdir(os.path.join(CSRC, f))]
    deps += [os.path.join(CSRC, "gen", f) for f in os.listdir(os.path.join(CSRC, "gen"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "zkgpu.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "zk_poseidon2_rc.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "tools", "sass_balance.py"))
    return deps


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _deps())


# kernels that carry the Poseidon2 round bodies and get the pipe-balancing SASS pass (tools/sass_balance.py)
BALANCE_KERNELS = r"hash_rows|compress_layer|compress_top|permute_states|grind_kernel"


def build(force=False, verbose=False, balance=None, so=None):
    """One nvcc pipeline per translation unit, in parallel, then one link.  Objects live in build/ (ignored).
    `balance` (default off; env ZK_SASS_BALANCE=1): run tools/sass_balance.py on every cubin between ptxas and
    fatbinary -- an experiment kept for the record: moving adds between the integer pipes at SASS level made the
    permutation monotonically slower in both directions (profiles/README.md).  `so` names a second library for A/B."""
    from concurrent.futures import ThreadPoolExecutor

    from .air import codegen
    codegen.write()  # csrc/gen/airs_gen.cuh (only rewritten when its text changes)
    if balance is None:
        balance = os.environ.get("ZK_SASS_BALANCE", "0") != "0"  # measured slower both ways (profiles/r2_p2bench_sass_sweep.txt)
    so = so or SO
    if not force and so == SO and not _stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build_bal" if balance else "build")
    os.makedirs(objdir, exist_ok=True)

    sources = SOURCES + sorted(os.path.join("gen", f) for f in os.listdir(os.path.join(CSRC, "gen")) if f.endswith(".cu"))

    def compile_one(src):
        base = os.path.basename(src).replace(".cu", "")
        obj = os.path.join(objdir, base + ".o")
        args = NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if balance and not src.startswith("gen"):
            rc, out = nvcc_with_sass_pass(nvcc, args, os.path.join(objdir, "keep_" + base), ["--kernels", BALANCE_KERNELS])
            return src, obj, rc, " ".join([nvcc] + args) + "  [+ sass_balance]\n" + out
        res = subprocess.run([nvcc] + args, capture_output=True, text=True)
        return src, obj, res.returncode, " ".join([nvcc] + args) + "\n" + res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources))
    log = "".join(r[3] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + [r[1] for r in results]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, "build.log" if so == SO else "build_ab.log"), "w") as fh:
        fh.write(log)
    if rc == 0 and so == SO:
        # integer-pipe budget of one permutation of the leaf sponge as compiled (bench.py: roofline_int)
        try:
            import importlib.util
            import json
            spec = importlib.util.spec_from_file_location("sass_balance", os.path.join(os.path.dirname(HERE), "tools", "sass_balance.py"))
            sb = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(sb)
            model = sb.perm_pipe_model(os.path.join(objdir, "zkgpu.o"))
            if model:
                with open(os.path.join(HERE, "pipe_model.json"), "w") as fh:
                    json.dump(model, fh, indent=1)
        except Exception as e:  # the model is reporting only
            log += f"pipe model not written: {e}\n"
    if verbose or rc:
        print(log)
    if rc:
        raise RuntimeError("nvcc failed building libzkgpu.so (see zkmips_b200/build.log)")
    return so


if __name__ == "__main__":
    build(force=True, verbose=True)
