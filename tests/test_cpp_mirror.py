"""The C++ host mirror (include/zkgpu.hpp) compiles, drives a Fibonacci shard end to end through the C ABI
and produces the same transcript as the Python mirror + oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from oracle import binding_fri as bf
from tests import backends, shard_util as su

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")


def _build(lib_path, out):
    libdir, libname = os.path.split(lib_path)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir, "-pthread"])
    return out


def _run_and_check(exe, log_n, nq, pw):
    res = subprocess.run([exe, str(log_n), str(nq), str(pw)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    got = json.loads(res.stdout)
    # oracle transcript for the same shard
    chip = su.fibonacci_chip(log_n)
    tree = ob.pcs_commit([chip.main], 1)
    assert got["main_commit"] == [int(x) for x in tree.root]
    # Dft + Mmcs used separately give the same tree as Pcs::commit
    assert got["mmcs_root_of_lde"] == [int(x) for x in tree.root]
    assert got["opened_row_5"] == [int(x) for x in tree.matrix(0)[5]]
    assert got["path_len"] == log_n + 1
    ch = bf.new_challenger()
    bf.observe(ch, su.M(chip.pvs))
    bf.observe(ch, tree.root)
    bf.sample_ext(ch)
    bf.sample_ext(ch)
    alpha = bf.sample_ext(ch)
    assert got["alpha"] == [int(x) for x in alpha]
    assert got["n_layers"] == log_n and got["n_queries"] == nq
    assert len(got["challenger_state"]) == 16
    return got


def test_cpp_mirror_on_emulator(tmp_path):
    exe = _build(backends.build_emu(), str(tmp_path / "host_mirror_emu"))
    _run_and_check(exe, 6, 6, 5)


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    from zkmips_b200 import native
    exe = _build(native.DEFAULT_SO, str(tmp_path / "host_mirror_gpu"))
    _run_and_check(exe, 12, 84, 16)
