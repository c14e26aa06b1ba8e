"""The C-ABI library loads without a GPU and exports every symbol include/zkgpu.h declares; the ctypes
prototype table mirrors the header; a context cannot be created without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from zkmips_b200 import ZkError, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "zkgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", src))


def test_header_symbols_are_exported():
    from zkmips_b200 import build
    build.build()
    lib = native.load()
    names = _declared()
    assert len(names) >= 40
    for n in sorted(names):
        assert hasattr(lib.dll, n), f"{n} declared in include/zkgpu.h but not exported"


def test_prototype_table_matches_header():
    assert set(native.PROTOTYPES) == _declared()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = native.load()
    with pytest.raises(ZkError, match="no CUDA device"):
        lib.ctx_create(0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ZkError, match="missing"):
        native.Lib(str(tmp_path / "libzkgpu.so"))


def _c_prototypes():
    """{function: number of parameters} parsed from include/zkgpu.h"""
    src = open(os.path.join(ROOT, "include", "zkgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(zk_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        args = " ".join(args.split())
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_sys_crate_declares_the_whole_abi():
    """rust/gpu-sys/src/lib.rs (the `-sys` crate a Ziren maintainer adds; not compiled here -- no Rust toolchain) must
    declare every function of the header with the same number of parameters, and the two #[repr(C)] structs must have
    the header's fields in the header's order."""
    rs = open(os.path.join(ROOT, "rust", "gpu-sys", "src", "lib.rs")).read()
    block = rs[rs.index('extern "C-unwind" {'):]
    rust = {}
    for name, args in re.findall(r"pub fn (zk_[a-z0-9_]+)\(([^)]*)\)", block):
        rust[name] = len(re.findall(r"(?:^|, )[a-z_0-9]+: ", args))
    c = _c_prototypes()
    assert set(rust) == set(c), sorted(set(rust) ^ set(c))
    for name, n in c.items():
        assert rust[name] == n, f"{name}: header has {n} parameters, gpu-sys declares {rust[name]}"
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "zkgpu.h")).read(), flags=re.S)
    desc = re.search(r"typedef struct \{([^}]*)\} zk_air_desc;", hdr).group(1)
    c_fields = re.findall(r"\b([a-z_]+)\s*[,;]", desc)
    rs_fields = re.findall(r"pub ([a-z_]+): u32", rs[rs.index("pub struct ZkAirDesc"):rs.index('#[link(name = "zkgpu")]')])
    assert c_fields == rs_fields
    ch = re.search(r"typedef struct \{([^}]*)\} zk_challenger;", hdr).group(1)
    assert re.findall(r"uint32_t ([a-z_]+)", ch) == ["state", "in", "n_in", "out", "n_out"]
    assert re.findall(r"pub ([a-z_]+): ", rs[rs.index("pub struct ZkChallenger"):rs.index("pub struct ZkAirDesc")]) == \
        ["state", "inp", "n_in", "out", "n_out"]
