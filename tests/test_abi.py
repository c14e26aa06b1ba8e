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
