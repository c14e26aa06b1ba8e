"""zkmips_b200 -- B200 (sm_100a) implementation of Ziren's STARK commit / quotient / FRI hot path.

The product is the C-ABI CUDA library `libzkgpu.so` (include/zkgpu.h).  This package only holds the
sources (csrc/), the build recipe (build.py) and a thin ctypes binding used by the tests and bench.py;
there is no CPU fallback: every compute entry point needs a CUDA device.
"""
from .native import Challenger, Lib, ZkError, load, pcs_open  # noqa: F401
