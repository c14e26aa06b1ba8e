#!/usr/bin/env python3
"""Generate tests/golden/poseidon2_vectors.json from the REFERENCE's own C++ sources
(oracle/_ref/libzkref.so = crates/core/machine/include/kb31_t.hpp + crates/recursion/core/include/
poseidon2*.hpp compiled in place) plus the reference's known-answer digest
(examples/poseidon2/host/src/main.rs:33-37).  Run here, where /root/reference exists; the JSON is
committed because the reference does not travel to the GPU box."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

P = ob.P
R = ob.ref()
assert R is not None, "needs /root/reference to build oracle/_ref"
rng = np.random.default_rng(20261018)


def ref_permute_canonical(v):
    s = np.array([R.ref_to_monty(int(x) % P) for x in v], np.uint32)
    R.ref_poseidon2_permute(s.ctypes.data_as(ob._u32p))
    return [R.ref_from_monty(int(x)) for x in s]


def ref_sponge(data: bytes) -> str:
    l = len(data)
    new = (l + 3) // 3 * 3
    pad = bytearray(data) + bytes(new - l)
    if l % 3 == 2:
        pad[l] = 0b10000001
    else:
        pad[l] = 1
        pad[new - 1] = 0b10000000
    felts = [int.from_bytes(pad[i:i + 3], "little") for i in range(0, new, 3)]
    st = [0] * 16
    for i in range(0, len(felts), 8):
        chunk = felts[i:i + 8]
        st[:len(chunk)] = chunk
        st = ref_permute_canonical(st)
    return np.array(st, "<u4").tobytes()[:32].hex()


out = {"source": "oracle/_ref/libzkref.so built from /root/reference (see tools/gen_golden.py)",
       "permute_canonical": [], "field_mul_monty": [], "field_inv_monty": [], "sponge_hex": []}
ins = [list(range(16)), [1] * 16, [0] * 16, [P - 1] * 16] + [rng.integers(0, P, 16).tolist() for _ in range(12)]
for v in ins:
    out["permute_canonical"].append({"in": [int(x) for x in v], "out": ref_permute_canonical(v)})
for _ in range(64):
    a, b = (int(x) for x in rng.integers(0, P, 2))
    out["field_mul_monty"].append({"a": a, "b": b, "out": R.ref_mul(a, b)})
for a in [1, 2, P - 1, 0x01FFFFFE] + [int(x) for x in rng.integers(1, P, 28)]:
    out["field_inv_monty"].append({"a": a, "out": R.ref_inv(a)})
kat = ref_sponge(bytes([1] * 1000))
assert kat == "ae45b14fe23b9f584c76c67d4d9ef6635a27b553a7114427584cc87ba8919866", kat
out["sponge_hex"].append({"in": (bytes([1] * 1000)).hex(), "out": kat})
for n in (0, 1, 2, 23, 24, 25, 100):
    d = bytes(rng.integers(0, 256, n).astype(np.uint8))
    out["sponge_hex"].append({"in": d.hex(), "out": ref_sponge(d)})
path = os.path.join(ROOT, "tests", "golden", "poseidon2_vectors.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", path)
