#!/usr/bin/env python3
"""Writes the constraint programs of the real Ziren chips as JSON (zkmips_b200/air/exported/<Chip>.json): the format a
Rust-side recording builder (ZKMAirBuilder + PairBuilder + MultiTableAirBuilder running `Chip::eval` once,
crates/stark/src/chip.rs:253-272, the way the reference already runs it symbolically to count constraints,
crates/stark/src/machine.rs:357-369) has to emit.  tests/test_air_ir.py checks that loading these files reproduces the
generated CUDA text."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zkmips_b200.air import library  # noqa: E402

OUT = os.path.join(ROOT, "zkmips_b200", "air", "exported")
os.makedirs(OUT, exist_ok=True)
for air in (library.add_sub(), library.lt(), library.bitwise(), library.poseidon2_wide(3), library.poseidon2_wide(9),
            library.memory_const(), library.base_alu(), library.memory_var(), library.ext_alu(), library.select(),
            library.batch_fri(3), library.exp_reverse_bits_len(3), library.public_values_chip(), library.fri_fold(3),
            library.poseidon2_skinny(9), library.mov_cond(), library.jump(), library.branch(), library.shift_left(),
            library.clo_clz(), library.byte_chip(), library.program_chip(),
            library.syscall_chip("Core"), library.syscall_chip("Precompile"), library.memory_local(),
            library.shift_right(), library.mul(), library.cpu(),
            library.div_rem()):
    with open(os.path.join(OUT, air.name + ".json"), "w") as fh:
        fh.write(air.to_json() + "\n")
    print(air.name, air.num_constraints, "constraints,", len(air.nodes), "nodes")
