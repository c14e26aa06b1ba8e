"""The constraint IR: JSON round trip of the real chips (the Rust-side exporter's target format) and the transcribed
chips' shapes against the reference's cost table."""
import json
import os

from zkmips_b200.air import codegen, library
from zkmips_b200.air.ir import Air

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPORTED = os.path.join(ROOT, "zkmips_b200", "air", "exported")


def test_exported_json_reproduces_the_generated_kernels():
    """Loading zkmips_b200/air/exported/<Chip>.json (what a recording-builder exporter would write) must give the same
    program as the hand transcription: identical JSON again, and byte-identical generated CUDA."""
    for make in (library.add_sub, library.lt, library.bitwise):
        air = make()
        text = open(os.path.join(EXPORTED, air.name + ".json")).read()
        assert json.loads(text) == json.loads(air.to_json()), f"{air.name}.json is stale: run tools/export_airs.py"
        loaded = Air.from_json(text)
        assert loaded.to_json() == air.to_json()
        assert loaded.commit_scope == air.commit_scope and loaded.local_only == air.local_only
        assert codegen.generate([loaded]) == codegen.generate([air])


def test_real_chip_shapes_match_mips_costs():
    """committed columns per row = main + 4 * permutation + 4 * quotient chunks (Chip::cost, crates/stark/src/chip.rs:
    151-162) must equal crates/core/executor/src/artifacts/mips_costs.json: AddSub 47, Lt 56, Bitwise 42; constraint
    counts as StarkMachine::setup computes them (own + count_permutation_constraints, permutation.rs:355-388)."""
    want = {"AddSub": (19, 47, 14, 8), "Lt": (36, 56, 32, 4), "Bitwise": (18, 42, 5, 5)}
    for make in (library.add_sub, library.lt, library.bitwise):
        air = make()
        width, cost, own, n_lookups = want[air.name]
        assert air.main_width == width
        assert air.main_width + 4 * air.perm_width + 4 * 2 == cost
        assert len(air.sends) + len(air.receives) == n_lookups
        assert air.perm_width == -(-n_lookups // 2) + 1
        assert air.num_constraints == own + (air.perm_width - 1) + 3
        assert air.local_only and air.commit_scope == "local"
        assert air.max_degree() == 3  # log_quotient_degree 1
