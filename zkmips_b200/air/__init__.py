"""AIR constraint IR, builder DSL and CUDA code generator for the quotient kernels.

The reference evaluates a chip's constraints by running `Air::eval` against a `ProverConstraintFolder`
(crates/stark/src/folder.rs:19-149) inside `quotient_values` (crates/stark/src/quotient.rs:19-171).  Here
the same `eval` is recorded ONCE, symbolically, into a small SSA program (ir.Air) -- what the reference's
own `get_symbolic_constraints` (crates/stark/src/machine.rs:357-362) does for constraint counting -- and
codegen.py turns that program into a straight-line sm_100a kernel per chip.
"""
from .ir import Air, AirBuilder, Expr  # noqa: F401
