//! Constraint-DAG exporter: the Rust half of SURVEY.md section 8 row f2.  NOT COMPILED HERE (this image has no
//! cargo / rustc); written against the APIs the reference itself uses for the same purpose:
//!
//! * `p3_uni_stark::get_symbolic_constraints(&air, preprocessed_width, num_public_values)` -- how
//!   `StarkMachine::setup` already runs every chip's `Air::eval` symbolically to count its constraints
//!   (crates/stark/src/machine.rs:357-362); it returns one `SymbolicExpression<F>` per `assert_zero`, in emission order;
//! * `Chip::new` -> `LookupBuilder` (crates/stark/src/chip.rs:66-91, lookup/builder.rs), which leaves the chip's
//!   `sends` / `receives` as `Lookup { values: Vec<VirtualPairCol>, multiplicity, kind, scope }`.
//!
//! The output is the JSON `zkmips_b200/air/ir.py::Air.from_exported_json` loads (same node vocabulary as
//! `Air.to_json`; `zkmips_b200/air/exported/*.json` are examples produced by the hand transcriptions):
//!
//! ```json
//! {"name": "AddSub", "main_width": 19, "prep_width": 0, "num_public_values": 231, "commit_scope": "local",
//!  "local_only": true, "batch_size": 2, "permutation_constraints_included": false,
//!  "nodes": [["main", 0, 17], ["main", 0, 18], ["add", 0, 1], ["const", 1], ...],
//!  "constraints": [57, 61, ...],
//!  "sends":    [{"kind": 4, "values": [[c, [["main", col, weight], ...]], ...], "mult": [c, [...]]}, ...],
//!  "receives": [...]}
//! ```
//!
//! The LogUp constraints are NOT exported: the loader appends them itself with the transliteration of
//! `eval_permutation_constraints` (crates/stark/src/permutation.rs:205-347), from the lookups and `batch_size`, exactly
//! where `Chip::eval` appends them (chip.rs:259-270).  Constants and weights are canonical `u32`s.
use std::collections::HashMap;
use std::sync::Arc;

use p3_air::{Air, PairCol, VirtualPairCol};
use p3_field::PrimeField32;
use p3_uni_stark::{get_symbolic_constraints, Entry, SymbolicAirBuilder, SymbolicExpression};
use serde_json::{json, Value};
use zkm_stark::air::{LookupScope, MachineAir};
use zkm_stark::lookup::Lookup;
use zkm_stark::{Chip, PROOF_MAX_NUM_PVS};

/// Hash-consing arena: a `SymbolicExpression` tree shares sub-trees through `Arc`, so nodes are keyed by the pointer of
/// their `Arc` first (no re-walk of shared sub-trees) and by their structural key second (same id for equal nodes).
#[derive(Default)]
struct Arena {
    nodes: Vec<Value>,
    by_key: HashMap<String, usize>,
    by_ptr: HashMap<usize, usize>,
}

impl Arena {
    fn intern(&mut self, node: Value) -> usize {
        let key = node.to_string();
        if let Some(&id) = self.by_key.get(&key) {
            return id;
        }
        self.nodes.push(node);
        self.by_key.insert(key, self.nodes.len() - 1);
        self.nodes.len() - 1
    }

    fn arc<F: PrimeField32>(&mut self, e: &Arc<SymbolicExpression<F>>) -> usize {
        let ptr = Arc::as_ptr(e) as usize;
        if let Some(&id) = self.by_ptr.get(&ptr) {
            return id;
        }
        let id = self.expr(e.as_ref());
        self.by_ptr.insert(ptr, id);
        id
    }

    fn expr<F: PrimeField32>(&mut self, e: &SymbolicExpression<F>) -> usize {
        match e {
            SymbolicExpression::Variable(v) => match v.entry {
                // row offset 0 = local, 1 = next (folder.rs:52-149 reads the same two rows)
                Entry::Preprocessed { offset } => self.intern(json!(["prep", offset, v.index])),
                Entry::Main { offset } => self.intern(json!(["main", offset, v.index])),
                Entry::Public => self.intern(json!(["pv", v.index])),
                // the symbolic run sees only the chip's own constraints: permutation columns and challenges appear
                // in the LogUp constraints, which the loader generates
                Entry::Permutation { .. } | Entry::Challenge => unreachable!("chip constraints do not read these"),
            },
            SymbolicExpression::IsFirstRow => self.intern(json!(["first"])),
            SymbolicExpression::IsLastRow => self.intern(json!(["last"])),
            SymbolicExpression::IsTransition => self.intern(json!(["trans"])),
            SymbolicExpression::Constant(c) => self.intern(json!(["const", c.as_canonical_u32()])),
            SymbolicExpression::Add { x, y, .. } => {
                let (a, b) = (self.arc(x), self.arc(y));
                self.intern(json!(["add", a, b]))
            }
            SymbolicExpression::Sub { x, y, .. } => {
                let (a, b) = (self.arc(x), self.arc(y));
                self.intern(json!(["sub", a, b]))
            }
            SymbolicExpression::Mul { x, y, .. } => {
                let (a, b) = (self.arc(x), self.arc(y));
                self.intern(json!(["mul", a, b]))
            }
            SymbolicExpression::Neg { x, .. } => {
                let a = self.arc(x);
                self.intern(json!(["neg", a]))
            }
        }
    }
}

/// `VirtualPairCol` -> `[constant, [[table, column, weight], ...]]` (ir.py `linear_form`)
fn linear_form<F: PrimeField32>(col: &VirtualPairCol<F>) -> Value {
    let (weights, constant) = col.clone().into_parts(); // (Vec<(PairCol, F)>, F)
    let mut terms: Vec<(u8, usize, u32)> = weights
        .iter()
        .map(|(c, w)| match c {
            PairCol::Main(i) => (0u8, *i, w.as_canonical_u32()),
            PairCol::Preprocessed(i) => (1u8, *i, w.as_canonical_u32()),
        })
        .filter(|t| t.2 != 0)
        .collect();
    terms.sort(); // ir.py sorts (table, column): "main" < "prep"
    let terms: Vec<Value> =
        terms.iter().map(|(t, i, w)| json!([if *t == 0 { "main" } else { "prep" }, i, w])).collect();
    json!([constant.as_canonical_u32(), terms])
}

fn lookups<F: PrimeField32>(ls: &[Lookup<F>]) -> Vec<Value> {
    ls.iter()
        .filter(|l| l.scope == LookupScope::Local)
        .map(|l| {
            json!({"kind": l.argument_index(),
                   "values": l.values.iter().map(linear_form).collect::<Vec<_>>(),
                   "mult": linear_form(&l.multiplicity)})
        })
        .collect()
}

/// The JSON program of one chip.
pub fn export_chip<F, A>(chip: &Chip<F, A>) -> Value
where
    F: PrimeField32,
    A: MachineAir<F> + Air<SymbolicAirBuilder<F>>,
{
    let mut arena = Arena::default();
    let constraints: Vec<usize> =
        get_symbolic_constraints(chip.air(), chip.preprocessed_width(), PROOF_MAX_NUM_PVS)
            .iter()
            .map(|c| arena.expr(c))
            .collect();
    json!({
        "name": chip.name(),
        "main_width": chip.width(),
        "prep_width": chip.preprocessed_width(),
        "perm_width": 0,                      // filled in by the loader from the lookups
        "num_public_values": PROOF_MAX_NUM_PVS,
        "num_challenges": 2,
        "commit_scope": match chip.commit_scope() { LookupScope::Local => "local", LookupScope::Global => "global" },
        "local_only": chip.local_only(),
        "batch_size": chip.logup_batch_size(),
        "permutation_constraints_included": false,
        "nodes": arena.nodes,
        "constraints": constraints,
        "sends": lookups(chip.sends()),
        "receives": lookups(chip.receives()),
    })
}

/// Every chip of a machine (`StarkMachine::chips()`), one JSON file per chip under `dir`.
pub fn export_machine<F, A>(chips: &[Chip<F, A>], dir: &std::path::Path) -> std::io::Result<()>
where
    F: PrimeField32,
    A: MachineAir<F> + Air<SymbolicAirBuilder<F>>,
{
    std::fs::create_dir_all(dir)?;
    for chip in chips {
        std::fs::write(dir.join(format!("{}.json", chip.name())), export_chip(chip).to_string())?;
    }
    Ok(())
}

// Usage from the Ziren workspace (a test or a small bin):
//
//     let machine = MipsAir::<KoalaBear>::machine(KoalaBearPoseidon2::new());
//     zkm_air_export::export_machine(machine.chips(), Path::new("zkgpu_airs/core"))?;
//     let compress = RecursionAir::<KoalaBear, 3>::compress_machine(InnerSC::default());
//     zkm_air_export::export_machine(compress.chips(), Path::new("zkgpu_airs/compress"))?;
//
// then `python -m zkmips_b200.air.codegen --from-json zkgpu_airs/core` regenerates csrc/gen/ for those chips.
