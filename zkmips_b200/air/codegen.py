"""CUDA code generator: ir.Air -> straight-line sm_100a quotient kernels (csrc/gen/airs_gen.cuh).

Every AIR becomes one or more `__global__` kernels ("parts"): the constraint list is cut so that a part
stays below MAX_NODES live SSA nodes, which bounds code size and register pressure for wide chips; part p
adds  inv_zeroifier * sum_{k in part} alpha^(n-1-k) C_k  into the quotient buffer.  One thread evaluates one
point of the quotient domain (crates/stark/src/quotient.rs:57-170 evaluates PackedVal::WIDTH of them)."""
import os

from .ir import P

# minimum resident CTAs per SM the generated quotient kernels are compiled for (register cap 65536 / (128 * n)); env
# ZK_QUOT_MIN_CTAS for A/B builds.  Real chips' programs keep 150-255 values live (quot_Cpu_p0 255 registers with
# spills at 1), which leaves 12 % of the warp slots occupied and the kernels latency-bound (profiles/README.md).
QUOT_MIN_CTAS = int(os.environ.get("ZK_QUOT_MIN_CTAS", "1"))

MAX_NODES = 1500
R = 1 << 32


def _monty(v):
    return (v % P) * R % P


def _needed(air, cons):
    need = set()
    stack = list(cons)
    while stack:
        n = stack.pop()
        if n in need:
            continue
        need.add(n)
        node = air.nodes[n]
        if node[0] in ("add", "sub", "mul"):
            stack += [node[1], node[2]]
        elif node[0] == "neg":
            stack.append(node[1])
    return sorted(need)


def _parts(air):
    parts, cur = [], []
    for k, c in enumerate(air.constraints):
        trial = cur + [(k, c)]
        if cur and len(_needed(air, [x[1] for x in trial])) > MAX_NODES:
            parts.append(cur)
            cur = [(k, c)]
        else:
            cur = trial
    if cur:
        parts.append(cur)
    return parts


def _vec_width(width):
    """columns fetched by one load: rows of a committed LDE start 8-byte aligned (even pitch, zkgpu_internal.cuh
    `lde_pitch`), 16-byte aligned when the pitch is a multiple of 4 words.  One thread reads one ROW, so a warp's scalar
    column load touches 32 sectors for 32 words; fetching the aligned group of 4 (2) adjacent columns at once divides
    the load-store unit's work by the group size (wide_bitwise_4096 at 2^16 rows: 537 M scalar loads per quotient)."""
    pitch = (width + 1) & ~1
    return 4 if pitch % 4 == 0 else 2


class _Loads:
    """vector loads of one kernel part: the first use of a column fetches its whole aligned group, provided at least
    two columns of the group are used by the part"""

    def __init__(self, air, needed):
        self.air = air
        self.count = {}
        for n in needed:
            node = air.nodes[n]
            if node[0] in ("main", "prep"):
                key = self.key(node)
                self.count[key] = self.count.get(key, 0) + 1
        self.loaded = set()

    def key(self, node):
        vw = _vec_width(self.air.main_width if node[0] == "main" else self.air.prep_width)
        return (node[0], node[1], node[2] // vw, vw)

    def emit(self, n):
        node = self.air.nodes[n]
        key = self.key(node)
        kind, row, grp, vw = key
        src = f"R.{'m' if kind == 'main' else 'p'}{row}"
        if self.count[key] < 2:
            return [f"const uint32_t n{n} = __ldg({src} + {node[2]});"]
        g = f"g{'m' if kind == 'main' else 'p'}{row}_{grp}"
        lines = []
        if key not in self.loaded:
            self.loaded.add(key)
            ty = "uint4" if vw == 4 else "uint2"
            lines.append(f"const {ty} {g} = __ldg(reinterpret_cast<const {ty}*>({src} + {grp * vw}));")
        lines.append(f"const uint32_t n{n} = {g}.{'xyzw'[node[2] % vw]};")
        return lines


def _emit_node(air, n):
    node, ty = air.nodes[n], air.types[n]
    k = node[0]
    v = f"n{n}"
    if k == "const":
        return f"const uint32_t {v} = 0x{_monty(node[1]):08x}u;"
    if k == "main":
        return f"const uint32_t {v} = __ldg(R.m{node[1]} + {node[2]});"
    if k == "prep":
        return f"const uint32_t {v} = __ldg(R.p{node[1]} + {node[2]});"
    if k == "perm":
        return f"const kb::Ext {v} = quot::ld_ext(R.q{node[1]} + {4 * node[2]});"
    if k == "pv":
        return f"const uint32_t {v} = __ldg(A.pvs + {node[1]});"
    if k == "gcs":
        return f"const uint32_t {v} = __ldg(A.gcs + {node[1]});"
    if k == "lcs":
        return f"const kb::Ext {v} = quot::ld_ext(A.lcs);"
    if k == "chal":
        return f"const kb::Ext {v} = quot::ld_ext(A.chal + {4 * node[1]});"
    if k in ("first", "last", "trans"):
        return f"const uint32_t {v} = R.is_{k};"
    if k == "neg":
        a = f"n{node[1]}"
        return f"const uint32_t {v} = kb::neg({a});" if ty == "b" else f"const kb::Ext {v} = kb::ext_neg({a});"
    a, b = f"n{node[1]}", f"n{node[2]}"
    ta, tb = air.types[node[1]], air.types[node[2]]
    if ty == "b":
        return f"const uint32_t {v} = kb::{k}({a}, {b});"
    if ta == "e" and tb == "e":
        return f"const kb::Ext {v} = kb::ext_{k}({a}, {b});"
    if ta == "e":  # ext op base
        return f"const kb::Ext {v} = kb::ext_{k}_base({a}, {b});"
    # base op ext
    if k == "add":
        return f"const kb::Ext {v} = kb::ext_add_base({b}, {a});"
    if k == "mul":
        return f"const kb::Ext {v} = kb::ext_mul_base({b}, {a});"
    return f"const kb::Ext {v} = kb::ext_neg(kb::ext_sub_base({b}, {a}));"  # a - b


def _emit_linear(lf, name):
    """VirtualPairCol::apply on the local row (base field), Montgomery constants as immediates"""
    c, terms = lf
    lines = [f"    uint32_t {name} = 0x{_monty(c):08x}u;"]
    for (t, col, w) in terms:
        src = "m" if t == "main" else "p"
        if w % P == 1:
            lines.append(f"    {name} = kb::add({name}, __ldg({src} + {col}));")
        else:
            lines.append(f"    {name} = kb::add({name}, kb::mul(__ldg({src} + {col}), 0x{_monty(w):08x}u));")
    return lines


def _emit_logup(air):
    """generate_permutation_trace for one chip: one thread per row (crates/stark/src/permutation.rs:29-69)"""
    fn = f"logup_{air.name}"
    lookups = [(l, True) for l in air.sends] + [(l, False) for l in air.receives]
    width, bs = air.permutation_width, air.batch_size
    maxv = max(len(l["values"]) for l, _ in lookups)
    o = [f"__global__ void __launch_bounds__(128) {fn}(logup::Args A) {{",
         "  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;",
         "  if (r >= A.h) return;",
         "  const uint32_t* m = A.main + r * A.wm;",
         "  const uint32_t* p = A.prep + r * A.wp;",
         "  const kb::Ext alpha = logup::ld_ext(A.chal);",
         "  kb::Ext bp[%d];" % max(maxv, 1),
         "  bp[0] = logup::ld_ext(A.chal + 4);"]
    for k in range(1, maxv):
        o.append(f"  bp[{k}] = kb::ext_mul(bp[{k - 1}], bp[0]);")
    o.append("  kb::Ext total = kb::ext_zero();")
    for b in range(width - 1):
        chunk = lookups[b * bs:(b + 1) * bs]
        n = len(chunk)
        o.append("  {")
        o.append(f"    kb::Ext d[{n}];")
        o.append(f"    uint32_t mu[{n}];")
        for i, (l, is_send) in enumerate(chunk):
            o.append(f"    d[{i}] = kb::ext_add_base(alpha, 0x{_monty(l['kind']):08x}u);")
            for k, lf in enumerate(l["values"]):
                o += _emit_linear(lf, f"v{i}_{k}")
                o.append(f"    d[{i}] = kb::ext_add(d[{i}], kb::ext_mul_base(bp[{k}], v{i}_{k}));")
            o += _emit_linear(l["mult"], f"mm{i}")
            o.append(f"    mu[{i}] = {'mm%d' % i if is_send else 'kb::neg(mm%d)' % i};")
        o.append(f"    kb::Ext e = logup::batch_entry<{n}>(d, mu);")
        o.append(f"    logup::st_ext(A.perm + r * A.wq + {4 * b}, e);")
        o.append("    total = kb::ext_add(total, e);")
        o.append("  }")
    o.append("  logup::st_ext(A.rowsum + 4 * r, total);")
    o.append("}")
    o.append("")
    return fn, o


N_KERNEL_FILES = 8


def generate(airs):
    """Returns {relative file name: text}: N_KERNEL_FILES translation units with the quotient kernels (compiled
    in parallel; ptxas time grows faster than linearly with the size of one unit) and airs_gen.cuh with the
    declarations, the LogUp kernels and the registry."""
    files = [["// GENERATED by zkmips_b200/air/codegen.py -- do not edit.", '#include "../quotient.cuh"', "",
              "namespace quotgen {", ""] for _ in range(N_KERNEL_FILES)]
    hdr = ["// GENERATED by zkmips_b200/air/codegen.py -- do not edit.", "#pragma once", '#include "../quotient.cuh"',
           '#include "../logup.cuh"', "", "namespace quotgen {", ""]
    table = []
    nk = 0
    for ai, air in enumerate(airs):
        parts = _parts(air)
        names = []
        for pi, part in enumerate(parts):
            fn = f"quot_{air.name}_p{pi}"
            names.append(fn)
            hdr.append(f"__global__ void {fn}(quot::Args A);")
            out = files[nk % N_KERNEL_FILES]
            nk += 1
            out.append(f"__global__ void __launch_bounds__(128, {QUOT_MIN_CTAS}) {fn}(quot::Args A) {{")
            out.append("  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;")
            out.append("  if (i >= (1u << (A.log_n + A.lqd))) return;")
            out.append("  quot::Row R;")
            out.append("  quot::prologue(A, i, R);")
            out.append("  kb::Ext acc = kb::ext_zero();")
            # demand-driven emission: a node is materialised right before the first constraint that needs
            # it, which keeps live ranges (registers) short for wide chips
            done = set()
            loads = _Loads(air, _needed(air, [c for _, c in part]))
            for k, c in part:
                for n in _needed(air, [c]):
                    if n not in done:
                        done.add(n)
                        if air.nodes[n][0] in ("main", "prep"):
                            out += ["  " + line for line in loads.emit(n)]
                        else:
                            out.append("  " + _emit_node(air, n))
                f = "fold_b" if air.types[c] == "b" else "fold_e"
                out.append(f"  acc = quot::{f}(acc, A.alpha_pows + {4 * k}, n{c});")
            out.append(f"  quot::epilogue(A, i, R, acc, {'true' if pi == 0 else 'false'});")
            out.append("}")
            out.append("")
        lfn = "nullptr"
        if air.sends or air.receives:
            lfn, code = _emit_logup(air)
            hdr += code
        table.append((air, names, lfn))
    hdr.append("")
    hdr.append("struct Entry {")
    hdr.append("  const char* name;")
    hdr.append("  uint32_t main_w, prep_w, perm_w, n_pv, n_chal, n_constraints, max_degree, n_parts;")
    hdr.append("  uint32_t n_lookups;")
    hdr.append("  void (*logup)(logup::Args);")
    hdr.append("  void (*parts[%d])(quot::Args);" % max(1, max(len(n) for _, n, _ in table)))
    hdr.append("};")
    hdr.append("static const Entry AIRS[] = {")
    for air, names, lfn in table:
        hdr.append(f'  {{"{air.name}", {air.main_width}, {air.prep_width}, {air.perm_width}, {air.num_public_values}, '
                   f"{air.num_challenges}, {air.num_constraints}, {air.max_degree()}, {len(names)}, "
                   f"{len(air.sends) + len(air.receives)}, {lfn}, {{{', '.join(names)}}}}},")
    hdr.append("};")
    hdr.append(f"static const int NUM_AIRS = {len(table)};")
    hdr.append("")
    hdr.append("}  // namespace quotgen")
    out = {"airs_gen.cuh": "\n".join(hdr) + "\n"}
    for i, f in enumerate(files):
        f.append("}  // namespace quotgen")
        out[f"airs_kernels_{i}.cu"] = "\n".join(f) + "\n"
    return out


def exported_airs(json_dir):
    """the chips a Rust-side exporter (rust/air-export) wrote into `json_dir`, one <Chip>.json each"""
    from .ir import Air
    return [Air.from_exported_json(open(os.path.join(json_dir, f)).read())
            for f in sorted(os.listdir(json_dir)) if f.endswith(".json")]


def write(gen_dir=None, json_dir=None):
    """Writes csrc/gen/ (files are only rewritten when their text changes, so make-style staleness works).
    `json_dir` (env ZK_AIR_JSON_DIR): chips exported from the Ziren workspace are compiled in as well; an exported chip
    replaces the library's transcription of the same name."""
    from . import library
    here = os.path.dirname(os.path.abspath(__file__))
    gen_dir = gen_dir or os.path.join(os.path.dirname(here), "csrc", "gen")
    os.makedirs(gen_dir, exist_ok=True)
    airs = library.all_airs()
    json_dir = json_dir or os.environ.get("ZK_AIR_JSON_DIR")
    if json_dir:
        extra = exported_airs(json_dir)
        names = {a.name for a in extra}
        airs = [a for a in airs if a.name not in names] + extra
    for name, text in generate(airs).items():
        path = os.path.join(gen_dir, name)
        old = open(path).read() if os.path.exists(path) else None
        if old != text:
            with open(path, "w") as fh:
                fh.write(text)
    return gen_dir


if __name__ == "__main__":
    import sys
    print(write(json_dir=sys.argv[sys.argv.index("--from-json") + 1] if "--from-json" in sys.argv else None))
