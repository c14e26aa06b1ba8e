"""SSA constraint programs and the builder that records them.

Node kinds (type 'b' = base field, 'e' = degree-4 extension):
  const v            b   canonical constant
  prep r c | main r c b  column c of the local (r=0) / next (r=1) row
  perm r c           e   extension column c of the permutation trace (4 base columns in memory)
  pv k | gcs k       b   public value k / coordinate k of the global cumulative sum (SepticDigest, 14 words)
  lcs                e   local cumulative sum
  chal k             e   permutation challenge k
  first|last|trans   b   selectors of the trace domain on the quotient coset
  add a b | sub a b | mul a b | neg a      (result is 'e' when any operand is 'e')

`AirBuilder` mirrors the methods Ziren's chips call on p3 `AirBuilder` / `PairBuilder` /
`PermutationAirBuilder` / `MultiTableAirBuilder` (crates/stark/src/folder.rs:52-149):
main(), preprocessed(), permutation(), permutation_randomness(), public_values(), local_cumulative_sum(),
global_cumulative_sum(), is_first_row(), is_last_row(), is_transition(), when(...), when_first_row(),
when_last_row(), when_transition(), assert_zero(), assert_eq(), assert_one(), assert_bool(),
assert_zero_ext(), assert_eq_ext().  Constraints keep their emission order: constraint k is folded with
alpha^(n-1-k) (crates/stark/src/prover.rs:453-456).
"""
import json

P = 0x7F000001


class Expr:
    __slots__ = ("air", "id")

    def __init__(self, air, nid):
        self.air, self.id = air, nid

    @property
    def ty(self):
        return self.air.types[self.id]

    def _lift(self, o):
        if isinstance(o, Expr):
            return o
        return self.air.const(int(o))

    def __add__(self, o):
        return self.air.op("add", self, self._lift(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self.air.op("sub", self, self._lift(o))

    def __rsub__(self, o):
        return self.air.op("sub", self._lift(o), self)

    def __mul__(self, o):
        return self.air.op("mul", self, self._lift(o))

    __rmul__ = __mul__

    def __neg__(self):
        return self.air.op("neg", self)


class Air:
    """One chip's constraint program."""

    def __init__(self, name, main_width, prep_width=0, perm_width=0, num_public_values=0, num_challenges=2,
                 commit_scope="local", local_only=False):
        self.name = name
        # MachineAir::commit_scope (LookupScope::Local / Global, crates/stark/src/air/machine.rs) and
        # MachineAir::local_only (no constraint reads the next row: opened at zeta only, prover.rs:503-531)
        assert commit_scope in ("local", "global")
        self.commit_scope, self.local_only = commit_scope, bool(local_only)
        self.main_width, self.prep_width, self.perm_width = main_width, prep_width, perm_width
        self.num_public_values, self.num_challenges = num_public_values, num_challenges
        self.nodes = []   # tuples
        self.types = []   # 'b' / 'e'
        self.constraints = []  # node ids, emission order
        self._memo = {}
        # LogUp lookups (crates/stark/src/lookup/lookup.rs:8-19), Local scope: dicts
        #   {"kind": argument_index, "values": [linear form...], "mult": linear form}
        # a linear form is (constant, [(table 'main'|'prep', column, weight), ...]) = p3 VirtualPairCol
        self.sends, self.receives = [], []
        self.batch_size = 2  # 2^log_quotient_degree (crates/stark/src/chip.rs:173-175)
        self.lookups_finalized = False

    # ---- node construction with hash-consing
    def _node(self, key, ty):
        nid = self._memo.get(key)
        if nid is None:
            nid = len(self.nodes)
            self.nodes.append(key)
            self.types.append(ty)
            self._memo[key] = nid
        return Expr(self, nid)

    def const(self, v):
        return self._node(("const", int(v) % P), "b")

    def leaf(self, kind, *args):
        ty = "e" if kind in ("perm", "lcs", "chal") else "b"
        return self._node((kind,) + tuple(int(a) for a in args), ty)

    def op(self, kind, a, b=None):
        if kind == "neg":
            return self._node(("neg", a.id), a.ty)
        ia, ib = a.id, b.id
        # constant folding / identities keep generated kernels small
        ka, kb_ = self.nodes[ia], self.nodes[ib]
        if ka[0] == "const" and kb_[0] == "const":
            va, vb = ka[1], kb_[1]
            return self.const({"add": va + vb, "sub": va - vb, "mul": va * vb}[kind])
        if kind == "add" and ka == ("const", 0):
            return b
        if kind in ("add", "sub") and kb_ == ("const", 0):
            return a
        if kind == "mul" and (ka == ("const", 1)):
            return b
        if kind == "mul" and (kb_ == ("const", 1)):
            return a
        if kind in ("add", "mul") and ia > ib and a.ty == b.ty:
            ia, ib = ib, ia  # commutative canonical order
        ty = "e" if "e" in (self.types[ia], self.types[ib]) else "b"
        return self._node((kind, ia, ib), ty)

    # ---- VirtualPairCol extraction: an Expr that is affine in the LOCAL row's columns
    def linear_form(self, e):
        def walk(nid):
            n = self.nodes[nid]
            k = n[0]
            if k == "const":
                return n[1], {}
            if k in ("main", "prep"):
                if n[1] != 0:
                    raise ValueError("lookup values may only use the local row")
                return 0, {(k, n[2]): 1}
            if k == "neg":
                c, t = walk(n[1])
                return (-c) % P, {kk: (-v) % P for kk, v in t.items()}
            if k in ("add", "sub"):
                ca, ta = walk(n[1])
                cb, tb = walk(n[2])
                sg = 1 if k == "add" else -1
                out = dict(ta)
                for kk, v in tb.items():
                    out[kk] = (out.get(kk, 0) + sg * v) % P
                return (ca + sg * cb) % P, out
            if k == "mul":
                ca, ta = walk(n[1])
                cb, tb = walk(n[2])
                if ta and tb:
                    raise ValueError("lookup values must be linear in the trace columns")
                if not ta:
                    return ca * cb % P, {kk: v * ca % P for kk, v in tb.items()}
                return ca * cb % P, {kk: v * cb % P for kk, v in ta.items()}
            raise ValueError(f"node kind {k} is not allowed in a lookup value")
        c, terms = walk(e.id)
        return (c % P, sorted((t, col, w) for (t, col), w in terms.items() if w % P))

    @property
    def permutation_width(self):
        """local_permutation_trace_width (crates/stark/src/permutation.rs:18-23), in extension columns"""
        nl = len(self.sends) + len(self.receives)
        return 0 if nl == 0 else -(-nl // self.batch_size) + 1

    # ---- serialisation (the format a Rust-side recording builder would emit)
    def to_json(self):
        return json.dumps({
            "name": self.name, "main_width": self.main_width, "prep_width": self.prep_width,
            "perm_width": self.perm_width, "num_public_values": self.num_public_values,
            "num_challenges": self.num_challenges, "nodes": [list(n) for n in self.nodes],
            "constraints": self.constraints, "sends": self.sends, "receives": self.receives,
            "batch_size": self.batch_size, "commit_scope": self.commit_scope, "local_only": self.local_only})

    @classmethod
    def from_json(cls, text):
        d = json.loads(text)
        a = cls(d["name"], d["main_width"], d["prep_width"], d["perm_width"], d["num_public_values"],
                d.get("num_challenges", 2), d.get("commit_scope", "local"), d.get("local_only", False))
        for n in d["nodes"]:
            n = tuple(n)
            if n[0] in ("add", "sub", "mul"):
                ty = "e" if "e" in (a.types[n[1]], a.types[n[2]]) else "b"
            elif n[0] == "neg":
                ty = a.types[n[1]]
            else:
                ty = "e" if n[0] in ("perm", "lcs", "chal") else "b"
            a.nodes.append(n)
            a.types.append(ty)
        a.constraints = list(d["constraints"])
        fix = lambda lf: (lf[0], [tuple(t) for t in lf[1]])
        for key in ("sends", "receives"):
            setattr(a, key, [{"kind": l["kind"], "values": [fix(v) for v in l["values"]], "mult": fix(l["mult"])}
                             for l in d.get(key, [])])
        a.batch_size = d.get("batch_size", 2)
        a.lookups_finalized = True
        return a

    @classmethod
    def from_exported_json(cls, text):
        """Load what the Rust-side exporter (rust/air-export) writes: the chip's own constraints as a node DAG plus its
        lookups, WITHOUT the LogUp constraints (`permutation_constraints_included: false`) -- those are appended here,
        after the chip's own, exactly as `Chip::eval` does (crates/stark/src/chip.rs:259-270).  A file that already
        includes them (`Air.to_json`) loads unchanged."""
        d = json.loads(text)
        if d.get("permutation_constraints_included", True):
            return cls.from_json(text)
        a = cls.from_json(text)
        a._memo = {n: i for i, n in enumerate(a.nodes)}
        a.lookups_finalized = False
        a.perm_width = 0
        AirBuilder(a).eval_permutation_constraints(batch_size=d.get("batch_size", 2))
        return a

    def uses_next_row(self):
        """does any constraint read the next row of the preprocessed / main trace?  (a `local_only` chip must not)"""
        need, stack = set(), list(self.constraints)
        while stack:
            n = stack.pop()
            if n in need:
                continue
            need.add(n)
            node = self.nodes[n]
            if node[0] in ("add", "sub", "mul"):
                stack += [node[1], node[2]]
            elif node[0] == "neg":
                stack.append(node[1])
        return any(self.nodes[n][0] in ("main", "prep") and self.nodes[n][1] == 1 for n in need)

    @property
    def num_constraints(self):
        return len(self.constraints)

    def max_degree(self):
        """Degree of the constraints in the trace polynomials as crates/stark/src/chip.rs:81-87 computes it for
        log_quotient_degree (p3 SymbolicExpression::degree_multiple: is_first_row / is_last_row count 1,
        is_transition 0 -- its factor X - g^-1 is paid for by the vanishing polynomial)."""
        deg = []
        for n in self.nodes:
            k = n[0]
            if k in ("prep", "main", "perm", "first", "last"):
                deg.append(1)
            elif k in ("add", "sub"):
                deg.append(max(deg[n[1]], deg[n[2]]))
            elif k == "mul":
                deg.append(deg[n[1]] + deg[n[2]])
            elif k == "neg":
                deg.append(deg[n[1]])
            else:
                deg.append(0)
        return max([deg[c] for c in self.constraints] or [0])


class _Rows:
    def __init__(self, air, kind, width):
        self.air, self.kind, self.width = air, kind, width

    def row_slice(self, r):
        return [self.air.leaf(self.kind, r, c) for c in range(self.width)]

    def local(self):
        return self.row_slice(0)

    def next(self):
        return self.row_slice(1)


class AirBuilder:
    def __init__(self, air, cond=None):
        self.air, self.cond = air, cond

    # --- inputs
    def main(self):
        return _Rows(self.air, "main", self.air.main_width)

    def preprocessed(self):
        return _Rows(self.air, "prep", self.air.prep_width)

    def permutation(self):
        return _Rows(self.air, "perm", self.air.perm_width)

    def permutation_randomness(self):
        return [self.air.leaf("chal", k) for k in range(self.air.num_challenges)]

    def public_values(self):
        return [self.air.leaf("pv", k) for k in range(self.air.num_public_values)]

    def local_cumulative_sum(self):
        return self.air.leaf("lcs")

    def global_cumulative_sum(self):
        return [self.air.leaf("gcs", k) for k in range(14)]

    def is_first_row(self):
        return self.air.leaf("first")

    def is_last_row(self):
        return self.air.leaf("last")

    def is_transition(self):
        return self.air.leaf("trans")

    def const(self, v):
        return self.air.const(v)

    # --- filtered builders
    def when(self, cond):
        c = cond if self.cond is None else self.cond * cond
        return AirBuilder(self.air, c)

    def when_not(self, cond):
        """p3 AirBuilder::when_not: filter by (1 - cond)"""
        return self.when(1 - self._lift(cond))

    def when_first_row(self):
        return self.when(self.is_first_row())

    def when_last_row(self):
        return self.when(self.is_last_row())

    def when_transition(self):
        return self.when(self.is_transition())

    # --- constraints
    def _lift(self, x):
        return x if isinstance(x, Expr) else self.air.const(int(x))

    def assert_zero(self, x):
        x = self._lift(x)
        if self.cond is not None:
            x = self.cond * x
        self.air.constraints.append(x.id)

    def assert_eq(self, a, b):
        self.assert_zero(self._lift(a) - self._lift(b))

    def assert_one(self, x):
        self.assert_zero(self._lift(x) - 1)

    def assert_bool(self, x):
        x = self._lift(x)
        self.assert_zero(x * (x - 1))

    assert_zero_ext = assert_zero
    assert_eq_ext = assert_eq

    # --- LogUp lookups (ZKMAirBuilder::send / receive -> Lookup, crates/stark/src/lookup/builder.rs)
    def send(self, kind, values, multiplicity):
        self.air.sends.append({"kind": int(kind), "values": [self.air.linear_form(self._lift(v)) for v in values],
                               "mult": self.air.linear_form(self._lift(multiplicity))})

    def receive(self, kind, values, multiplicity):
        self.air.receives.append({"kind": int(kind), "values": [self.air.linear_form(self._lift(v)) for v in values],
                                  "mult": self.air.linear_form(self._lift(multiplicity))})

    def _apply(self, lf):
        """VirtualPairCol::apply on the local row"""
        c, terms = lf
        acc = self.air.const(c)
        for (t, col, w) in terms:
            acc = acc + self.air.leaf(t, 0, col) * w
        return acc

    def eval_permutation_constraints(self, batch_size=2):
        """Transliteration of eval_permutation_constraints (crates/stark/src/permutation.rs:205-347), appended
        after the chip's own constraints exactly as Chip::eval does (crates/stark/src/chip.rs:259-270) -- for EVERY
        chip: without local lookups and with Local scope it adds nothing (count_permutation_constraints,
        permutation.rs:355-388)."""
        air = self.air
        assert not air.lookups_finalized
        air.lookups_finalized = True
        air.batch_size = batch_size
        lookups = [(l, True) for l in air.sends] + [(l, False) for l in air.receives]
        if lookups:
            self._eval_local_lookups(lookups, batch_size)
        # Global scope (permutation.rs:333-346): the last row's final 14 main columns are the chip's contribution to
        # the shard's global cumulative sum (a SepticDigest: x[7], y[7]); 14 constraints, x_i and y_i interleaved
        if air.commit_scope == "global":
            main_local = self.main().local()
            gcs = self.global_cumulative_sum()
            w = air.main_width
            for i in range(7):
                self.when_last_row().assert_eq(main_local[w - 14 + i], gcs[i])
                self.when_last_row().assert_eq(main_local[w - 7 + i], gcs[7 + i])
        if air.local_only:
            assert not air.uses_next_row(), f"{air.name} is declared local_only but reads the next row"

    def _eval_local_lookups(self, lookups, batch_size):
        air = self.air
        air.perm_width = air.permutation_width
        perm, permn = self.permutation().local(), self.permutation().next()
        alpha, beta = self.permutation_randomness()[:2]
        lcs = self.local_cumulative_sum()
        width = air.perm_width
        for b in range(width - 1):
            chunk = lookups[b * batch_size:(b + 1) * batch_size]
            rlcs, mults = [], []
            for l, is_send in chunk:
                rlc = alpha + l["kind"]           # betas.next() == 1 multiplies the argument index
                bp = beta
                for lf in l["values"]:
                    rlc = rlc + bp * self._apply(lf)
                    bp = bp * beta
                rlcs.append(rlc)
                m = self._apply(l["mult"])
                mults.append(m if is_send else -m)
            product, numerator = None, None
            for i, (m, rlc) in enumerate(zip(mults, rlcs)):
                product = rlc if product is None else product * rlc
                others = None
                for j, o in enumerate(rlcs):
                    if j != i:
                        others = o if others is None else others * o
                term = m if others is None else others * m
                numerator = term if numerator is None else numerator + term
            self.assert_eq_ext(product * perm[b], numerator)
        sum_local, sum_next = perm[0], permn[0]
        for b in range(1, width - 1):
            sum_local, sum_next = sum_local + perm[b], sum_next + permn[b]
        phi_local, phi_next = perm[width - 1], permn[width - 1]
        self.when_first_row().assert_eq_ext(phi_local, sum_local)
        self.when_transition().assert_eq_ext(phi_next - phi_local, sum_next)
        self.when_last_row().assert_eq_ext(phi_local, lcs)
