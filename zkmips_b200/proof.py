"""`ShardProof` as the reference defines it, the repackaging of the library's flat PCS proof into it, and the
wire format the reference (de)serialises.

Reference:
  * structs and field order: crates/stark/src/types.rs:37-83 (`ShardCommitment`, `AirOpenedValues`,
    `ChipOpenedValues`, `ShardOpenedValues`, `ShardProof`); FRI proof crates/recursion/circuit/src/witness/stark.rs:
    60-142 and the aliases crates/stark/src/kb31_poseidon2.rs:37-44 (`FriProof{commit_phase_commits, query_proofs,
    final_poly, pow_witness}`, `QueryProof{input_proof, commit_phase_openings}`, `BatchOpening{opened_values,
    opening_proof}`, `CommitPhaseProofStep{sibling_value, opening_proof}`);
  * repackaging of `Pcs::open`'s output per chip: crates/stark/src/prover.rs:558-652;
  * serialisation: `#[derive(Serialize, Deserialize)]` + bincode 1.x default options (little endian, fixed-width
    integers, `usize`/lengths as u64, fixed arrays without a length prefix, `PhantomData` as nothing); a `KoalaBear`
    is written as its CANONICAL u32 (Plonky3 `MontyField31::serialize` = `serialize_u32(as_canonical_u32())`,
    [P3-recalled]), so the Montgomery words this library computes in are converted at this boundary.

All arrays held by these classes are uint32 MONTGOMERY words (the in-memory `KoalaBear`), exactly what the C ABI
returns.  `chip_ordering` is a `HashMap<String, usize>`: its serialised entry order is the map's iteration order
(unspecified in the reference); `to_bincode` writes the entries by index and `from_bincode` accepts any order."""
import struct
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

P = 0x7F000001
R_INV = pow(1 << 32, -1, P)
R_MOD = (1 << 32) % P


def from_monty(a):
    """Montgomery words -> canonical u32 (numpy, exact in uint64: both factors are < 2^31)."""
    return ((np.asarray(a, dtype=np.uint64) * np.uint64(R_INV)) % np.uint64(P)).astype(np.uint32)


def to_monty(a):
    return ((np.asarray(a, dtype=np.uint64) * np.uint64(R_MOD)) % np.uint64(P)).astype(np.uint32)


def _ext(n=0):
    return np.zeros((n, 4), np.uint32)


@dataclass
class AirOpenedValues:
    """types.rs:43-49: `local` / `next` rows of extension elements, (width, 4) words each"""
    local: np.ndarray = field(default_factory=_ext)
    next: np.ndarray = field(default_factory=_ext)


@dataclass
class ChipOpenedValues:
    """types.rs:51-63"""
    preprocessed: AirOpenedValues
    main: AirOpenedValues
    permutation: AirOpenedValues
    quotient: List[np.ndarray]              # per chunk: (4, 4) = the chunk's 4 base columns opened at zeta
    global_cumulative_sum: np.ndarray       # SepticDigest: x[7] then y[7]
    local_cumulative_sum: np.ndarray        # 4 words
    log_degree: int


@dataclass
class BatchOpening:
    opened_values: List[np.ndarray]         # one row per matrix of the batch
    opening_proof: np.ndarray               # (log_max_height, 8) siblings, bottom-up


@dataclass
class CommitPhaseProofStep:
    sibling_value: np.ndarray               # 4 words
    opening_proof: np.ndarray               # (depth, 8)


@dataclass
class QueryProof:
    input_proof: List[BatchOpening]         # one per round
    commit_phase_openings: List[CommitPhaseProofStep]


@dataclass
class FriProof:
    commit_phase_commits: np.ndarray        # (n_layers, 8)
    query_proofs: List[QueryProof]
    final_poly: np.ndarray                  # 4 words (one extension element: the constant)
    pow_witness: int                        # Montgomery word


@dataclass
class ShardCommitment:
    main_commit: np.ndarray
    permutation_commit: np.ndarray
    quotient_commit: np.ndarray


@dataclass
class ShardProof:
    """types.rs:77-83"""
    commitment: ShardCommitment
    opened_values: List[ChipOpenedValues]   # ShardOpenedValues.chips, in chip_ordering order
    opening_proof: FriProof
    chip_ordering: Dict[str, int]
    public_values: np.ndarray

    def local_cumulative_sum(self):
        """types.rs:98-100 (Montgomery words; addition is the same in both forms)"""
        acc = np.zeros(4, np.uint64)
        for c in self.opened_values:
            acc = (acc + c.local_cumulative_sum) % np.uint64(P)
        return acc.astype(np.uint32)


# ------------------------------------------------------------------------------------------------------------------
# flat proof (include/zkgpu.h "flat proof layout")  <->  structured proof
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class RoundShape:
    """What `Pcs::open` was called with for one round: committed (LDE) heights, widths and the number of opening
    points of every matrix of the batch, in commit order."""
    heights: List[int]
    widths: List[int]
    n_points: List[int]

    @property
    def log_max(self):
        return max(int(h).bit_length() - 1 for h in self.heights)


class _QueryView:
    """Sequence of QueryProof records backed by the flat proof buffer (built when indexed)."""

    def __init__(self, flat, off0, stride, n, rounds, log_max, n_layers):
        self.flat, self.off0, self.stride, self.n = flat, off0, stride, n
        self.rounds, self.log_max, self.n_layers = rounds, log_max, n_layers

    def __len__(self):
        return self.n

    def __iter__(self):
        return (self[k] for k in range(self.n))

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(self.n))]
        if k < 0:
            k += self.n
        if not 0 <= k < self.n:
            raise IndexError(k)
        flat, off = self.flat, self.off0 + k * self.stride
        inp = []
        for r in self.rounds:
            rows = []
            for w in r.widths:
                rows.append(flat[off:off + w])
                off += w
            path = flat[off:off + 8 * r.log_max].reshape(r.log_max, 8)
            off += 8 * r.log_max
            inp.append(BatchOpening(rows, path))
        steps = []
        for i in range(self.n_layers):
            sib = flat[off:off + 4]
            off += 4
            d = self.log_max - i - 1
            steps.append(CommitPhaseProofStep(sib, flat[off:off + 8 * d].reshape(d, 8)))
            off += 8 * d
        return QueryProof(inp, steps)


def split_flat_proof(flat, rounds: List[RoundShape], log_blowup, num_queries):
    """-> (opened[round][matrix][point] = (width, 4) array, FriProof)"""
    flat = np.asarray(flat, np.uint32)  # the structured proof holds VIEWS into this buffer
    off = 0
    opened = []
    for r in rounds:
        rnd = []
        for w, npts in zip(r.widths, r.n_points):
            mat = []
            for _ in range(npts):
                mat.append(flat[off:off + 4 * w].reshape(w, 4))
                off += 4 * w
            rnd.append(mat)
        opened.append(rnd)
    log_max = max(r.log_max for r in rounds)
    n_layers = max(log_max - log_blowup, 0)
    commits = flat[off:off + 8 * n_layers].reshape(n_layers, 8)
    off += 8 * n_layers
    final_poly = flat[off:off + 4]
    off += 4
    pow_witness = int(flat[off])
    off += 1
    # query records are fixed-size: sliced on demand (84 queries x (rounds x matrices + layers) views cost 2 ms when
    # built eagerly -- more than the device spends on the query phase)
    per_query = sum(sum(r.widths) + 8 * r.log_max for r in rounds) + sum(4 + 8 * (log_max - i - 1) for i in range(n_layers))
    assert off + num_queries * per_query == flat.size, \
        f"flat proof has {flat.size} words, layout accounts for {off + num_queries * per_query}"
    queries = _QueryView(flat, off, per_query, num_queries, rounds, log_max, n_layers)
    return opened, FriProof(commits, queries, final_poly, pow_witness)


def join_flat_proof(opened, fri: FriProof):
    """inverse of split_flat_proof"""
    parts = [np.asarray(v, np.uint32).reshape(-1) for rnd in opened for mat in rnd for v in mat]
    parts += [fri.commit_phase_commits.reshape(-1), np.asarray(fri.final_poly, np.uint32),
              np.array([fri.pow_witness], np.uint32)]
    for q in fri.query_proofs:
        for bo in q.input_proof:
            parts += [np.asarray(r, np.uint32) for r in bo.opened_values] + [bo.opening_proof.reshape(-1)]
        for st in q.commit_phase_openings:
            parts += [np.asarray(st.sibling_value, np.uint32), st.opening_proof.reshape(-1)]
    return np.concatenate([p.astype(np.uint32) for p in parts]) if parts else np.zeros(0, np.uint32)


# ------------------------------------------------------------------------------------------------------------------
# bincode
# ------------------------------------------------------------------------------------------------------------------
class _W:
    def __init__(self):
        self.b = bytearray()

    def u64(self, v):
        self.b += struct.pack("<Q", int(v))

    def felts(self, a):
        """field elements, canonical u32 LE, no length prefix (fixed arrays / single values)"""
        self.b += from_monty(np.asarray(a, np.uint32).reshape(-1)).astype("<u4").tobytes()

    def vec_felts(self, a, per=1):
        a = np.asarray(a, np.uint32)
        self.u64(a.size // per)
        self.felts(a)


class _R:
    def __init__(self, data):
        self.d, self.o = memoryview(bytes(data)), 0

    def u64(self):
        v = struct.unpack_from("<Q", self.d, self.o)[0]
        self.o += 8
        return v

    def felts(self, n, shape=None):
        a = np.frombuffer(self.d, dtype="<u4", count=n, offset=self.o).astype(np.uint32)
        self.o += 4 * n
        if a.size and int(a.max()) >= P:
            raise ValueError("non-canonical field element in a serialised proof")
        a = to_monty(a)
        return a.reshape(shape) if shape is not None else a

    def vec_felts(self, per=1):
        n = self.u64()
        return self.felts(n * per, (n, per) if per > 1 else None)


def _w_air(w: _W, v: AirOpenedValues):
    w.vec_felts(v.local, 4)
    w.vec_felts(v.next, 4)


def _r_air(r: _R):
    return AirOpenedValues(r.vec_felts(4).reshape(-1, 4), r.vec_felts(4).reshape(-1, 4))


def to_bincode(sp: ShardProof) -> bytes:
    """bincode::serialize(&ShardProof) (types.rs:77-83)"""
    w = _W()
    w.felts(sp.commitment.main_commit)
    w.felts(sp.commitment.permutation_commit)
    w.felts(sp.commitment.quotient_commit)
    w.u64(len(sp.opened_values))
    for c in sp.opened_values:
        _w_air(w, c.preprocessed)
        _w_air(w, c.main)
        _w_air(w, c.permutation)
        w.u64(len(c.quotient))
        for q in c.quotient:
            w.vec_felts(q, 4)
        w.felts(c.global_cumulative_sum)
        w.felts(c.local_cumulative_sum)
        w.u64(c.log_degree)
    f = sp.opening_proof
    w.vec_felts(f.commit_phase_commits, 8)
    w.u64(len(f.query_proofs))
    for q in f.query_proofs:
        w.u64(len(q.input_proof))
        for bo in q.input_proof:
            w.u64(len(bo.opened_values))
            for row in bo.opened_values:
                w.vec_felts(row)
            w.vec_felts(bo.opening_proof, 8)
        w.u64(len(q.commit_phase_openings))
        for st in q.commit_phase_openings:
            w.felts(st.sibling_value)
            w.vec_felts(st.opening_proof, 8)
    w.felts(f.final_poly)
    w.felts([f.pow_witness])
    w.u64(len(sp.chip_ordering))
    for name, idx in sorted(sp.chip_ordering.items(), key=lambda kv: kv[1]):
        raw = name.encode()
        w.u64(len(raw))
        w.b += raw
        w.u64(idx)
    w.vec_felts(sp.public_values)
    return bytes(w.b)


def from_bincode(data) -> ShardProof:
    r = _R(data)
    com = ShardCommitment(r.felts(8), r.felts(8), r.felts(8))
    chips = []
    for _ in range(r.u64()):
        prep, main, perm = _r_air(r), _r_air(r), _r_air(r)
        quotient = [r.vec_felts(4).reshape(-1, 4) for _ in range(r.u64())]
        gcs, lcs = r.felts(14), r.felts(4)
        chips.append(ChipOpenedValues(prep, main, perm, quotient, gcs, lcs, r.u64()))
    commits = r.vec_felts(8).reshape(-1, 8)
    queries = []
    for _ in range(r.u64()):
        inp = []
        for _ in range(r.u64()):
            rows = [r.vec_felts() for _ in range(r.u64())]
            inp.append(BatchOpening(rows, r.vec_felts(8).reshape(-1, 8)))
        steps = []
        for _ in range(r.u64()):
            sib = r.felts(4)
            steps.append(CommitPhaseProofStep(sib, r.vec_felts(8).reshape(-1, 8)))
        queries.append(QueryProof(inp, steps))
    final_poly = r.felts(4)
    pow_witness = int(r.felts(1)[0])
    order = {}
    for _ in range(r.u64()):
        n = r.u64()
        name = bytes(r.d[r.o:r.o + n]).decode()
        r.o += n
        order[name] = r.u64()
    pvs = r.vec_felts()
    if r.o != len(r.d):
        raise ValueError(f"{len(r.d) - r.o} trailing bytes after the ShardProof")
    return ShardProof(com, chips, FriProof(commits, queries, final_poly, pow_witness), order, pvs)
