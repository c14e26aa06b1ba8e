"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of oracle/_ref/libzkref.so.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (zkmips_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0x7F000001

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("zk_oracle.c", "zk_oracle_fri.c", "zk_oracle.h", "kb31.h")]
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-C", HERE, "-s", os.path.join(HERE, "liboracle.so")])
    if os.path.isdir("/root/reference/crates/recursion/core/include"):
        ref = os.path.join(HERE, "_ref", "libzkref.so")
        if force or not os.path.exists(ref):
            subprocess.check_call(["make", "-C", HERE, "-s", "ref"])
    return so


_lib = None
_ref = None


def _ptr(a):
    return a.ctypes.data_as(_u32p)


def lib():
    global _lib
    if _lib is None:
        so = build()
        L = C.CDLL(so)
        L.ork_to_monty.restype = C.c_uint32
        L.ork_to_monty.argtypes = [C.c_uint32]
        L.ork_from_monty.restype = C.c_uint32
        L.ork_from_monty.argtypes = [C.c_uint32]
        L.ork_mul.restype = C.c_uint32
        L.ork_mul.argtypes = [C.c_uint32, C.c_uint32]
        L.ork_inv.restype = C.c_uint32
        L.ork_inv.argtypes = [C.c_uint32]
        L.ork_two_adic_generator.restype = C.c_uint32
        L.ork_two_adic_generator.argtypes = [C.c_uint32]
        L.ork_ext_mul.argtypes = [_u32p, _u32p, _u32p]
        L.ork_ext_inv.argtypes = [_u32p, _u32p]
        L.ork_to_monty_vec.argtypes = [_u32p, _u32p, C.c_uint64]
        L.ork_from_monty_vec.argtypes = [_u32p, _u32p, C.c_uint64]
        L.ork_poseidon2_permute.argtypes = [_u32p]
        L.ork_poseidon2_permute_canonical.argtypes = [_u32p]
        L.ork_poseidon2_wide_trace.argtypes = [_u32p, C.c_uint64, C.c_uint64, C.c_int32, _u32p]
        L.ork_poseidon2_wide_prep.argtypes = [_u32p, C.c_uint64, C.c_uint64, _u32p]
        L.ork_add_sub_trace.argtypes = [_u32p, C.c_uint64, C.c_uint64, _u32p]
        L.ork_hash.argtypes = [_u32p, C.c_uint64, _u32p]
        L.ork_compress.argtypes = [_u32p, _u32p, _u32p]
        L.ork_hash_rows.argtypes = [_u32p, C.c_uint64, C.c_uint64, _u32p]
        L.ork_dft_batch.argtypes = [_u32p, C.c_uint64, C.c_uint64, _u32p]
        L.ork_coset_lde.argtypes = [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _u32p]
        L.ork_mmcs_commit.restype = C.c_int32
        L.ork_mmcs_commit.argtypes = [C.c_uint32, C.POINTER(_u32p), _u64p, _u64p, C.c_int32, _u32p,
                                      C.POINTER(C.c_void_p)]
        L.ork_pcs_commit.restype = C.c_int32
        L.ork_pcs_commit.argtypes = [C.c_uint32, C.POINTER(_u32p), _u64p, _u64p, _u32p, C.c_uint32, _u32p,
                                     C.POINTER(C.c_void_p)]
        L.ork_tree_free.argtypes = [C.c_void_p]
        L.ork_tree_num_matrices.restype = C.c_uint32
        L.ork_tree_num_matrices.argtypes = [C.c_void_p]
        L.ork_tree_height.restype = C.c_uint64
        L.ork_tree_height.argtypes = [C.c_void_p, C.c_uint32]
        L.ork_tree_width.restype = C.c_uint64
        L.ork_tree_width.argtypes = [C.c_void_p, C.c_uint32]
        L.ork_tree_matrix.restype = _u32p
        L.ork_tree_matrix.argtypes = [C.c_void_p, C.c_uint32]
        L.ork_tree_log_max_height.restype = C.c_uint32
        L.ork_tree_log_max_height.argtypes = [C.c_void_p]
        L.ork_tree_layer.restype = _u32p
        L.ork_tree_layer.argtypes = [C.c_void_p, C.c_uint32]
        L.ork_tree_open.argtypes = [C.c_void_p, C.c_uint64, _u32p, _u32p]
        L.ork_mmcs_verify.restype = C.c_int32
        L.ork_mmcs_verify.argtypes = [_u32p, C.c_uint32, _u64p, _u64p, C.c_uint64, _u32p, _u32p, C.c_uint32]
        L.ork_num_threads.restype = C.c_int32
        L.ork_set_num_threads.argtypes = [C.c_int32]
        _bind_fri(L)
        _lib = L
    return _lib


def _bind_fri(L):
    """Signatures of oracle/zk_oracle_fri.c (filled in as that file grows)."""
    if not hasattr(L, "ork_ch_init"):
        return
    from . import binding_fri
    binding_fri.bind(L)


def ref():
    """oracle/_ref/libzkref.so (the reference's own C++ field class + Poseidon2), or None."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libzkref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/crates/recursion/core/include"):
                build()
            if not os.path.exists(path):
                return None
        R = C.CDLL(path)
        for f in ("ref_add", "ref_sub", "ref_mul"):
            getattr(R, f).restype = C.c_uint32
            getattr(R, f).argtypes = [C.c_uint32, C.c_uint32]
        for f in ("ref_inv", "ref_to_monty", "ref_from_monty"):
            getattr(R, f).restype = C.c_uint32
            getattr(R, f).argtypes = [C.c_uint32]
        R.ref_poseidon2_permute.argtypes = [_u32p]
        if hasattr(R, "ref_poseidon2_wide_event_to_row"):
            R.ref_poseidon2_wide_event_to_row.argtypes = [_u32p, _u32p, C.c_int]
            R.ref_poseidon2_wide_instr_to_row.argtypes = [_u32p, _u32p]
            R.ref_add_sub_event_to_row.argtypes = [_u32p, _u32p]
        if hasattr(R, "ref_poseidon2_skinny_event_to_rows"):
            R.ref_poseidon2_skinny_event_to_rows.argtypes = [_u32p, _u32p]
            R.ref_poseidon2_skinny_instr_to_row.argtypes = [_u32p, C.c_uint64, _u32p]
        _ref = R
    return _ref


# ------------------------------------------------------------------ numpy helpers
def to_monty(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty_like(a)
    lib().ork_to_monty_vec(_ptr(a), _ptr(out), a.size)
    return out


def from_monty(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty_like(a)
    lib().ork_from_monty_vec(_ptr(a), _ptr(out), a.size)
    return out


def permute(state):
    s = np.ascontiguousarray(state, dtype=np.uint32).copy()
    assert s.shape == (16,)
    lib().ork_poseidon2_permute(_ptr(s))
    return s


def permute_canonical(state):
    s = np.ascontiguousarray(state, dtype=np.uint32).copy()
    lib().ork_poseidon2_permute_canonical(_ptr(s))
    return s


def poseidon2_wide_trace(inputs, rows, sbox=True):
    """Poseidon2WideChip<3|9> main trace (rows x 313 | 172) from [n, 16] Montgomery inputs; padding = zero-input row"""
    x = np.ascontiguousarray(inputs, dtype=np.uint32).reshape(-1, 16)
    out = np.empty((rows, 313 if sbox else 172), np.uint32)
    lib().ork_poseidon2_wide_trace(_ptr(x), len(x), rows, int(sbox), _ptr(out))
    return out


def poseidon2_wide_prep(instrs, rows):
    x = np.ascontiguousarray(instrs, dtype=np.uint32).reshape(-1, 48)
    out = np.empty((rows, 49), np.uint32)
    lib().ork_poseidon2_wide_prep(_ptr(x), len(x), rows, _ptr(out))
    return out


def add_sub_trace(events, rows):
    x = np.ascontiguousarray(events, dtype=np.uint32).reshape(-1, 7)
    out = np.empty((rows, 19), np.uint32)
    lib().ork_add_sub_trace(_ptr(x), len(x), rows, _ptr(out))
    return out


def hash_slice(x):
    x = np.ascontiguousarray(x, dtype=np.uint32)
    out = np.empty(8, dtype=np.uint32)
    lib().ork_hash(_ptr(x), x.size, _ptr(out))
    return out


def compress(l, r):
    l = np.ascontiguousarray(l, dtype=np.uint32)
    r = np.ascontiguousarray(r, dtype=np.uint32)
    out = np.empty(8, dtype=np.uint32)
    lib().ork_compress(_ptr(l), _ptr(r), _ptr(out))
    return out


def hash_rows(mat):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    h, w = mat.shape
    out = np.empty((h, 8), dtype=np.uint32)
    lib().ork_hash_rows(_ptr(mat), h, w, _ptr(out))
    return out


def dft_batch(mat):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    h, w = mat.shape
    out = np.empty_like(mat)
    lib().ork_dft_batch(_ptr(mat), h, w, _ptr(out))
    return out


def coset_lde(mat, log_blowup, shift_monty):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    h, w = mat.shape
    out = np.empty((h << log_blowup, w), dtype=np.uint32)
    lib().ork_coset_lde(_ptr(mat), h, w, log_blowup, int(shift_monty), _ptr(out))
    return out


class Tree:
    """Owning wrapper of an ork_tree (MerkleTreeMmcs prover data)."""

    def __init__(self, handle, root, keep=None):
        self.handle = handle
        self.root = root
        self._keep = keep

    def __del__(self):
        if getattr(self, "handle", None):
            lib().ork_tree_free(self.handle)
            self.handle = None

    @property
    def num_matrices(self):
        return lib().ork_tree_num_matrices(self.handle)

    @property
    def log_max_height(self):
        return lib().ork_tree_log_max_height(self.handle)

    def dims(self, i):
        return int(lib().ork_tree_height(self.handle, i)), int(lib().ork_tree_width(self.handle, i))

    def matrix(self, i):
        h, w = self.dims(i)
        p = lib().ork_tree_matrix(self.handle, i)
        return np.ctypeslib.as_array(p, shape=(h, w)).copy()

    def layer(self, l):
        n = 1 << (self.log_max_height - l)
        p = lib().ork_tree_layer(self.handle, l)
        return np.ctypeslib.as_array(p, shape=(n, 8)).copy()

    def open(self, index):
        n = self.num_matrices
        widths = [self.dims(i)[1] for i in range(n)]
        opened = np.empty(max(1, sum(widths)), dtype=np.uint32)
        proof = np.empty((self.log_max_height, 8), dtype=np.uint32)
        lib().ork_tree_open(self.handle, index, _ptr(opened), _ptr(proof))
        rows, off = [], 0
        for w in widths:
            rows.append(opened[off:off + w].copy())
            off += w
        return rows, proof


def _mat_args(mats):
    mats = [np.ascontiguousarray(m, dtype=np.uint32) for m in mats]
    n = len(mats)
    ptrs = (_u32p * n)(*[_ptr(m) for m in mats])
    heights = np.array([m.shape[0] for m in mats], dtype=np.uint64)
    widths = np.array([m.shape[1] for m in mats], dtype=np.uint64)
    return mats, n, ptrs, heights, widths


def mmcs_commit(mats):
    mats, n, ptrs, heights, widths = _mat_args(mats)
    root = np.empty(8, dtype=np.uint32)
    h = C.c_void_p()
    rc = lib().ork_mmcs_commit(n, ptrs, heights.ctypes.data_as(_u64p), widths.ctypes.data_as(_u64p), 1,
                               _ptr(root), C.byref(h))
    if rc != 0:
        raise ValueError(f"ork_mmcs_commit failed: {rc}")
    return Tree(h, root)


def mmcs_verify(root, dims, index, rows, proof):
    heights = np.array([d[0] for d in dims], dtype=np.uint64)
    widths = np.array([d[1] for d in dims], dtype=np.uint64)
    opened = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint32) for r in rows])
                                  if rows else np.zeros(1, np.uint32), dtype=np.uint32)
    if opened.size == 0:
        opened = np.zeros(1, np.uint32)
    proof = np.ascontiguousarray(proof, dtype=np.uint32).reshape(-1, 8)
    root = np.ascontiguousarray(root, dtype=np.uint32)
    return bool(lib().ork_mmcs_verify(_ptr(root), len(dims), heights.ctypes.data_as(_u64p),
                                      widths.ctypes.data_as(_u64p), index, _ptr(opened), _ptr(proof),
                                      proof.shape[0]))


def pcs_commit(mats, log_blowup=1, domain_shifts=None):
    mats, n, ptrs, heights, widths = _mat_args(mats)
    if domain_shifts is None:
        domain_shifts = [lib().ork_to_monty(1)] * n
    shifts = np.array(domain_shifts, dtype=np.uint32)
    root = np.empty(8, dtype=np.uint32)
    h = C.c_void_p()
    rc = lib().ork_pcs_commit(n, ptrs, heights.ctypes.data_as(_u64p), widths.ctypes.data_as(_u64p),
                              _ptr(shifts), log_blowup, _ptr(root), C.byref(h))
    if rc != 0:
        raise ValueError(f"ork_pcs_commit failed: {rc}")
    return Tree(h, root)


def use_all_cores():
    """OpenMP thread count = the cores this process may run on, whatever OMP_NUM_THREADS says (torch.distributed.run
    exports OMP_NUM_THREADS=1 to its workers).  Returns the count."""
    n = len(os.sched_getaffinity(0))
    lib().ork_set_num_threads(n)
    return lib().ork_num_threads()
