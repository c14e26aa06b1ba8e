"""ctypes signatures + numpy wrappers for oracle/zk_oracle_fri.c (TEST INFRASTRUCTURE)."""
import ctypes as C

import numpy as np

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class Challenger(C.Structure):
    _fields_ = [("state", C.c_uint32 * 16), ("inp", C.c_uint32 * 8), ("n_in", C.c_uint32),
                ("out", C.c_uint32 * 8), ("n_out", C.c_uint32)]

    def words(self):
        """flat 34-word image (state, in, n_in, out, n_out) -- the zk_challenger layout of include/zkgpu.h"""
        return np.frombuffer(bytes(self), dtype=np.uint32).copy()

    @classmethod
    def from_words(cls, w):
        return cls.from_buffer_copy(np.ascontiguousarray(w, dtype=np.uint32).tobytes())


def bind(L):
    cp = C.POINTER(Challenger)
    L.ork_ch_init.argtypes = [cp]
    L.ork_ch_observe.argtypes = [cp, _u32p, C.c_uint32]
    L.ork_ch_sample.restype = C.c_uint32
    L.ork_ch_sample.argtypes = [cp]
    L.ork_ch_sample_ext.argtypes = [cp, _u32p]
    L.ork_ch_sample_bits.restype = C.c_uint32
    L.ork_ch_sample_bits.argtypes = [cp, C.c_uint32]
    L.ork_ch_check_witness.restype = C.c_int32
    L.ork_ch_check_witness.argtypes = [cp, C.c_uint32, C.c_uint32]
    L.ork_ch_grind.restype = C.c_uint32
    L.ork_ch_grind.argtypes = [cp, C.c_uint32]
    L.ork_pcs_proof_words.restype = C.c_uint64
    L.ork_pcs_proof_words.argtypes = [C.c_uint32, _u32p, _u64p, _u32p, _u32p, C.c_uint32, C.c_uint32]
    L.ork_pcs_open.restype = C.c_int32
    L.ork_pcs_open.argtypes = [C.c_uint32, C.POINTER(C.c_void_p), _u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32,
                               cp, C.c_int64, _u32p, C.c_uint64]
    L.ork_pcs_verify.restype = C.c_int32
    L.ork_pcs_verify.argtypes = [C.c_uint32, _u32p, _u32p, _u64p, _u32p, _u32p, _u32p, C.c_uint32, C.c_uint32,
                                 C.c_uint32, cp, _u32p, C.c_uint64]


def _p(a):
    return a.ctypes.data_as(_u32p)


def new_challenger(observe=None):
    from . import binding as ob
    ch = Challenger()
    ob.lib().ork_ch_init(C.byref(ch))
    if observe is not None:
        v = np.ascontiguousarray(observe, dtype=np.uint32)
        ob.lib().ork_ch_observe(C.byref(ch), _p(v), v.size)
    return ch


def observe(ch, vals):
    from . import binding as ob
    v = np.ascontiguousarray(vals, dtype=np.uint32).reshape(-1)
    ob.lib().ork_ch_observe(C.byref(ch), _p(v), v.size)


def sample_ext(ch):
    from . import binding as ob
    out = np.empty(4, np.uint32)
    ob.lib().ork_ch_sample_ext(C.byref(ch), _p(out))
    return out


def shapes_of(trees):
    n_mats = np.array([t.num_matrices for t in trees], np.uint32)
    hs, ws = [], []
    for t in trees:
        for i in range(t.num_matrices):
            h, w = t.dims(i)
            hs.append(h)
            ws.append(w)
    return n_mats, np.array(hs, np.uint64), np.array(ws, np.uint32)


def pcs_open(trees, points_per_mat, ch, log_blowup=1, num_queries=84, pow_bits=16, inject_witness=-1):
    """trees: list of binding.Tree (one per round); points_per_mat: flat list (round-major) of lists of
    4-word extension points.  Returns the flat proof (see zk_oracle.h) and mutates ch."""
    from . import binding as ob
    n_mats, hs, ws = shapes_of(trees)
    n_points = np.array([len(p) for p in points_per_mat], np.uint32)
    pts = np.ascontiguousarray(np.array([q for p in points_per_mat for q in p], np.uint32).reshape(-1))
    if pts.size == 0:
        pts = np.zeros(4, np.uint32)
    words = ob.lib().ork_pcs_proof_words(len(trees), _p(n_mats), hs.ctypes.data_as(_u64p), _p(ws), _p(n_points),
                                         log_blowup, num_queries)
    proof = np.zeros(words, np.uint32)
    handles = (C.c_void_p * len(trees))(*[t.handle for t in trees])
    rc = ob.lib().ork_pcs_open(len(trees), handles, _p(n_points), _p(pts), log_blowup, num_queries, pow_bits,
                               C.byref(ch), inject_witness, _p(proof), words)
    if rc != 0:
        raise ValueError(f"ork_pcs_open failed: {rc}")
    return proof


def pcs_verify(roots, n_mats, lde_heights, widths, points_per_mat, ch, proof, log_blowup=1, num_queries=84,
               pow_bits=16):
    from . import binding as ob
    roots = np.ascontiguousarray(np.array(roots, np.uint32).reshape(-1))
    n_mats = np.ascontiguousarray(n_mats, dtype=np.uint32)
    hs = np.ascontiguousarray(lde_heights, dtype=np.uint64)
    ws = np.ascontiguousarray(widths, dtype=np.uint32)
    n_points = np.array([len(p) for p in points_per_mat], np.uint32)
    pts = np.ascontiguousarray(np.array([q for p in points_per_mat for q in p], np.uint32).reshape(-1))
    if pts.size == 0:
        pts = np.zeros(4, np.uint32)
    proof = np.ascontiguousarray(proof, dtype=np.uint32)
    return ob.lib().ork_pcs_verify(len(n_mats), _p(roots), _p(n_mats), hs.ctypes.data_as(_u64p), _p(ws), _p(n_points),
                                   _p(pts), log_blowup, num_queries, pow_bits, C.byref(ch), _p(proof), proof.size)
