"""ctypes binding of libzkgpu.so (include/zkgpu.h).  No torch types cross this boundary.

`load()` opens the in-tree CUDA library and nothing else; it raises if the library is missing.  The
optional `path` argument exists for the test-only emulator build (tests/emu), never for production.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(HERE, "libzkgpu.so")

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p
i32, u32, u64 = C.c_int32, C.c_uint32, C.c_uint64

# name -> (restype, argtypes); mirrors include/zkgpu.h one to one (checked by tests/test_abi.py)
PROTOTYPES = {
    "zk_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "zk_ctx_create_on_stream": (i32, [i32, vp, C.POINTER(vp)]),
    "zk_ctx_destroy": (None, [vp]),
    "zk_ctx_sync": (i32, [vp]),
    "zk_last_error": (C.c_char_p, []),
    "zk_build_info": (C.c_char_p, []),
    "zk_prof_enable": (i32, [vp, i32]),
    "zk_prof_reset": (i32, [vp]),
    "zk_prof_count": (i32, [vp]),
    "zk_prof_get": (i32, [vp, i32, C.c_char_p, i32, C.POINTER(C.c_float), u64p]),
    "zk_prof_start": (i32, [vp, i32, C.POINTER(C.c_float)]),
    "zk_launch_count": (u64, [vp]),
    "zk_ntt_tma_passes": (u64, []),
    "zk_dev_alloc": (i32, [vp, u64, u64p]),
    "zk_dev_free": (i32, [vp, u64]),
    "zk_h2d": (i32, [vp, u64, vp, u64]),
    "zk_d2h": (i32, [vp, vp, u64, u64]),
    "zk_poseidon2_permute": (i32, [vp, u32p, u64]),
    "zk_hash_rows": (i32, [vp, u32p, u64, u32, u32p]),
    "zk_compress_layer": (i32, [vp, u32p, u64, u32p]),
    "zk_dft_batch": (i32, [vp, u32p, u64, u32, u32p]),
    "zk_coset_lde": (i32, [vp, u32p, u64, u32, u32, u32, u32p]),
    "zk_coset_lde_dev": (i32, [vp, u64, u64, u32, u32, u32, u64]),
    "zk_commit": (i32, [vp, u32, C.POINTER(vp), u64p, u32p, u32p, u32, u32p, C.POINTER(vp)]),
    "zk_commit_dev": (i32, [vp, u32, u64p, u64p, u32p, u32p, u32, u32p, C.POINTER(vp)]),
    "zk_mmcs_commit": (i32, [vp, u32, C.POINTER(vp), u64p, u32p, u32p, C.POINTER(vp)]),
    "zk_mmcs_commit_dev": (i32, [vp, u32, u64p, u64p, u32p, u32p, C.POINTER(vp)]),
    "zk_pdata_free": (None, [vp]),
    "zk_pdata_num_matrices": (u32, [vp]),
    "zk_pdata_height": (u64, [vp, u32]),
    "zk_pdata_width": (u32, [vp, u32]),
    "zk_pdata_log_max_height": (u32, [vp]),
    "zk_pdata_root": (i32, [vp, u32p]),
    "zk_pdata_lde": (u64, [vp, u32]),
    "zk_pdata_pitch": (u32, [vp, u32]),
    "zk_pdata_copy_lde": (i32, [vp, u32, u32p]),
    "zk_pdata_copy_layer": (i32, [vp, u32, u32p]),
    "zk_pdata_import": (i32, [vp, u32, C.POINTER(vp), u64p, u32p, C.POINTER(vp), u32, C.POINTER(vp), u32, C.POINTER(vp)]),
    "zk_pdata_open_batch": (i32, [vp, u32, u64p, u32p, u32p]),
    "zk_air_count": (i32, []),
    "zk_air_name": (C.c_char_p, [i32]),
    "zk_air_find": (i32, [C.c_char_p]),
    "zk_air_info": (i32, [i32, vp]),
    "zk_permutation_trace": (i32, [vp, i32, u64, u64, u64, u32p, u64p, u32p]),
    "zk_ctx_keep_traces": (i32, [vp, i32]),
    "zk_ctx_set_upload_helper": (i32, [vp, i32]),
    "zk_pdata_trace": (u64, [vp, u32]),
    "zk_quotient": (i32, [vp, i32, vp, u32, vp, u32, vp, u32, u32, u32, u32p, u32p, u32p, u32, u32p, u32p, u64p]),
    "zk_tracegen_alu_width": (u32, [i32]),
    "zk_tracegen_alu": (i32, [vp, i32, vp, u64, u64, u64p]),
    "zk_tracegen_alu_dev": (i32, [vp, i32, u64, u64, u64, u64p]),
    "zk_tracegen_cpu_width": (u32, []),
    "zk_tracegen_cpu": (i32, [vp, vp, u64, u64, u64p]),
    "zk_tracegen_cpu_dev": (i32, [vp, u64, u64, u64, u64p]),
    "zk_tracegen_poseidon2_wide_width": (u32, [i32]),
    "zk_tracegen_poseidon2_wide": (i32, [vp, u32p, u64, u64, i32, u64p]),
    "zk_tracegen_poseidon2_wide_dev": (i32, [vp, u64, u64, u64, i32, u64p]),
    "zk_tracegen_poseidon2_wide_prep": (i32, [vp, u32p, u64, u64, u64p]),
    "zk_tracegen_poseidon2_skinny_width": (u32, []),
    "zk_tracegen_poseidon2_skinny": (i32, [vp, u32p, u64, u64, u64p]),
    "zk_tracegen_poseidon2_skinny_dev": (i32, [vp, u64, u64, u64, u64p]),
    "zk_tracegen_poseidon2_skinny_prep": (i32, [vp, u32p, u64, u64, u64p]),
    "zk_challenger_init": (i32, [vp]),
    "zk_challenger_observe": (i32, [vp, vp, u32p, u32]),
    "zk_challenger_sample_ext": (i32, [vp, vp, u32, u32p]),
    "zk_challenger_sample_bits": (i32, [vp, vp, u32, u32, u64p]),
    "zk_challenger_grind": (i32, [vp, vp, u32, u32p]),
    "zk_pcs_proof_words": (u64, [u32, C.POINTER(vp), u32p, u32, u32]),
    "zk_pcs_open": (i32, [vp, u32, C.POINTER(vp), u32p, u32p, u32, u32, u32, vp, C.c_int64, u32p, u64]),
}


class ZkError(RuntimeError):
    pass


def _p32(a):
    return a.ctypes.data_as(u32p)


def _arr(x, dt):
    return np.ascontiguousarray(np.asarray(x, dtype=dt))


class Lib:
    """One loaded libzkgpu.so.  Methods are 1:1 with the C ABI, with numpy arrays for host buffers."""

    def __init__(self, path=None):
        path = path or os.environ.get("ZK_LIB") or DEFAULT_SO  # ZK_LIB: A/B builds of the same sources (tools only)
        if not os.path.exists(path):
            raise ZkError(f"{path} is missing: build it with `python -m zkmips_b200.build` "
                          "(there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            if name not in EXTRA_OPTIONAL or hasattr(self.dll, name):
                fn = getattr(self.dll, name)
                fn.restype = res
                fn.argtypes = args

    def check(self, rc):
        if rc != 0:
            raise ZkError(f"libzkgpu status {rc}: {self.dll.zk_last_error().decode()}")

    # ---- context
    def ctx_create(self, device=0, stream=None):
        h = vp()
        if stream is None:
            self.check(self.dll.zk_ctx_create(device, C.byref(h)))
        else:
            self.check(self.dll.zk_ctx_create_on_stream(device, vp(stream), C.byref(h)))
        return Ctx(self, h)


EXTRA_OPTIONAL = set()


class PData:
    """Mmcs::ProverData handle (LDE matrices + digest layers resident on the device)."""

    def __init__(self, ctx, h, root):
        self.ctx, self.h, self.root = ctx, h, root

    @property
    def d(self):
        return self.ctx.lib.dll

    def free(self):
        if self.h:
            self.d.zk_pdata_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def num_matrices(self):
        return self.d.zk_pdata_num_matrices(self.h)

    def height(self, i):
        return self.d.zk_pdata_height(self.h, i)

    def width(self, i):
        return self.d.zk_pdata_width(self.h, i)

    def log_max_height(self):
        return self.d.zk_pdata_log_max_height(self.h)

    def lde_ptr(self, i):
        return self.d.zk_pdata_lde(self.h, i)

    def pitch(self, i):
        """row stride of LDE matrix i on the device, in words (>= width: odd widths are padded to even)"""
        return self.d.zk_pdata_pitch(self.h, i)

    def trace_ptr(self, i):
        return self.d.zk_pdata_trace(self.h, i)

    def lde(self, i):
        out = np.empty((self.height(i), self.width(i)), np.uint32)
        self.ctx.lib.check(self.d.zk_pdata_copy_lde(self.h, i, _p32(out)))
        return out

    def layer(self, l):
        out = np.empty(((1 << self.log_max_height()) >> l, 8), np.uint32)
        self.ctx.lib.check(self.d.zk_pdata_copy_layer(self.h, l, _p32(out)))
        return out

    def export(self):
        """host image of the prover data (what `Serialize` writes): LDE matrices + every digest layer"""
        n = self.num_matrices()
        return {"ldes": [self.lde(i) for i in range(n)], "layers": [self.layer(l) for l in range(self.log_max_height() + 1)]}

    def open_batch(self, indices):
        idx = _arr(indices, np.uint64)
        n = len(idx)
        sw = sum(self.width(i) for i in range(self.num_matrices()))
        L = self.log_max_height()
        opened = np.empty((n, sw), np.uint32)
        proofs = np.empty((n, L, 8), np.uint32)
        self.ctx.lib.check(self.d.zk_pdata_open_batch(self.h, n, idx.ctypes.data_as(u64p), _p32(opened), _p32(proofs)))
        return opened, proofs


class Ctx:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    @property
    def d(self):
        return self.lib.dll

    def destroy(self):
        if self.h:
            self.d.zk_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        self.lib.check(self.d.zk_ctx_sync(self.h))

    def launch_count(self):
        return self.d.zk_launch_count(self.h)

    # ---- profiling
    def prof_enable(self, on=True):
        self.lib.check(self.d.zk_prof_enable(self.h, 1 if on else 0))

    def prof_reset(self):
        self.lib.check(self.d.zk_prof_reset(self.h))

    def prof_records(self):
        out = []
        buf = C.create_string_buffer(64)
        ms = C.c_float()
        nl = C.c_uint64()
        for i in range(self.d.zk_prof_count(self.h)):
            self.lib.check(self.d.zk_prof_get(self.h, i, buf, 64, C.byref(ms), C.byref(nl)))
            out.append((buf.value.decode(), ms.value, nl.value))
        return out

    def prof_timeline(self):
        """(name, start ms after the first record, duration ms) of every record: a timeline of the ctx stream"""
        out = []
        t0 = C.c_float()
        for i, (name, ms, _) in enumerate(self.prof_records()):
            self.lib.check(self.d.zk_prof_start(self.h, i, C.byref(t0)))
            out.append((name, t0.value, ms))
        return out

    # ---- device memory
    def dev_alloc(self, nbytes):
        p = u64()
        self.lib.check(self.d.zk_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        self.lib.check(self.d.zk_dev_free(self.h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.lib.check(self.d.zk_h2d(self.h, dptr, arr.ctypes.data_as(vp), arr.nbytes))

    def d2h(self, arr, dptr):
        self.lib.check(self.d.zk_d2h(self.h, arr.ctypes.data_as(vp), dptr, arr.nbytes))

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.dev_alloc(arr.nbytes)
        self.h2d(p, arr)
        return p

    # ---- unit entry points
    def poseidon2_permute(self, states):
        s = _arr(states, np.uint32).reshape(-1, 16).copy()
        self.lib.check(self.d.zk_poseidon2_permute(self.h, _p32(s), s.shape[0]))
        return s

    def hash_rows(self, mat):
        m = _arr(mat, np.uint32)
        h, w = m.shape
        out = np.empty((h, 8), np.uint32)
        self.lib.check(self.d.zk_hash_rows(self.h, _p32(m), h, w, _p32(out)))
        return out

    def compress_layer(self, prev):
        p = _arr(prev, np.uint32).reshape(-1, 8)
        n = p.shape[0] // 2
        out = np.empty((n, 8), np.uint32)
        self.lib.check(self.d.zk_compress_layer(self.h, _p32(p), n, _p32(out)))
        return out

    def dft_batch(self, mat):
        m = _arr(mat, np.uint32)
        h, w = m.shape
        out = np.empty((h, w), np.uint32)
        self.lib.check(self.d.zk_dft_batch(self.h, _p32(m), h, w, _p32(out)))
        return out

    def coset_lde(self, mat, log_blowup, shift):
        m = _arr(mat, np.uint32)
        h, w = m.shape
        out = np.empty((h << log_blowup, w), np.uint32)
        self.lib.check(self.d.zk_coset_lde(self.h, _p32(m), h, w, log_blowup, shift, _p32(out)))
        return out

    # ---- commit
    @staticmethod
    def _shape_args(mats_hw):
        heights = _arr([h for h, _ in mats_hw], np.uint64)
        widths = _arr([w for _, w in mats_hw], np.uint32)
        return heights, widths

    def commit(self, mats, domain_shifts, log_blowup=1):
        """TwoAdicFriPcs::commit on host matrices (list of (h, w) uint32 Montgomery arrays)."""
        ms = [_arr(m, np.uint32) for m in mats]
        heights, widths = self._shape_args([m.shape for m in ms])
        ptrs = (vp * len(ms))(*[m.ctypes.data_as(vp) for m in ms])
        shifts = _arr(domain_shifts, np.uint32)
        root = np.empty(8, np.uint32)
        pd = vp()
        self.lib.check(self.d.zk_commit(self.h, len(ms), ptrs, heights.ctypes.data_as(u64p), _p32(widths),
                                        _p32(shifts), log_blowup, _p32(root), C.byref(pd)))
        return root, PData(self, pd, root)

    def commit_dev(self, dptrs, shapes, domain_shifts, log_blowup=1):
        heights, widths = self._shape_args(shapes)
        ptrs = _arr(dptrs, np.uint64)
        shifts = _arr(domain_shifts, np.uint32)
        root = np.empty(8, np.uint32)
        pd = vp()
        self.lib.check(self.d.zk_commit_dev(self.h, len(shapes), ptrs.ctypes.data_as(u64p),
                                            heights.ctypes.data_as(u64p), _p32(widths), _p32(shifts),
                                            log_blowup, _p32(root), C.byref(pd)))
        return root, PData(self, pd, root)

    def quotient(self, air_name, main, log_degree, log_quotient_degree, alpha, prep=None, perm=None,
                 perm_challenges=None, public_values=(), local_cumsum=None, global_cumsum=None):
        """quotient_values for one chip.  main / prep / perm are (PData, matrix index) pairs.  Returns the device
        pointer of the chunk matrices (2^lqd matrices of 2^log_degree x 4)."""
        aid = self.d.zk_air_find(air_name.encode())
        if aid < 0:
            raise ZkError(f"unknown AIR {air_name}")
        pv = _arr(public_values, np.uint32)
        al = _arr(alpha, np.uint32)
        ch = _arr(perm_challenges, np.uint32).reshape(-1) if perm_challenges is not None else None
        lc = _arr(local_cumsum, np.uint32) if local_cumsum is not None else None
        gc = _arr(global_cumsum, np.uint32) if global_cumsum is not None else None
        out = u64()
        self.lib.check(self.d.zk_quotient(
            self.h, aid, prep[0].h if prep else None, prep[1] if prep else 0, main[0].h, main[1],
            perm[0].h if perm else None, perm[1] if perm else 0, log_degree, log_quotient_degree, _p32(al),
            _p32(ch) if ch is not None else None, _p32(pv) if pv.size else None, pv.size,
            _p32(lc) if lc is not None else None, _p32(gc) if gc is not None else None, C.byref(out)))
        return out.value

    def keep_traces(self, on=True):
        self.lib.check(self.d.zk_ctx_keep_traces(self.h, 1 if on else 0))

    def set_upload_helper(self, device):
        """An idle peer GPU whose PCIe link carries half of every trace slab (forwarded over NVLink); -1: off."""
        self.lib.check(self.d.zk_ctx_set_upload_helper(self.h, int(device)))

    def air_info(self, air_name):
        aid = self.d.zk_air_find(air_name.encode())
        if aid < 0:
            raise ZkError(f"unknown AIR {air_name}")
        buf = (u32 * 9)()
        self.lib.check(self.d.zk_air_info(aid, C.cast(buf, vp)))
        keys = ["main_width", "prep_width", "perm_width", "num_public_values", "num_challenges", "num_constraints",
                "max_degree", "num_kernels", "num_lookups"]
        return dict(zip(keys, list(buf)))

    def permutation_trace(self, air_name, prep_trace, main_trace, height, perm_challenges):
        """generate_permutation_trace on the device.  Returns (device pointer of the h x 4*perm_width matrix,
        local cumulative sum[4])."""
        aid = self.d.zk_air_find(air_name.encode())
        if aid < 0:
            raise ZkError(f"unknown AIR {air_name}")
        ch = _arr(perm_challenges, np.uint32).reshape(-1)
        out = u64()
        lcs = np.empty(4, np.uint32)
        self.lib.check(self.d.zk_permutation_trace(self.h, aid, prep_trace or 0, main_trace, height, _p32(ch),
                                                   C.byref(out), _p32(lcs)))
        return out.value, lcs

    # ---- device trace generation (csrc/tracegen.cuh)
    ALU_CHIPS = {"AddSub": 0, "Bitwise": 1, "Lt": 2, "ShiftLeft": 3, "ShiftRight": 4, "CloClz": 5}

    def tracegen_alu(self, chip, events, rows):
        """`AluEvent` records ([n, 7] uint32, host array or device pointer + count) -> device pointer of the padded
        rows x width main trace of AddSubChip / BitwiseChip / LtChip, and its width."""
        cid = self.ALU_CHIPS[chip]
        out = u64()
        if isinstance(events, tuple):
            dptr, n = events
            self.lib.check(self.d.zk_tracegen_alu_dev(self.h, cid, dptr, n, rows, C.byref(out)))
        else:
            ev = _arr(events, np.uint32).reshape(-1, 7)
            self.lib.check(self.d.zk_tracegen_alu(self.h, cid, ev.ctypes.data_as(vp), len(ev), rows, C.byref(out)))
        return out.value, self.d.zk_tracegen_alu_width(cid)

    def tracegen_poseidon2_wide(self, inputs, rows, sbox_state=True):
        """Poseidon2WideChip<3 | 9> main trace from the permutation inputs ([n, 16] Montgomery words, host array or
        (device pointer, n)); returns (device pointer, width)."""
        out = u64()
        if isinstance(inputs, tuple):
            dptr, n = inputs
            self.lib.check(self.d.zk_tracegen_poseidon2_wide_dev(self.h, dptr, n, rows, int(sbox_state), C.byref(out)))
        else:
            x = _arr(inputs, np.uint32).reshape(-1, 16)
            self.lib.check(self.d.zk_tracegen_poseidon2_wide(self.h, _p32(x), len(x), rows, int(sbox_state), C.byref(out)))
        return out.value, self.d.zk_tracegen_poseidon2_wide_width(int(sbox_state))

    def tracegen_poseidon2_wide_prep(self, instrs, rows):
        """Poseidon2WideChip preprocessed trace from [n, 48] instruction words; returns (device pointer, 49)."""
        x = _arr(instrs, np.uint32).reshape(-1, 48)
        out = u64()
        self.lib.check(self.d.zk_tracegen_poseidon2_wide_prep(self.h, _p32(x), len(x), rows, C.byref(out)))
        return out.value, 49

    def tracegen_cpu(self, events, rows):
        """CpuChip main trace from packed CPU events ([n, 22] uint32: zk_cpu_event, host array or (device pointer, n));
        returns (device pointer, 67)."""
        out = u64()
        if isinstance(events, tuple):
            dptr, n = events
            self.lib.check(self.d.zk_tracegen_cpu_dev(self.h, dptr, n, rows, C.byref(out)))
        else:
            ev = _arr(events, np.uint32).reshape(-1, 22)
            self.lib.check(self.d.zk_tracegen_cpu(self.h, ev.ctypes.data_as(vp), len(ev), rows, C.byref(out)))
        return out.value, self.d.zk_tracegen_cpu_width()

    def tracegen_poseidon2_skinny(self, inputs, rows):
        """Poseidon2SkinnyChip main trace (eleven 28-word rows per permutation) from the permutation inputs ([n, 16]
        Montgomery words, host array or (device pointer, n)); returns (device pointer, 28)."""
        out = u64()
        if isinstance(inputs, tuple):
            dptr, n = inputs
            self.lib.check(self.d.zk_tracegen_poseidon2_skinny_dev(self.h, dptr, n, rows, C.byref(out)))
        else:
            x = _arr(inputs, np.uint32).reshape(-1, 16)
            self.lib.check(self.d.zk_tracegen_poseidon2_skinny(self.h, _p32(x), len(x), rows, C.byref(out)))
        return out.value, self.d.zk_tracegen_poseidon2_skinny_width()

    def tracegen_poseidon2_skinny_prep(self, instrs, rows):
        """Poseidon2SkinnyChip preprocessed trace from [n, 48] instruction words; returns (device pointer, 51)."""
        x = _arr(instrs, np.uint32).reshape(-1, 48)
        out = u64()
        self.lib.check(self.d.zk_tracegen_poseidon2_skinny_prep(self.h, _p32(x), len(x), rows, C.byref(out)))
        return out.value, 51

    def download(self, dptr, shape):
        out = np.empty(shape, np.uint32)
        self.d2h(out, dptr)
        return out

    def import_pdata(self, image, traces=None, log_blowup=1):
        """`DeserializeOwned` for Mmcs::ProverData: rebuild device-resident prover data from PData.export()'s image
        (optionally with the retained input traces, e.g. `StarkProvingKey::traces`)."""
        ldes = [_arr(m, np.uint32) for m in image["ldes"]]
        layers = [_arr(l, np.uint32) for l in image["layers"]]
        heights, widths = self._shape_args([m.shape for m in ldes])
        lp = (vp * len(ldes))(*[m.ctypes.data_as(vp) for m in ldes])
        yp = (vp * len(layers))(*[l.ctypes.data_as(vp) for l in layers])
        tp = None
        keep = []
        if traces is not None:
            keep = [None if t is None else _arr(t, np.uint32) for t in traces]
            tp = (vp * len(ldes))(*[None if t is None else t.ctypes.data_as(vp) for t in keep])
        pd = vp()
        self.lib.check(self.d.zk_pdata_import(self.h, len(ldes), lp, heights.ctypes.data_as(u64p), _p32(widths), yp,
                                              len(layers), tp, log_blowup, C.byref(pd)))
        return PData(self, pd, layers[-1].reshape(-1)[:8].copy())

    def mmcs_commit(self, mats):
        ms = [_arr(m, np.uint32) for m in mats]
        heights, widths = self._shape_args([m.shape for m in ms])
        ptrs = (vp * len(ms))(*[m.ctypes.data_as(vp) for m in ms])
        root = np.empty(8, np.uint32)
        pd = vp()
        self.lib.check(self.d.zk_mmcs_commit(self.h, len(ms), ptrs, heights.ctypes.data_as(u64p), _p32(widths),
                                             _p32(root), C.byref(pd)))
        return root, PData(self, pd, root)


class Challenger:
    """Host image of a DuplexChallenger (34 words) whose operations run on the device."""

    def __init__(self, ctx, words=None):
        self.ctx = ctx
        self.w = np.zeros(34, np.uint32) if words is None else np.ascontiguousarray(words, dtype=np.uint32).copy()

    def _p(self):
        return self.w.ctypes.data_as(vp)

    def observe(self, vals):
        v = _arr(vals, np.uint32).reshape(-1)
        self.ctx.lib.check(self.ctx.d.zk_challenger_observe(self.ctx.h, self._p(), _p32(v), v.size))

    def observe_many(self, parts):
        """Consecutive observe / observe_slice calls in ONE device round trip: observing slices one after the other is
        observing their concatenation (DuplexChallenger::observe_slice, challenger.rs)."""
        parts = [np.asarray(p, np.uint32).reshape(-1) for p in parts]
        if parts:
            self.observe(np.concatenate(parts))

    def sample_ext(self, n=1):
        out = np.empty((n, 4), np.uint32)
        self.ctx.lib.check(self.ctx.d.zk_challenger_sample_ext(self.ctx.h, self._p(), n, _p32(out)))
        return out[0] if n == 1 else out

    def sample_bits(self, bits, n=1):
        out = np.empty(n, np.uint64)
        self.ctx.lib.check(self.ctx.d.zk_challenger_sample_bits(self.ctx.h, self._p(), bits, n, out.ctypes.data_as(u64p)))
        return out

    def grind(self, bits):
        w = u32()
        self.ctx.lib.check(self.ctx.d.zk_challenger_grind(self.ctx.h, self._p(), bits, C.byref(w)))
        return w.value


def pcs_open(ctx, rounds, points_per_mat, ch, log_blowup=1, num_queries=84, pow_bits=16, inject_witness=-1):
    """TwoAdicFriPcs::open.  rounds: list of PData; points_per_mat: flat list (round-major) of lists of 4-word
    points; ch: Challenger (advanced in place).  Returns the flat proof (layout: include/zkgpu.h)."""
    handles = (vp * len(rounds))(*[r.h for r in rounds])
    n_points = _arr([len(p) for p in points_per_mat], np.uint32)
    pts = _arr([q for p in points_per_mat for q in p], np.uint32).reshape(-1)
    if pts.size == 0:
        pts = np.zeros(4, np.uint32)
    words = ctx.d.zk_pcs_proof_words(len(rounds), handles, _p32(n_points), log_blowup, num_queries)
    proof = np.zeros(words, np.uint32)
    ctx.lib.check(ctx.d.zk_pcs_open(ctx.h, len(rounds), handles, _p32(n_points), _p32(pts), log_blowup, num_queries,
                                    pow_bits, ch._p(), inject_witness, _p32(proof), words))
    return proof


_default = None


def load(path=None):
    """Load the product library (default) or an explicit path (tests/emu only)."""
    global _default
    if path is not None:
        return Lib(path)
    if _default is None:
        _default = Lib()
    return _default
