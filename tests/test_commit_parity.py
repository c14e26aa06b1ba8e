"""Parity of the commit path (Poseidon2, sponge, compression, NTT, coset LDE, MMCS, Pcs::commit)
with the CPU oracle, bit for bit.  Every case runs twice: on the CPU emulator build (small sizes, no
GPU needed; debugging aid only) and -- marked `gpu` -- on the real library through the C ABI."""
import numpy as np
import pytest

from oracle import binding as ob
from tests import backends, util

P = util.P


def _backend(name):
    return backends.gpu() if name == "gpu" else backends.emu()


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _mont(h, w, kind="rand", seed=0):
    return ob.to_monty(util.canon_matrix(h, w, kind, seed=0x5A4B4D49 + seed))


@pytest.mark.parametrize("be", BACKENDS)
def test_permute_matches_oracle(be):
    ctx = _backend(be)
    rng = np.random.default_rng(11)
    s = rng.integers(0, P, (67, 16)).astype(np.uint32)
    s[0] = 0
    s[1] = P - 1
    got = ctx.poseidon2_permute(s)
    exp = np.stack([ob.permute(x) for x in s])
    assert (got == exp).all()


@pytest.mark.parametrize("be", BACKENDS)
@pytest.mark.parametrize("h,w", [(1, 8), (4, 1), (64, 7), (32, 8), (128, 24), (16, 100), (8, 0)])
def test_hash_rows(be, h, w):
    ctx = _backend(be)
    m = _mont(h, w)
    assert (ctx.hash_rows(m) == ob.hash_rows(m)).all()


@pytest.mark.parametrize("be", BACKENDS)
def test_compress_layer(be):
    ctx = _backend(be)
    d = _mont(64, 8)
    exp = np.stack([ob.compress(d[2 * i], d[2 * i + 1]) for i in range(32)])
    assert (ctx.compress_layer(d) == exp).all()


# log heights chosen to hit every pass kernel: register-only (1..5), shared-memory (6..10), and
# two-pass plans (11 = 1+10, 12 = 2+10)
EMU_LOGS = [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12]
GPU_LOGS = EMU_LOGS + [13, 14, 15, 16, 17, 20, 21]


def _dft_cases():
    out = []
    for n in GPU_LOGS:
        w = (3 if n % 2 else 4) if n >= 9 else 20
        marks = [] if n in EMU_LOGS else [pytest.mark.gpu]
        out.append(pytest.param("emu", n, w, id=f"emu-2^{n}x{w}", marks=marks + ([pytest.mark.skip("gpu only size")] if n not in EMU_LOGS else [])))
        out.append(pytest.param("gpu", n, w, id=f"gpu-2^{n}x{w}", marks=[pytest.mark.gpu]))
    return out


@pytest.mark.parametrize("be,n,w", _dft_cases())
def test_dft_batch(be, n, w):
    ctx = _backend(be)
    m = _mont(1 << n, w, seed=n)
    assert (ctx.dft_batch(m) == ob.dft_batch(m)).all()


@pytest.mark.parametrize("be,n,w", _dft_cases())
@pytest.mark.parametrize("log_blowup", [1, 2])
def test_coset_lde(be, n, w, log_blowup):
    if n + log_blowup > 22 and be == "emu":
        pytest.skip("too large for the emulator")
    ctx = _backend(be)
    m = _mont(1 << n, w, seed=100 + n)
    shift = ob.lib().ork_to_monty(3)
    assert (ctx.coset_lde(m, log_blowup, shift) == ob.coset_lde(m, log_blowup, shift)).all()


@pytest.mark.gpu
def test_upload_helper_commit(monkeypatch):
    """zk_ctx_set_upload_helper: half of the rows of every slab travel over an idle peer GPU's PCIe link and NVLink.
    Column slabs, uneven cuts, a whole-matrix odd-width upload and retained traces, against the oracle; needs 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs a second (idle) GPU")
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    ctx = native.load().ctx_create(0)
    one = ob.lib().ork_to_monty(1)

    def pin(m):
        return torch.from_numpy(m.view(np.int32)).pin_memory().numpy().view(np.uint32)

    try:
        ctx.set_upload_helper(1)
        ctx.keep_traces(True)
        mats = [pin(_mont(1 << 16, 300, seed=801)), pin(_mont(1 << 17, 47, seed=802)), pin(_mont(1 << 15, 256, seed=803))]
        root, pd = ctx.commit(mats, [one] * 3, 1)
        tree = ob.pcs_commit(mats, 1, [one] * 3)
        assert (root == tree.root).all()
        for i, m in enumerate(mats):
            assert (pd.lde(i) == tree.matrix(i)).all()
            assert (ctx.download(pd.trace_ptr(i), m.shape) == m).all()
        pd.free()
        ctx.set_upload_helper(-1)
        root2, pd2 = ctx.commit(mats, [one] * 3, 1)
        assert (root2 == root).all()
        pd2.free()
        with pytest.raises(Exception):
            ctx.set_upload_helper(0)  # the context's own GPU is not a helper
    finally:
        ctx.destroy()


def _ragged_cases():
    out = []
    for be, sizes in (("emu", [(8, 20), (9, 6), (10, 24), (9, 36), (12, 8), (10, 22)]),
                      ("gpu", [(8, 20), (9, 6), (13, 20), (14, 36), (15, 8), (16, 24), (17, 72), (16, 56), (18, 22), (13, 6), (19, 20)])):
        for n, w in sizes:
            out.append(pytest.param(be, n, w, id=f"{be}-2^{n}x{w}", marks=[pytest.mark.gpu] if be == "gpu" else []))
    return out


@pytest.mark.parametrize("be,n,w", _ragged_cases())
def test_ntt_ragged_widths(be, n, w):
    """widths that are not a multiple of the 16-column tile (the last column group has idle lanes; with ZK_NTT_SPLIT=1 it
    runs as its own launch on 8- / 4-column tiles, an experiment that is bit-exact and did not pay) -- passes of 8, 9 and
    10 stages, first / middle / last passes, both directions, blowup 1 and 2, against the oracle."""
    ctx = _backend(be)
    shift = ob.lib().ork_to_monty(3)
    m = _mont(1 << n, w, seed=1200 + n + w)
    assert (ctx.dft_batch(m) == ob.dft_batch(m)).all()
    for lb in (1, 2):
        assert (ctx.coset_lde(m, lb, shift) == ob.coset_lde(m, lb, shift)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n,w", [(17, 4), (18, 4), (18, 2), (19, 4), (20, 2)])
def test_ntt_narrow_tiles(n, w):
    """4-column tiles for matrices of <= 4 columns (tiles of 512 and 1024 rows: passes of 9 and 10 stages), first,
    middle and last passes, against the oracle."""
    ctx = _backend("gpu")
    shift = ob.lib().ork_to_monty(3)
    m = _mont(1 << n, w, seed=900 + n + w)
    assert (ctx.dft_batch(m) == ob.dft_batch(m)).all()
    assert (ctx.coset_lde(m, 1, shift) == ob.coset_lde(m, 1, shift)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 21])
def test_ntt_tma_path(n, monkeypatch):
    """The persistent TMA-fed pass (csrc/ntt_tma.cuh) for every tile shape k = 6..10, first / middle / last passes,
    both directions, a ragged last column group (20 = 16 + 4 columns) and a full-width one: forced on for small
    matrices, compared with the oracle and with the plain shared-memory kernels (ZK_NTT_TMA=0)."""
    ctx = _backend("gpu")
    shift = ob.lib().ork_to_monty(3)
    w = 20 if n < 20 else 8
    m = _mont(1 << n, w, seed=700 + n)
    monkeypatch.setenv("ZK_NTT_TMA_MIN_TILES", "1")
    monkeypatch.setenv("ZK_NTT_TMA", "1")
    t0 = ctx.lib.dll.zk_ntt_tma_passes()
    d_tma = ctx.dft_batch(m)
    lde_tma = ctx.coset_lde(m, 1 if n >= 20 else 2, shift)
    npass = 1 if n <= 11 else 2 if n <= 20 else 3  # passes of >= 6 stages (11 = 6 + a register pass)
    assert ctx.lib.dll.zk_ntt_tma_passes() - t0 == npass * (1 + 1 + (2 if n >= 20 else 4)), "the TMA path was not taken"
    monkeypatch.setenv("ZK_NTT_TMA", "0")
    d_plain = ctx.dft_batch(m)
    lde_plain = ctx.coset_lde(m, 1 if n >= 20 else 2, shift)
    assert (d_tma == d_plain).all() and (lde_tma == lde_plain).all()
    if n <= 17:
        assert (d_tma == ob.dft_batch(m)).all()
        assert (lde_tma == ob.coset_lde(m, 2, shift)).all()


@pytest.mark.parametrize("be", BACKENDS)
def test_coset_lde_other_shift_and_wide(be):
    ctx = _backend(be)
    L = ob.lib()
    m = _mont(1 << 7, 37, seed=5)
    # quotient-chunk style shift: GENERATOR / (GENERATOR * g_8) = g_8^-1
    shift = L.ork_inv(L.ork_two_adic_generator(8))
    assert (ctx.coset_lde(m, 1, shift) == ob.coset_lde(m, 1, shift)).all()
    assert (ctx.coset_lde(m, 0, shift) == ob.coset_lde(m, 0, shift)).all()


def _check_commit(ctx, mats, shifts, log_blowup):
    root, pd = ctx.commit(mats, shifts, log_blowup)
    tree = ob.pcs_commit(mats, log_blowup, shifts)
    assert (root == tree.root).all()
    assert pd.log_max_height() == tree.log_max_height
    for i in range(len(mats)):
        assert (pd.lde(i) == tree.matrix(i)).all()
    for l in range(tree.log_max_height + 1):
        assert (pd.layer(l) == tree.layer(l)).all()
    # open_batch + the oracle's transliterated verify_batch
    hmax = 1 << tree.log_max_height
    idx = sorted({0, hmax - 1, hmax // 3, 5 % hmax})
    opened, proofs = pd.open_batch(idx)
    dims = [tree.dims(i) for i in range(len(mats))]
    for k, index in enumerate(idx):
        rows_o, proof_o = tree.open(index)
        off = 0
        rows = []
        for (h, w) in dims:
            rows.append(opened[k, off:off + w])
            off += w
        for a, b in zip(rows, rows_o):
            assert (a == b).all()
        assert (proofs[k] == proof_o).all()
        assert ob.mmcs_verify(root, dims, index, rows, proofs[k])
    pd.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_commit_single(be):
    ctx = _backend(be)
    one = ob.lib().ork_to_monty(1)
    _check_commit(ctx, [_mont(64, 16)], [one], 1)
    _check_commit(ctx, [_mont(32, 5)], [one], 2)


@pytest.mark.parametrize("be", BACKENDS)
def test_commit_mixed_heights(be):
    """size_gaps-style batch (crates/recursion/circuit/src/fri.rs:580-624): several matrices per height
    class, gaps between classes, widths that are not multiples of 8, a zero-width matrix."""
    ctx = _backend(be)
    L = ob.lib()
    one = L.ork_to_monty(1)
    mats = [_mont(16, 3, seed=1), _mont(128, 9, seed=2), _mont(128, 8, seed=3), _mont(2, 12, seed=4),
            _mont(16, 0, seed=5), _mont(1, 4, seed=6), _mont(64, 1, seed=7)]
    shifts = [one] * len(mats)
    # one quotient-chunk-like shifted domain: shift = GENERATOR * g^c
    shifts[2] = L.ork_mul(L.ork_to_monty(3), L.ork_two_adic_generator(8))
    _check_commit(ctx, mats, shifts, 1)


@pytest.mark.parametrize("be", BACKENDS)
def test_mmcs_commit_no_lde(be):
    ctx = _backend(be)
    mats = [_mont(32, 8, seed=9), _mont(8, 4, seed=10)]
    root, pd = ctx.mmcs_commit(mats)
    tree = ob.mmcs_commit(mats)
    assert (root == tree.root).all()
    pd.free()


@pytest.mark.parametrize("be", BACKENDS)
def test_bad_arguments(be):
    from zkmips_b200 import ZkError
    ctx = _backend(be)
    one = ob.lib().ork_to_monty(1)
    with pytest.raises(ZkError):
        ctx.commit([np.zeros((3, 4), np.uint32)], [one], 1)  # height not a power of two
    with pytest.raises(ZkError):
        ctx.commit([np.zeros((4, 4), np.uint32)], [0], 1)  # zero shift


@pytest.mark.parametrize("be", BACKENDS)
def test_streaming_commit_multi_slab(be, monkeypatch):
    """zk_commit streams column slabs (H2D || LDE || resumable sponge).  Force many small slabs and compare
    with the oracle: solo tallest matrix (streamed leaves), ragged last slab, and a mixed batch."""
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_SLAB_COLS", "16")
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    monkeypatch.setenv("ZK_HASH_VEC_MIN_ROWS", "0")  # block-aligned slabs take the vector-load sponge kernel at any height
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    one = ob.lib().ork_to_monty(1)
    try:
        _check_commit(ctx, [_mont(64, 40, seed=31)], [one], 1)          # slabs 16,16,8, leaves streamed
        _check_commit(ctx, [_mont(32, 24, seed=32)], [one], 2)
        _check_commit(ctx, [_mont(64, 35, seed=33)], [one], 1)          # width not a multiple of 8: LDE streamed only
        _check_commit(ctx, [_mont(64, 48, seed=34), _mont(16, 20, seed=35), _mont(64, 3, seed=36)], [one] * 3, 1)
        # single-matrix height classes: leaves AND injected digests hashed while streaming
        _check_commit(ctx, [_mont(128, 32, seed=37), _mont(32, 24, seed=38), _mont(8, 40, seed=39)], [one] * 3, 1)
    finally:
        ctx.destroy()


@pytest.mark.parametrize("be", BACKENDS)
def test_streaming_commit_tapered_slabs(be, monkeypatch):
    """Default slab schedule of the streaming commit: the last full slab is cut in halves so that the tail
    after the last upload is short (256 columns -> slabs of 64, 64, 64, 32, 32 columns).  Pinned host memory
    on the GPU (pageable traces take the one-copy path), forced streaming for small matrices."""
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    monkeypatch.delenv("ZK_SLAB_COLS", raising=False)
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    one = ob.lib().ork_to_monty(1)

    def pin(m):
        if be != "gpu":
            return m
        import torch
        return torch.from_numpy(m.view(np.int32)).pin_memory().numpy().view(np.uint32)

    try:
        _check_commit(ctx, [pin(_mont(64, 256, seed=41))], [one], 1)      # 64, 64, 64, 32, 32
        _check_commit(ctx, [pin(_mont(32, 320, seed=42))], [one], 2)      # 80, 80, 80, 80 (80 is not cut: 40 % 8 != 0 is avoided)
        _check_commit(ctx, [pin(_mont(64, 512, seed=43)), pin(_mont(16, 24, seed=44))], [one] * 2, 1)  # 128 x 3, 64, 64
    finally:
        ctx.destroy()
    # Uneven cuts: the widest slab of the schedule can exceed the nominal slab width -- 68 and 90 columns
    # round to ONE slab of the whole width, 300 columns are cut 72, 72, 72, 84 (nominal 80), 119 is odd.  Each in a FRESH
    # context, so the slab buffers are sized by this matrix alone (a larger earlier matrix would hide an overflow).
    for k, w in enumerate((68, 90, 119, 300, 1000)):
        ctx = lib.ctx_create(0)
        try:
            _check_commit(ctx, [pin(_mont(64, w, seed=45 + k))], [one], 1)
            _check_commit(ctx, [pin(_mont(32, w, seed=145 + k)), pin(_mont(32, 3, seed=245 + k))], [one] * 2, 2)
        finally:
            ctx.destroy()


@pytest.mark.parametrize("be", BACKENDS)
def test_streaming_commit_multi_member_classes(be, monkeypatch):
    """Height classes of SEVERAL matrices are streamed too: the class sponge is resumed matrix after matrix at any
    alignment (a 2-column matrix in front shifts every later chunk boundary by 2), slabs of 16 columns, zero-width
    members first / in the middle / last, classes interleaved in input order."""
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_SLAB_COLS", "16")
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    one = ob.lib().ork_to_monty(1)
    try:
        # keccak-like: narrow + wide chip of the same height (BASELINE config 3 shape)
        _check_commit(ctx, [_mont(64, 2, seed=51), _mont(64, 72, seed=52)], [one] * 2, 1)
        # three members, widths 5 + 40 + 3 (partial block carried across two matrix boundaries), shorter classes between
        _check_commit(ctx, [_mont(64, 5, seed=53), _mont(16, 3, seed=54), _mont(64, 40, seed=55), _mont(16, 13, seed=56),
                            _mont(64, 3, seed=57), _mont(4, 8, seed=58)], [one] * 6, 1)
        # zero-width members: first, middle and LAST of their class (the last one must still finalise the digests)
        _check_commit(ctx, [_mont(32, 0, seed=59), _mont(32, 9, seed=60), _mont(32, 0, seed=61), _mont(8, 24, seed=62),
                            _mont(8, 0, seed=63)], [one] * 5, 2)
        # exactly block-aligned members use the vector kernel, the misaligned ones after them the general one
        _check_commit(ctx, [_mont(64, 16, seed=64), _mont(64, 8, seed=65), _mont(64, 6, seed=66), _mont(64, 32, seed=67)],
                      [one] * 4, 1)
    finally:
        ctx.destroy()


@pytest.mark.parametrize("be", BACKENDS)
def test_commit_execution_shard_shape(be, monkeypatch):
    """The chips of a maximal log-21 execution shard (bench.py EXEC21_SHAPE: heights from maximal_shapes.json, widths
    from mips_costs.json), scaled down by 2^5 rows on the GPU (2^13 on the emulator) so the oracle finishes in seconds:
    eleven matrices, a three-member
    tallest class with widths 47 + 119 + 115, in the prover's (-height, name) order, streamed from host memory."""
    import bench
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    monkeypatch.setenv("ZK_SLAB_COLS", "32")
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    one = ob.lib().ork_to_monty(1)
    order = sorted(bench.EXEC21_SHAPE.items(), key=lambda kv: (-kv[1][0], kv[0]))
    down = 13 if be == "emu" else 5
    mats = [_mont(1 << max(lg - down, 0), w, seed=70 + k) for k, (name, (lg, w)) in enumerate(order)]
    try:
        _check_commit(ctx, mats, [one] * len(mats), 1)
    finally:
        ctx.destroy()


@pytest.mark.parametrize("be", BACKENDS)
def test_streaming_commit_retained_traces(be, monkeypatch):
    """zk_ctx_keep_traces: the slabs are uploaded straight into the retained trace (strided) and transformed from
    there.  The retained trace must equal the input and the commitment the oracle's, for several slabs per matrix,
    a ragged last slab, an odd width and a one-slab matrix."""
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_SLAB_COLS", "16")
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    ctx.keep_traces(True)
    one = ob.lib().ork_to_monty(1)
    try:
        mats = [_mont(64, 40, seed=81), _mont(64, 35, seed=82), _mont(16, 8, seed=83), _mont(64, 2, seed=84)]
        root, pd = ctx.commit(mats, [one] * len(mats), 1)
        assert (root == ob.pcs_commit(mats, 1, [one] * len(mats)).root).all()
        for i, m in enumerate(mats):
            assert (ctx.download(pd.trace_ptr(i), m.shape) == m).all(), f"retained trace {i} differs from the input"
        pd.free()
    finally:
        ctx.destroy()


def test_lde_two_pass_k10_emu():
    """2^20 rows x 2 columns: both k=10 passes of the second-generation kernel (coset scale + bit-reversed
    gather fused in the first, pass twiddles) on the emulator; the GPU twins are the 2^20 cases above."""
    ctx = backends.emu()
    m = _mont(1 << 20, 2, seed=77)
    shift = ob.lib().ork_to_monty(3)
    assert (ctx.coset_lde(m, 1, shift) == ob.coset_lde(m, 1, shift)).all()


@pytest.mark.parametrize("be", BACKENDS)
def test_prover_data_export_import_round_trip(be):
    """`PcsProverData: Serialize + DeserializeOwned` (crates/stark/src/prover.rs:221, machine.rs:56-57): the host image
    of a mixed-height commitment (LDE matrices + digest layers) is imported into a fresh handle, which must answer
    open_batch exactly like the original (rows and Merkle paths) and carry the same root; with the retained traces
    restored, the handle also serves the LogUp stage (`pk.traces`).  The context is destroyed while the imported handle
    is still alive: destruction is deferred to the last zk_pdata_free."""
    from zkmips_b200 import native
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    ctx = lib.ctx_create(0)
    one = ob.lib().ork_to_monty(1)
    mats = [_mont(64, 12, seed=301), _mont(16, 5, seed=302), _mont(64, 0, seed=303), _mont(4, 9, seed=304)]
    root, pd = ctx.commit(mats, [one] * len(mats), 1)
    image = pd.export()
    idx = [0, 1, 77, 127]
    opened, proofs = pd.open_batch(idx)
    pd.free()
    pd2 = ctx.import_pdata(image, traces=mats, log_blowup=1)
    assert (pd2.root == root).all()
    assert pd2.num_matrices() == 4 and pd2.log_max_height() == 7
    o2, p2 = pd2.open_batch(idx)
    assert (o2 == opened).all() and (p2 == proofs).all()
    for i, m in enumerate(mats):
        assert (pd2.lde(i) == image["ldes"][i]).all()
        if m.size:
            assert (ctx.download(pd2.trace_ptr(i), m.shape) == m).all()
    tree = ob.pcs_commit(mats, 1, [one] * len(mats))
    dims = [tree.dims(i) for i in range(4)]
    rows = [o2[2, :12], o2[2, 12:17], o2[2, 17:17], o2[2, 17:26]]
    assert ob.mmcs_verify(root, dims, 77, rows, p2[2])
    ctx.destroy()      # deferred: pd2 is still alive
    assert pd2.num_matrices() == 4
    pd2.free()         # the last handle tears the context down


@pytest.mark.parametrize("be", BACKENDS)
def test_even_pitch_of_odd_width_ldes(be, monkeypatch):
    """Committed LDEs of odd width carry one padding column (rows start 8-byte aligned: the two-column NTT kernels and
    64-bit loads then apply to 47-, 115-, 119-column chips).  The padding is invisible: same root, same digest layers,
    same exported LDE and same openings as the dense layout (ZK_EVEN_PITCH=0) and as the oracle -- from host memory
    (streamed slabs, retained traces) and from device-resident traces."""
    from zkmips_b200 import native
    monkeypatch.setenv("ZK_STREAM_MIN_BYTES", "0")
    lib = native.load() if be == "gpu" else native.load(backends.build_emu())
    one = ob.lib().ork_to_monty(1)
    logs = (6, 6, 4) if be == "emu" else (12, 12, 9)
    mats = [_mont(1 << logs[0], 47, seed=401), _mont(1 << logs[1], 119, seed=402), _mont(1 << logs[2], 5, seed=403),
            _mont(1 << logs[2], 8, seed=404)]
    tree = ob.pcs_commit(mats, 1, [one] * len(mats))
    results = []
    for even in ("1", "0"):
        monkeypatch.setenv("ZK_EVEN_PITCH", even)
        ctx = lib.ctx_create(0)
        ctx.keep_traces(True)
        try:
            root, pd = ctx.commit(mats, [one] * len(mats), 1)
            assert [pd.pitch(i) for i in range(4)] == ([48, 120, 6, 8] if even == "1" else [47, 119, 5, 8])
            assert (root == tree.root).all()
            for i, m in enumerate(mats):
                assert (pd.lde(i) == tree.matrix(i)).all()
                assert (ctx.download(pd.trace_ptr(i), m.shape) == m).all()
            idx = [0, 3, (1 << (logs[0] + 1)) - 1]
            results.append(pd.open_batch(idx))
            pd.free()
            dptrs = [ctx.upload(m) for m in mats]
            root2, pd2 = ctx.commit_dev(dptrs, [m.shape for m in mats], [one] * len(mats), 1)
            assert (root2 == tree.root).all() and (pd2.lde(1) == tree.matrix(1)).all()
            pd2.free()
            for p in dptrs:
                ctx.dev_free(p)
        finally:
            ctx.destroy()
    assert (results[0][0] == results[1][0]).all() and (results[0][1] == results[1][1]).all()
