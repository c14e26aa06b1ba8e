"""Parity of the device challenger and of TwoAdicFriPcs::open (opening reduction, FRI commit phase, grind,
queries) with the CPU oracle: the flat proof buffers must be identical word for word, the transcripts must
end in the same state, and the oracle's transliteration of the reference verifier
(crates/recursion/circuit/src/fri.rs:34-405) must accept the proof the library produced."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from oracle import binding_fri as bf
from tests import backends, util
from zkmips_b200 import Challenger, ZkError, pcs_open

P = util.P
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _backend(name):
    return backends.gpu() if name == "gpu" else backends.emu()


def _mont(h, w, seed=0):
    return ob.to_monty(util.canon_matrix(h, w, "rand", seed=0x5A4B4D49 + seed))


@pytest.mark.parametrize("be", BACKENDS)
def test_device_challenger_matches_oracle(be):
    ctx = _backend(be)
    L = ob.lib()
    och = bf.new_challenger()
    dch = Challenger(ctx)
    rng = np.random.default_rng(5)
    for n in [3, 8, 1, 13, 5]:
        vals = ob.to_monty(rng.integers(0, P, n).astype(np.uint32))
        bf.observe(och, vals)
        dch.observe(vals)
        assert (dch.w == och.words()).all()
        assert (dch.sample_ext() == bf.sample_ext(och)).all()
        assert int(dch.sample_bits(11)[0]) == L.ork_ch_sample_bits(C.byref(och), 11)
        assert (dch.w == och.words()).all()
    w_dev = dch.grind(9)
    w_orc = L.ork_ch_grind(C.byref(och), 9)
    assert w_dev == w_orc
    assert (dch.w == och.words()).all()


def _setup(ctx, mats_per_round, log_blowup, n_points_fn):
    L = ob.lib()
    one = L.ork_to_monty(1)
    trees, pds = [], []
    for ms in mats_per_round:
        trees.append(ob.pcs_commit(ms, log_blowup))
        root, pd = ctx.commit(ms, [one] * len(ms), log_blowup)
        assert (root == trees[-1].root).all()
        pds.append(pd)
    och = bf.new_challenger()
    for t in trees:
        bf.observe(och, t.root)
    zeta = bf.sample_ext(och)
    pts = []
    for ms in mats_per_round:
        for m in ms:
            g = L.ork_two_adic_generator(int(np.log2(m.shape[0])))
            zg = np.array([L.ork_mul(int(z), g) for z in zeta], np.uint32)
            pts.append([zeta, zg][:n_points_fn(m)])
    return trees, pds, och, pts


def _run(ctx, mats_per_round, log_blowup, n_points_fn, nq=5, pow_bits=5):
    trees, pds, och, pts = _setup(ctx, mats_per_round, log_blowup, n_points_fn)
    ch_v = bf.Challenger.from_words(och.words())
    dch = Challenger(ctx, och.words())
    proof_o = bf.pcs_open(trees, pts, och, log_blowup, nq, pow_bits)
    proof_d = pcs_open(ctx, pds, pts, dch, log_blowup, nq, pow_bits)
    assert proof_d.size == proof_o.size
    bad = np.nonzero(proof_d != proof_o)[0]
    assert bad.size == 0, f"first differing word {bad[:5]} of {proof_o.size}"
    assert (dch.w == och.words()).all()
    n_mats, hs, ws = bf.shapes_of(trees)
    assert bf.pcs_verify([t.root for t in trees], n_mats, hs, ws, pts, ch_v, proof_d, log_blowup, nq, pow_bits) == 1
    for pd in pds:
        pd.free()
    return proof_d


@pytest.mark.parametrize("be", BACKENDS)
def test_open_single_matrix(be):
    _run(_backend(be), [[_mont(64, 5)]], 1, lambda m: 2)


@pytest.mark.parametrize("be", BACKENDS)
def test_open_shard_like_rounds(be):
    """four rounds (preprocessed / main / permutation / quotient chunks), mixed heights, one or two points
    per matrix as in crates/stark/src/prover.rs:503-544, a height-1 and a zero-width matrix."""
    mats = [
        [_mont(32, 3, 1), _mont(8, 2, 2)],
        [_mont(32, 11, 3), _mont(16, 40, 4), _mont(8, 6, 5), _mont(1, 3, 6)],
        [_mont(32, 8, 7), _mont(16, 4, 8), _mont(8, 0, 9)],
        [_mont(32, 4, 10), _mont(32, 4, 11), _mont(16, 4, 12)],
    ]
    _run(_backend(be), mats, 1, lambda m: 1 if m.shape[1] == 4 else 2)


@pytest.mark.parametrize("be", BACKENDS)
def test_open_blowup4(be):
    _run(_backend(be), [[_mont(16, 3, 1)], [_mont(64, 2, 2), _mont(4, 5, 3)]], 2, lambda m: 2, nq=3, pow_bits=3)


@pytest.mark.parametrize("be", BACKENDS)
def test_open_injected_witness(be):
    """A transcript with a given pow_witness (the reference's grind is non-deterministic) is reproduced when
    that witness is injected; a wrong witness is rejected."""
    ctx = _backend(be)
    mats = [[_mont(32, 4, 21)]]
    trees, pds, och, pts = _setup(ctx, mats, 1, lambda m: 2)
    och2 = bf.Challenger.from_words(och.words())
    proof_o = bf.pcs_open(trees, pts, och, 1, 4, 6)
    n_layers = 5
    wit = int(proof_o[2 * 4 * 4 + n_layers * 8 + 4])
    dch = Challenger(ctx, och2.words())
    proof_d = pcs_open(ctx, pds, pts, dch, 1, 4, 6, inject_witness=wit)
    assert (proof_d == proof_o).all()
    dch = Challenger(ctx, och2.words())
    with pytest.raises(ZkError):
        pcs_open(ctx, pds, pts, dch, 1, 4, 6, inject_witness=(wit + 1) % P if (wit + 1) % P != wit else 5)


@pytest.mark.gpu
def test_open_large_default_params():
    """reference parameters (84 queries, 16 PoW bits: crates/stark/src/kb31_poseidon2.rs:203-213) on a
    2^14 x 64 main trace plus smaller tables."""
    ctx = backends.gpu()
    mats = [[_mont(1 << 12, 7, 1)], [_mont(1 << 14, 64, 2), _mont(1 << 10, 19, 3)], [_mont(1 << 14, 8, 4)]]
    _run(ctx, mats, 1, lambda m: 2, nq=84, pow_bits=16)


@pytest.mark.gpu
def test_reference_pcs_inner_inputs():
    """The deterministic inputs of the reference's own PCS test, `test_verify_two_adic_pcs_inner`
    (crates/recursion/circuit/src/fri.rs:817-871): two 2^19 x 100 matrices with entries from_canonical_u32(i),
    inner_fri_config (blowup 2, 84 queries, 16 PoW bits), transcript = observe(commit), zeta = sample_ext,
    every matrix opened at zeta.  The commitment must equal the oracle's and the proof must be accepted by the
    transliterated verifier (the reference asserts exactly `pcs.verify(...).unwrap()`)."""
    ctx = backends.gpu()
    h, w = 1 << 19, 100
    canon = (np.arange(h * w, dtype=np.uint64) % P).astype(np.uint32).reshape(h, w)
    mat = ob.to_monty(canon)
    one = ob.lib().ork_to_monty(1)
    root, pd = ctx.commit([mat, mat], [one, one], 1)
    tree = ob.pcs_commit([mat, mat], 1)
    assert (root == tree.root).all()
    och = bf.new_challenger()
    bf.observe(och, root)
    zeta = bf.sample_ext(och)
    ch_v = bf.Challenger.from_words(och.words())
    dch = Challenger(ctx, och.words())
    pts = [[zeta], [zeta]]
    proof = pcs_open(ctx, [pd], pts, dch, 1, 84, 16)
    n_mats, hs, ws = bf.shapes_of([tree])
    assert bf.pcs_verify([root], n_mats, hs, ws, pts, ch_v, proof, 1, 84, 16) == 1
    assert (dch.w == ch_v.words()).all()
    # opened values = p_j(zeta): spot-check column 0 against the oracle's own opening of the same commitment
    och2 = bf.Challenger.from_words(och.words())
    wit = int(proof[2 * w * 4 + 19 * 8 + 4])  # opened values, 19 commit-phase roots, final_poly, then the witness
    proof_o = bf.pcs_open([tree], pts, och2, 1, 84, 16, inject_witness=wit)
    assert (proof_o == proof).all()
    pd.free()
