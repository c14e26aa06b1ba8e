"""CPU tests of the oracle's FRI/PCS layer: the prover restatement must be accepted by the transliterated
in-repo verifier (crates/recursion/circuit/src/fri.rs:34-405), tampering must be rejected, and the
challenger must follow crates/recursion/circuit/src/challenger.rs."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from oracle import binding_fri as bf
from tests import util

P = util.P


def _mont(h, w, seed=0):
    return ob.to_monty(util.canon_matrix(h, w, "rand", seed=0x5A4B4D49 + seed))


def test_challenger_semantics():
    """observe clears the output buffer and duplexes at 8 inputs; sample pops from the END of state[0..8]."""
    L = ob.lib()
    ch = bf.new_challenger()
    vals = ob.to_monty(np.arange(1, 9, dtype=np.uint32))
    bf.observe(ch, vals)  # exactly RATE inputs -> one duplexing
    st = np.zeros(16, np.uint32)
    st[:8] = vals
    exp = ob.permute(st)
    assert ch.n_in == 0 and ch.n_out == 8
    assert list(ch.out) == list(exp[:8])
    assert L.ork_ch_sample(C.byref(ch)) == exp[7]
    assert L.ork_ch_sample(C.byref(ch)) == exp[6]
    # a pending input forces a new duplexing on the next sample
    bf.observe(ch, vals[:1])
    assert ch.n_out == 0
    st2 = exp.copy()
    st2[0] = vals[0]
    exp2 = ob.permute(st2)
    e = bf.sample_ext(ch)
    assert list(e) == [exp2[7], exp2[6], exp2[5], exp2[4]]
    # sample_bits = low bits of the canonical value
    v = L.ork_ch_sample_bits(C.byref(ch), 5)
    assert v == (int(ob.from_monty(np.array([exp2[3]], np.uint32))[0]) & 31)


def test_grind_then_check():
    L = ob.lib()
    ch = bf.new_challenger(ob.to_monty(np.arange(3, dtype=np.uint32)))
    ch2 = bf.Challenger.from_words(ch.words())
    w = L.ork_ch_grind(C.byref(ch), 8)
    assert L.ork_ch_check_witness(C.byref(ch2), 8, w) == 1
    assert (ch.words() == ch2.words()).all()


def _open_and_verify(mats_per_round, log_blowup, n_points_fn, num_queries=6, pow_bits=4):
    L = ob.lib()
    trees = [ob.pcs_commit(ms, log_blowup) for ms in mats_per_round]
    ch = bf.new_challenger()
    for t in trees:
        bf.observe(ch, t.root)
    zeta = bf.sample_ext(ch)
    pts = []
    for ms in mats_per_round:
        for m in ms:
            h = m.shape[0]
            g = L.ork_two_adic_generator(int(np.log2(h)))
            zg = np.array([L.ork_mul(int(z), g) for z in zeta], np.uint32)
            pts.append([zeta, zg][:n_points_fn(m)])
    ch_v = bf.Challenger.from_words(ch.words())
    proof = bf.pcs_open(trees, pts, ch, log_blowup, num_queries, pow_bits)
    n_mats, hs, ws = bf.shapes_of(trees)
    roots = [t.root for t in trees]
    rc = bf.pcs_verify(roots, n_mats, hs, ws, pts, ch_v, proof, log_blowup, num_queries, pow_bits)
    assert rc == 1, rc
    assert (ch.words() == ch_v.words()).all()  # prover and verifier transcripts end in the same state
    return trees, pts, proof, roots, (n_mats, hs, ws)


def test_open_verify_single_matrix():
    _open_and_verify([[_mont(64, 5)]], 1, lambda m: 2)


def test_open_verify_mixed_heights_rounds_blowup2():
    mats = [[_mont(32, 3, 1), _mont(8, 6, 2)], [_mont(32, 4, 3), _mont(2, 2, 4), _mont(16, 1, 5)], [_mont(1, 3, 6)]]
    _open_and_verify(mats, 2, lambda m: 1 if m.shape[1] == 4 else 2)


def test_verify_rejects_tampering():
    L = ob.lib()
    mats = [[_mont(32, 3, 1), _mont(8, 6, 2)]]
    trees, pts, proof, roots, (n_mats, hs, ws) = _open_and_verify(mats, 1, lambda m: 2)
    ch0 = bf.new_challenger()
    bf.observe(ch0, roots[0])
    bf.sample_ext(ch0)
    for pos in [0, 5, proof.size // 2, proof.size - 1]:
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % P
        ch = bf.Challenger.from_words(ch0.words())
        assert bf.pcs_verify(roots, n_mats, hs, ws, pts, ch, bad, 1, 6, 4) != 1, pos
    ch = bf.Challenger.from_words(ch0.words())
    assert bf.pcs_verify(roots, n_mats, hs, ws, pts, ch, proof, 1, 6, 4) == 1


def test_opened_values_are_polynomial_evaluations():
    """ys must equal direct evaluation of the interpolating polynomial (pure-Python check, tiny)."""
    h, w = 8, 2
    canon = util.canon_matrix(h, w, "rand", seed=77)
    m = ob.to_monty(canon)
    trees = [ob.pcs_commit([m], 1)]
    ch = bf.new_challenger()
    z = ob.to_monty(np.array([5, 6, 7, 8], np.uint32))
    proof = bf.pcs_open(trees, [[z]], ch, 1, 2, 1)
    ys = ob.from_monty(proof[:w * 4]).reshape(w, 4)
    # coefficients by naive inverse DFT over canonical ints
    g = util.two_adic_generator(3)
    for c in range(w):
        coeffs = []
        for k in range(h):
            acc = 0
            for j in range(h):
                acc = (acc + int(canon[j, c]) * pow(g, -j * k % h, P)) % P
            coeffs.append(acc * pow(h, -1, P) % P)
        zc = [5, 6, 7, 8]
        acc = [0, 0, 0, 0]
        zp = [1, 0, 0, 0]
        for k in range(h):
            acc = util.ext_add(acc, [x * coeffs[k] % P for x in zp])
            zp = util.ext_mul(zp, zc)
        assert list(ys[c]) == acc
