"""CPU tests that PIN the oracle (prompt section 3): the reference's Poseidon2 known-answer test, the
reference's own C++ field class / permutation compiled into oracle/_ref (when available), the golden
vectors committed under tests/golden/, and definition-level checks of LDE / MMCS semantics."""
import json
import os

import numpy as np
import pytest

from oracle import binding as ob
from tests import util

P = util.P
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _guest_sponge(data: bytes) -> str:
    """crates/zkvm/lib/src/poseidon2.rs:32-70 (pad 1*01 to a multiple of 3 bytes, 3 bytes per field
    element, rate-8 overwrite sponge on canonical words)."""
    l = len(data)
    new = (l + 3) // 3 * 3
    pad = bytearray(data) + bytes(new - l)
    if l % 3 == 2:
        pad[l] = 0b10000001
    else:
        pad[l] = 1
        pad[new - 1] = 0b10000000
    felts = [int.from_bytes(pad[i:i + 3], "little") for i in range(0, new, 3)]
    st = np.zeros(16, np.uint32)
    for i in range(0, len(felts), 8):
        chunk = felts[i:i + 8]
        st[:len(chunk)] = chunk
        st = ob.permute_canonical(st)
    return st.astype("<u4").tobytes()[:32].hex()


def test_poseidon2_known_answer():
    # examples/poseidon2/host/src/main.rs:11,33-37
    assert _guest_sponge(bytes([1] * 1000)) == \
        "ae45b14fe23b9f584c76c67d4d9ef6635a27b553a7114427584cc87ba8919866"


def test_golden_vectors():
    with open(os.path.join(GOLDEN, "poseidon2_vectors.json")) as fh:
        g = json.load(fh)
    for v in g["permute_canonical"]:
        assert ob.permute_canonical(np.array(v["in"], np.uint32)).tolist() == v["out"]
    for v in g["field_mul_monty"]:
        assert ob.lib().ork_mul(v["a"], v["b"]) == v["out"]
    for v in g["field_inv_monty"]:
        assert ob.lib().ork_inv(v["a"]) == v["out"]
    for v in g["sponge_hex"]:
        assert _guest_sponge(bytes.fromhex(v["in"])) == v["out"]


def test_against_reference_cpp():
    R = ob.ref()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    L = ob.lib()
    rng = np.random.default_rng(7)
    edge = [0, 1, 2, P - 1, P - 2, 0x01FFFFFE, 0x7EFFFFFF]
    vals = edge + [int(x) for x in rng.integers(0, P, 500)]
    for a in vals:
        assert L.ork_to_monty(a) == R.ref_to_monty(a)
        assert L.ork_from_monty(a) == R.ref_from_monty(a)
        if a:
            assert L.ork_inv(a) == R.ref_inv(a)
        for b in vals[:40]:
            assert L.ork_mul(a, b) == R.ref_mul(a, b)
    for _ in range(200):
        s = rng.integers(0, P, 16).astype(np.uint32)
        r = s.copy()
        R.ref_poseidon2_permute(r.ctypes.data_as(ob._u32p))
        assert (ob.permute(s) == r).all()


def test_field_constants():
    L = ob.lib()
    assert L.ork_to_monty(1) == 0x01FFFFFE  # kb31_t.hpp:29
    assert L.ork_from_monty(L.ork_two_adic_generator(24)) == 0x6AC49F88 == pow(3, 127, P)
    assert L.ork_from_monty(L.ork_two_adic_generator(1)) == P - 1
    for k in range(1, 24):
        g = L.ork_two_adic_generator(k + 1)
        assert L.ork_mul(g, g) == L.ork_two_adic_generator(k)


def test_extension_field():
    rng = np.random.default_rng(3)
    for _ in range(50):
        a = [int(x) for x in rng.integers(0, P, 4)]
        b = [int(x) for x in rng.integers(0, P, 4)]
        am, bm = util.monty(a), util.monty(b)
        out = np.empty(4, np.uint32)
        ob.lib().ork_ext_mul(am.ctypes.data_as(ob._u32p), bm.ctypes.data_as(ob._u32p),
                             out.ctypes.data_as(ob._u32p))
        assert ob.from_monty(out).tolist() == util.ext_mul(a, b)
        inv = np.empty(4, np.uint32)
        ob.lib().ork_ext_inv(am.ctypes.data_as(ob._u32p), inv.ctypes.data_as(ob._u32p))
        assert util.ext_mul(a, ob.from_monty(inv).tolist()) == [1, 0, 0, 0]


def test_sponge_and_compress_semantics():
    # recursion/circuit/src/hash.rs:40-49 (overwrite mode, no padding) and :76-81
    x = util.monty(util.canon_matrix(1, 21)[0])
    st = np.zeros(16, np.uint32)
    for off in range(0, 21, 8):
        ch = x[off:off + 8]
        st[:len(ch)] = ch
        st = ob.permute(st)
    assert (ob.hash_slice(x) == st[:8]).all()
    assert (ob.hash_slice(np.zeros(0, np.uint32)) == 0).all()  # empty input: zero state, no permutation
    l, r = x[:8], x[8:16]
    assert (ob.compress(l, r) == ob.permute(np.concatenate([l, r]))[:8]).all()


@pytest.mark.parametrize("log_h,w,log_blowup", [(0, 3, 1), (1, 2, 1), (3, 5, 1), (4, 2, 2), (5, 1, 1)])
def test_coset_lde_definition(log_h, w, log_blowup):
    """row r of the committed LDE holds p(shift * g^{bitrev(r)}) where p interpolates the
    input over the subgroup.  Checked against a direct O(n^2) evaluation in Python integers."""
    h = 1 << log_h
    a = util.canon_matrix(h, w).astype(object)
    g = util.two_adic_generator(log_h)
    hinv = pow(h, -1, P)
    # coefficients by the inverse DFT definition
    coef = [[sum(int(a[j][c]) * pow(g, -j * i % h if h > 1 else 0, P) for j in range(h)) * hinv % P
             for c in range(w)] for i in range(h)]
    shift = 3
    H = h << log_blowup
    G = util.two_adic_generator(log_h + log_blowup)
    lde = ob.from_monty(ob.coset_lde(util.monty(a.astype(np.uint64)), log_blowup, int(util.monty([shift])[0])))
    for r in range(H):
        x = shift * pow(G, util.bitrev(r, log_h + log_blowup), P) % P
        for c in range(w):
            want = sum(coef[i][c] * pow(x, i, P) for i in range(h)) % P
            assert int(lde[r][c]) == want
    # restricted to a subgroup coset with shift 1 the LDE reproduces the input (blowup rows j*2^b)
    lde1 = ob.from_monty(ob.coset_lde(util.monty(a.astype(np.uint64)), log_blowup, int(util.monty([1])[0])))
    for j in range(h):
        r = util.bitrev(j << log_blowup, log_h + log_blowup)
        assert lde1[r].tolist() == [int(v) for v in a[j]]


def test_dft_batch_definition():
    h, w = 8, 3
    a = util.canon_matrix(h, w)
    g = util.two_adic_generator(3)
    out = ob.from_monty(ob.dft_batch(util.monty(a)))
    for k in range(h):
        for c in range(w):
            assert int(out[k][c]) == sum(int(a[j][c]) * pow(g, j * k, P) for j in range(h)) % P


def _mmcs_reference(mats):
    """MerkleTree::new semantics written directly from SURVEY A.5 with the scalar oracle hash."""
    order = sorted(range(len(mats)), key=lambda i: -mats[i].shape[0])  # stable
    hmax = mats[order[0]].shape[0]
    k = 0
    tall = []
    while k < len(order) and mats[order[k]].shape[0] == hmax:
        tall.append(mats[order[k]])
        k += 1
    layer = [ob.hash_slice(np.concatenate([m[r] for m in tall])) for r in range(hmax)]
    layers = [layer]
    while len(layer) > 1:
        n = len(layer) // 2
        inj = []
        while k < len(order) and mats[order[k]].shape[0] == n:
            inj.append(mats[order[k]])
            k += 1
        nxt = []
        for i in range(n):
            d = ob.compress(layer[2 * i], layer[2 * i + 1])
            if inj:
                d = ob.compress(d, ob.hash_slice(np.concatenate([m[i] for m in inj])))
            nxt.append(d)
        layer = nxt
        layers.append(layer)
    return layers


@pytest.mark.parametrize("dims", [[(8, 3)], [(8, 3), (8, 9)], [(4, 2), (16, 5), (4, 7), (1, 3)],
                                  [(1, 4)], [(2, 8), (2, 8), (1, 1)], [(16, 0), (16, 3)]])
def test_mmcs_commit_open_verify(dims):
    mats = [util.monty(util.canon_matrix(h, max(w, 1), seed=11 + i))[:, :w] for i, (h, w) in enumerate(dims)]
    mats = [np.ascontiguousarray(m) for m in mats]
    tree = ob.mmcs_commit(mats)
    layers = _mmcs_reference(mats)
    assert (tree.root == layers[-1][0]).all()
    for l, layer in enumerate(layers):
        assert (tree.layer(l) == np.array(layer)).all()
    hmax = max(h for h, _ in dims)
    for index in range(hmax):
        rows, proof = tree.open(index)
        for (h, w), m, row in zip(dims, mats, rows):
            assert (row == m[index >> (tree.log_max_height - (h.bit_length() - 1))]).all()
        assert ob.mmcs_verify(tree.root, dims, index, rows, proof)
        if proof.shape[0]:
            bad = proof.copy()
            bad[0, 0] ^= 1
            assert not ob.mmcs_verify(tree.root, dims, index, rows, bad)
        if rows and rows[0].size:
            rows2 = [r.copy() for r in rows]
            rows2[0][0] ^= 1
            assert not ob.mmcs_verify(tree.root, dims, index, rows2, proof)


def test_pcs_commit_is_lde_then_mmcs():
    mats = [util.monty(util.canon_matrix(16, 5)), util.monty(util.canon_matrix(4, 3)),
            util.monty(util.canon_matrix(16, 2, seed=5))]
    one = int(util.monty([1])[0])
    g5 = int(util.monty([util.two_adic_generator(5)])[0])
    gen = int(util.monty([3])[0])
    # third matrix lives on the coset 3*g_32 (a quotient-chunk domain): shift = 3 / (3 g) = 1/g
    dshift = ob.lib().ork_mul(gen, g5)
    tree = ob.pcs_commit(mats, 1, [one, one, dshift])
    ldes = [ob.coset_lde(mats[0], 1, gen), ob.coset_lde(mats[1], 1, gen),
            ob.coset_lde(mats[2], 1, ob.lib().ork_inv(g5))]
    tree2 = ob.mmcs_commit(ldes)
    assert (tree.root == tree2.root).all()
    for i in range(3):
        assert (tree.matrix(i) == ldes[i]).all()
