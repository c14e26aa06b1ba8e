"""Size-independent properties at BASELINE.json's full size (2^20 x 256, blowup 2) through the C ABI, where
the oracle would take too long to recompute everything:
  * column-sum identity: for a polynomial of degree < N evaluated over a coset of 2N points,
    sum_x p(x) = 2N * c_0 = 2 * sum of the trace column (all other monomials sum to zero over the coset);
  * random Mmcs openings of the committed LDE verify against the root with the oracle's transliteration of
    verify_batch (crates/recursion/circuit/src/fri.rs:363-405);
  * the first LDE rows are what the barycentric opening reads: Pcs::open at a random point agrees with the
    oracle's verifier on the same commitment (full FRI verify, 84 queries);
  * linearity of the LDE;  the streaming (host) and device-resident commits agree."""
import numpy as np
import pytest

from oracle import binding as ob
from oracle import binding_fri as bf
from tests import backends, util
from zkmips_b200 import Challenger, pcs_open

P = util.P
ONE = 0x01FFFFFE
pytestmark = pytest.mark.gpu


def _trace(log_n, w, seed):
    n = (1 << log_n) * w
    out = np.empty(n, np.uint32)
    step = 1 << 24
    for off in range(0, n, step):
        m = min(step, n - off)
        out[off:off + m] = util.monty(util.splitmix64(seed + off // step, m))
    return out.reshape(1 << log_n, w)


def _colsum(mat):
    return (mat.astype(np.uint64).sum(axis=0) % P).astype(np.uint64)


def test_full_size_commit_properties():
    ctx = backends.gpu()
    log_n, w = 20, 256
    tr = _trace(log_n, w, 0xC0FFEE)
    root, pd = ctx.commit([tr], [ONE], 1)
    lde = pd.lde(0)
    assert lde.shape == (1 << (log_n + 1), w)
    # Montgomery form is linear, so the identity holds on the Montgomery words as well
    assert (_colsum(lde) == (2 * _colsum(tr)) % P).all()
    # pinned host memory takes the streaming path (slabs, resumable sponge), pageable the contiguous one
    import torch
    pinned = torch.from_numpy(tr.view(np.int32)).pin_memory().numpy().view(np.uint32)
    root_p, pd_p = ctx.commit([pinned], [ONE], 1)
    assert (root_p == root).all()
    pd_p.free()
    # device-resident commit gives the same root
    dptr = ctx.upload(tr)
    root2, pd2 = ctx.commit_dev([dptr], [tr.shape], [ONE], 1)
    assert (root == root2).all()
    pd2.free()
    ctx.dev_free(dptr)
    # random openings verify
    rng = np.random.default_rng(3)
    idx = [int(x) for x in rng.integers(0, 1 << (log_n + 1), 6)] + [0, (1 << (log_n + 1)) - 1]
    opened, proofs = pd.open_batch(idx)
    for k, i in enumerate(idx):
        assert (opened[k] == lde[i]).all()
        assert ob.mmcs_verify(root, [lde.shape], i, [opened[k]], proofs[k])
    # Pcs::open on the full-size commitment, checked by the transliterated verifier
    och = bf.new_challenger()
    bf.observe(och, root)
    zeta = bf.sample_ext(och)
    ch_v = bf.Challenger.from_words(och.words())
    dch = Challenger(ctx, och.words())
    proof = pcs_open(ctx, [pd], [[zeta]], dch, 1, 84, 16)
    assert bf.pcs_verify([root], [1], [lde.shape[0]], [w], [[zeta]], ch_v, proof, 1, 84, 16) == 1
    assert (dch.w == ch_v.words()).all()
    pd.free()


def test_lde_linearity_large():
    ctx = backends.gpu()
    a, b = _trace(20, 16, 1), _trace(20, 16, 2)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    shift = 0x05FFFFFA  # GENERATOR
    la, lb, ls = ctx.coset_lde(a, 1, shift), ctx.coset_lde(b, 1, shift), ctx.coset_lde(s, 1, shift)
    assert (((la.astype(np.uint64) + lb) % P).astype(np.uint32) == ls).all()


def test_max_height_and_two_adicity_limit():
    """2^22-row trace (MAX_CPU_LOG_DEGREE, crates/core/machine/src/cpu/mod.rs:8): passes [2, 10, 10]; blowup 4
    reaches the two-adicity of the field (2^24) and one more bit must be refused."""
    from zkmips_b200 import ZkError
    ctx = backends.gpu()
    tr = _trace(22, 2, 7)
    root, pd = ctx.commit([tr], [ONE], 2)
    lde = pd.lde(0)
    assert lde.shape[0] == 1 << 24
    assert (_colsum(lde) == (4 * _colsum(tr)) % P).all()
    opened, proofs = pd.open_batch([12345678])
    assert ob.mmcs_verify(root, [lde.shape], 12345678, [opened[0]], proofs[0])
    pd.free()
    with pytest.raises(ZkError):
        ctx.commit([tr], [ONE], 3)


@pytest.mark.parametrize("kind", ["a", "b"])
def test_config2_commit_bit_exact_with_oracle(kind):
    """BASELINE config 2 at FULL size (2^20 x 256, blowup 2), both inputs of SURVEY 8(d): (a) (r*W + c) mod p,
    (b) splitmix64(0x5A4B4D49).  The oracle's Pcs::commit of the same trace takes seconds on the box's host cores,
    so the headline configuration is pinned bit for bit: root, three sampled digest layers (leaves, middle, near
    the top), and LDE rows + Merkle paths at sampled indices -- from pinned host memory (streaming commit) and from
    a device-resident trace (zk_commit_dev), which is the path bench.py's `value` times."""
    import torch
    ctx = backends.gpu()
    tr = util.config2_trace(kind)
    tree = ob.pcs_commit([tr], 1)
    pinned = torch.from_numpy(tr.view(np.int32)).pin_memory().numpy().view(np.uint32)
    root, pd = ctx.commit([pinned], [ONE], 1)
    assert (root == tree.root).all(), "streaming commit: root differs from the oracle"
    for layer in (0, 9, 17):
        assert (pd.layer(layer) == tree.layer(layer)).all(), f"digest layer {layer} differs from the oracle"
    idx = [0, 1, (1 << 21) - 1, 1 << 20, 1234567, 777777]
    opened, proofs = pd.open_batch(idx)
    for k, i in enumerate(idx):
        rows, proof = tree.open(i)
        assert (opened[k] == rows[0]).all(), f"LDE row {i} differs from the oracle"
        assert (proofs[k] == proof).all(), f"Merkle path {i} differs from the oracle"
    pd.free()
    dptr = ctx.upload(tr)
    root2, pd2 = ctx.commit_dev([dptr], [tr.shape], [ONE], 1)
    assert (root2 == tree.root).all(), "device-resident commit: root differs from the oracle"
    assert (pd2.layer(0)[::4099] == tree.layer(0)[::4099]).all()
    pd2.free()
    ctx.dev_free(dptr)
