"""world_size-2 gloo test of the multi-GPU path (shard placement, commitment gather, max-over-ranks timing),
with the emulator build standing in for the per-rank device.  The roots gathered in shard order must equal
the oracle's."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(i):
    from oracle import binding as ob
    return ob.to_monty(util.canon_matrix(32, 8 + 8 * (i % 2), "rand", seed=900 + i))


def _worker(rank, world, port, n_shards, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import backends
    from zkmips_b200 import dispatch
    ctx = backends.emu()
    mine = dispatch.shards_for_rank(n_shards, world, rank)
    roots = {}
    for i in mine:
        root, pd = ctx.commit([_shard(i)], [0x01FFFFFE], 1)
        roots[i] = root
        pd.free()
    allr = dispatch.gather_commitments(roots, n_shards)
    t = dispatch.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((mine, allr, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dispatch_and_gather():
    from oracle import binding as ob
    from tests import backends
    backends.build_emu()  # build once, before forking
    n_shards, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_shards, q)) for r in range(world)]
    for p in procs:
        p.start()
    mine, allr, t = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert mine == [0, 2, 4]
    assert t == 11.0
    for i in range(n_shards):
        assert (allr[i] == ob.pcs_commit([_shard(i)], 1).root).all()


def test_placement_covers_all_shards():
    from zkmips_b200 import dispatch
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in dispatch.shards_for_rank(13, world, r))
        assert got == list(range(13))
