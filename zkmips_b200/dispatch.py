"""Shard dispatch across GPUs (one process and one zk_ctx per GPU).

The reference proves the shards of a batch with `records.into_par_iter()` inside one process
(crates/core/machine/src/utils/prove.rs:487-521); shards are independent (the challenger is cloned per
shard, prove.rs:496), so here shard i simply goes to rank i mod world and no collective touches the data
path.  `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used only to gather
the 8-word commitments in shard order and to reduce timings to the maximum over ranks."""
import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs next to GPU `device_index` (NVML's ideal affinity) BEFORE it allocates
    its pinned trace buffers: first-touch then puts them on the GPU's own NUMA node, so that with one process per
    GPU the uploads do not all read one socket's memory.  (The 8-GPU pool box is a 32-vCPU guest with a single NUMA
    node, so the binding changes nothing there: its aggregate upload rate tops out at ~176 GB/s = 44 Gelem/s for
    4 and for 8 GPUs, profiles/r1_bench_8gpu.json.)  Returns the CPU list it bound to, or None when NVML / the
    affinity call is unavailable."""
    try:
        import os

        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            device_index = int(vis.split(",")[device_index])  # NVML enumerates the physical devices
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def shards_for_rank(n_shards, world, rank):
    """round-robin placement: shard i -> rank i mod world"""
    return list(range(rank, n_shards, world))


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_commitments(local_roots, n_shards):
    """local_roots: {shard index: 8 uint32 words} of this rank.  Returns an (n_shards, 8) uint32 array in shard
    order on every rank (one all_gather of a padded int64 tensor)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    per_rank = (n_shards + world - 1) // world
    buf = torch.full((per_rank, 9), -1, dtype=torch.int64)
    for k, (idx, root) in enumerate(sorted(local_roots.items())):
        buf[k, 0] = idx
        buf[k, 1:] = torch.from_numpy(np.asarray(root, dtype=np.int64))
    if world == 1:
        parts = [buf]
    else:
        dev = _device()
        parts = [torch.empty_like(buf, device=dev) for _ in range(world)]
        dist.all_gather(parts, buf.to(dev))
        parts = [p.cpu() for p in parts]
    out = np.zeros((n_shards, 8), np.uint32)
    seen = 0
    for p in parts:
        for row in p.tolist():
            if row[0] >= 0:
                out[row[0]] = np.array(row[1:], dtype=np.uint32)
                seen += 1
    if seen != n_shards:
        raise RuntimeError(f"gathered {seen} commitments, expected {n_shards}")
    return out


def max_over_ranks(value):
    """device-time reductions for multi-GPU numbers are the MAX over ranks"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
