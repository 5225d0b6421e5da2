"""Sharding of the batch across GPUs / ranks (host logic, no compute).

Files are independent, so the path shards by file with NO data-path collective: the shard key is the
fuse-zstd inode (/root/reference/src/main.rs:744-753; inodes are a dense descending counter, :719-742, so a modulo
balances counts).  libfzgpu.so applies the same rule for the fd entry points (fz_api.cu: ctx_for_key).
The only cross-rank traffic is measurement plumbing: a barrier and the max over ranks of the timed region.
"""
import numpy as np


def device_for_key(shard_key, n_devices):
    """GPU index that owns inode `shard_key` (same rule as fzg_decode_fd / fzg_encode_fd)."""
    return int(shard_key) % int(n_devices)


def partition_by_inode(inodes, n_devices):
    """inodes -> list of index arrays, one per device (static `ino mod n`)."""
    ino = np.asarray(inodes, dtype=np.uint64)
    dev = (ino % np.uint64(n_devices)).astype(np.int64)
    return [np.nonzero(dev == d)[0] for d in range(n_devices)]


def files_for_rank(rank, files_per_gpu):
    """Weak scaling of bench.py: rank r owns the contiguous file indices [r * F, (r + 1) * F)."""
    return range(rank * files_per_gpu, (rank + 1) * files_per_gpu)


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks of the default process group (identity without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
