"""N > 1 host logic on CPU: world_size-2 gloo process group (the GPU path uses nccl for the same two calls)."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

shard = importlib.import_module("fuse-zstd_b200.shard")


def test_partition_by_inode_is_a_disjoint_cover():
    inodes = np.array([2**64 - 1 - i for i in range(1000)], dtype=np.uint64)      # descending counter, src/main.rs:719-742
    for n in (1, 2, 4, 8):
        parts = shard.partition_by_inode(inodes, n)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(inodes)))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        for d, p in enumerate(parts):
            assert all(shard.device_for_key(int(inodes[i]), n) == d for i in p[:20])


def test_weak_scaling_file_ranges_are_disjoint():
    a, b = shard.files_for_rank(0, 10000), shard.files_for_rank(1, 10000)
    assert a[-1] + 1 == b[0] and len(a) == len(b) == 10000


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        files = shard.files_for_rank(rank, 100)
        dist.barrier()
        t = shard.max_over_ranks(10.0 + rank)                  # the slowest rank defines the step time
        total = shard.sum_over_ranks(len(files))
        q.put((rank, files[0], t, total))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_max_time_and_total_units():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(60)
    assert [r[1] for r in res] == [0, 100]
    assert all(r[2] == 11.0 for r in res)                      # max over ranks
    assert all(r[3] == 200 for r in res)                       # whole-job units = sum over ranks
