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


def test_config3_partition_is_what_bench_uses():
    """bench.py --config 3: rank r of n decodes the files whose inode (2^64 - 1 - i, the reference's descending counter) is r mod n;
    the ranks' shares are disjoint, cover the corpus and differ by at most one file"""
    import argparse
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    args = argparse.Namespace(config=3, files=1001, file_size=1 << 20, level=3)
    for world in (1, 2, 4, 8):
        parts = [bench.rank_files(args, r, world) for r in range(world)]
        assert all(p[1] == "strong" for p in parts)
        idx = np.sort(np.concatenate([p[0] for p in parts]))
        assert np.array_equal(idx, np.arange(1001))
        sizes = [len(p[0]) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        for r, p in enumerate(parts):
            assert all((2**64 - 1 - int(i)) % world == r for i in p[0][:16])
    args2 = argparse.Namespace(config=2, files=100, file_size=1 << 20, level=3)
    a, b = bench.rank_files(args2, 0, 2), bench.rank_files(args2, 1, 2)
    assert a[1] == "weak" and a[0][-1] + 1 == b[0][0] and len(a[0]) == len(b[0]) == 100


def test_reference_arm_runs_for_every_config():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line per configuration, same metric and unit as the GPU arm"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cfg, extra in ((2, ["--files", "24"]), (3, ["--files", "24"]), (4, ["--files", "2", "--file-size", str(8 << 20)]), (5, ["--files", "4"])):
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", str(cfg), "--steps", "1", "--warmup", "1"] + extra,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-400:]
        line = json.loads(r.stdout.decode().strip().splitlines()[-1])
        if "unavailable" in line:
            continue
        assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
        assert line["metric"] == ("zstd_encode_uncompressed_GBps" if cfg == 5 else "zstd_decode_uncompressed_GBps")
        assert line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0
