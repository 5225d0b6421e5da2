#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its config 2:
batched decode of 10 000 synthetic 1 MiB JSON files compressed at zstd level 3 (reference-writer
framing: FCS + XXH64, /root/reference/src/main.rs:781-791), device-resident, per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the hot path (fzg_decode_batch, the replacement of copy_decode at
/root/reference/src/main.rs:463-467) over the whole batch.  N > 1: one process per GPU (torchrun), the files
are sharded by inode (rank r owns files r*F .. r*F+F-1), no data-path collective; the time is the max over ranks.

  value         uncompressed-output GB/s, inputs and outputs resident in HBM (CUDA events on the codec stream)
  e2e           same metric through the C ABI with pinned HOST buffers (H2D + decode + D2H inside the timed region)
  roofline      dominant kernel: (compressed in + uncompressed out) bytes / its CUDA-event time vs measured HBM GB/s
  cpu_baseline  the reference's CPU path (zstd-rs copy_decode restated on libzstd, oracle/_ref) on this box's cores

--impl reference times only that CPU path (all host threads, the same files per step as the GPU arm).

--config selects BASELINE.json's other configurations (the default, 2, is the headline and what the driver runs):
  3  the 64 GiB corpus (65 536 x 1 MiB), partitioned over the ranks by `inode mod n` (shard.partition_by_inode; file i
     carries inode 2^64 - 1 - i, fuse-zstd's descending counter, src/main.rs:719-753): STRONG scaling, total work fixed
  4  large single-frame files (default 16 x 1 GiB per GPU), windowLog 23 (8 MiB window); level 3 stands in for level 19, which
     compresses ~1 MB/s per core (a true level-19 file is in the parity tests): long match distances, few frames
  5  the write path: fzg_encode_batch of N x 4 MiB plain files per GPU (default 10 000), every emitted file of a sample
     decoded by libzstd and compared; ratio against libzstd level 3 reported
"""
import argparse
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "zstd_decode_uncompressed_GBps"
ENC_METRIC = "zstd_encode_uncompressed_GBps"
ENC_FIRST = 7000000          # first file index of the config-5 corpus
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (see the module docstring)")
    ap.add_argument("--files", type=int, default=0, help="files per GPU (configs 2, 4, 5: 10 000 / 16 / 10 000) or in total (config 3: 65 536)")
    ap.add_argument("--file-size", type=int, default=0, help="bytes per file (configs 2, 3: 1 MiB; 4: 1 GiB; 5: 4 MiB)")
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--e2e-files", type=int, default=0, help="files per GPU in the host-buffer (e2e) leg (0: all, or half when host RAM is short)")
    ap.add_argument("--cpu-files", type=int, default=0, help="files per step of the CPU arm (0: the inline cpu_baseline leg takes a bounded sample, --impl reference all the files of a step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mount", action="store_true", help="skip the read-through-the-mount leg (tools/mount_bench.py)")
    a = ap.parse_args()
    a.files = a.files or {2: 10000, 3: 65536, 4: 16, 5: 10000}[a.config]
    a.file_size = a.file_size or {2: 1 << 20, 3: 1 << 20, 4: 1 << 30, 5: 4 << 20}[a.config]
    return a


# ----------------------------------------------------------------------------------------------- workload
class Workload:
    """The files `indices` of the synthetic corpus, S bytes each: plain -> level-`level` frames by the reference-writer
    restatement (oracle/_ref, libzstd; window_log != 0 sets ZSTD_c_windowLog); the packed compressed bytes stay on the host
    (numpy), offsets 16-byte aligned."""

    def __init__(self, indices, size, level, threads, keep_plain=512, window_log=0):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        corpus = importlib.import_module("fuse-zstd_b200.corpus")
        pyoracle.build()
        self.ref = pyoracle.Ref()
        idx = np.asarray(list(indices), dtype=np.int64)
        n = len(idx)
        self.n, self.size = n, size
        chunk = max(1, min(500, (1 << 30) // max(size, 1)))            # <= 1 GiB of plain bytes on the host at a time
        keep_plain = min(keep_plain, max(1, (64 << 20) // max(size, 1)))
        self.comp_len = np.zeros(n, dtype=np.uint64)
        self.comp_off = np.zeros(n, dtype=np.uint64)
        parts, off = [], 0
        self.plain_head = None
        bound = (self.ref.bound(size) if self.ref.available else size + size // 128 + 512) + 64
        contiguous = n > 0 and bool((np.diff(idx) == 1).all())
        t_gen = t_cmp = 0.0
        for c0 in range(0, n, chunk):
            m = min(chunk, n - c0)
            t0 = time.time()
            plain = corpus.json_files(int(idx[c0]), m, size, threads=threads) if contiguous else corpus.json_files_idx(idx[c0:c0 + m], size, threads=threads)
            t1 = time.time()
            comp = np.empty((m, bound), dtype=np.uint8)
            if self.ref.available:
                sp = plain.ctypes.data + np.arange(m, dtype=np.uint64) * np.uint64(size)
                dp = comp.ctypes.data + np.arange(m, dtype=np.uint64) * np.uint64(bound)
                _, ol, st = self.ref.batch(2, threads, sp.astype(np.uint64), np.full(m, size, dtype=np.uint64),
                                           dp.astype(np.uint64), np.full(m, bound, dtype=np.uint64), level | (window_log << 8))
                assert not st.any(), "libzstd encode failed"
            else:                                   # same image should carry libzstd; pyarrow's bundled zstd otherwise
                import pyarrow as pa
                from concurrent.futures import ThreadPoolExecutor
                codec = pa.Codec("zstd", compression_level=level)
                def one(i):
                    b = codec.compress(plain[i].tobytes(), asbytes=True)
                    comp[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
                    return len(b)
                with ThreadPoolExecutor(threads) as ex:
                    ol = np.array(list(ex.map(one, range(m))), dtype=np.uint64)
            t2 = time.time()
            t_gen += t1 - t0; t_cmp += t2 - t1
            for i in range(m):
                L = int(ol[i])
                self.comp_len[c0 + i] = L; self.comp_off[c0 + i] = off
                off += (L + 15 & ~15) + 16
            parts.append([comp[j, :int(ol[j])].copy() for j in range(m)])
            del comp
            if c0 == 0:
                k = min(keep_plain, m)
                self.plain_head = plain[:k].copy()
        self.packed = np.zeros(off + 64, dtype=np.uint8)
        i = 0
        for part in parts:
            for blob in part:
                o = int(self.comp_off[i])
                self.packed[o:o + len(blob)] = blob
                i += 1
        del parts
        self.comp_bytes = int(self.comp_len.sum())
        self.plain_bytes = n * size
        self.gen_s, self.cmp_s = t_gen, t_cmp


def rank_files(args, rank, world):
    """The file indices rank `rank` decodes, the scaling mode, and the name of the workload."""
    shard = importlib.import_module("fuse-zstd_b200.shard")
    F, S = args.files, args.file_size
    if args.config == 3:
        inodes = (np.uint64(0xFFFFFFFFFFFFFFFF) - np.arange(F, dtype=np.uint64))      # fuse-zstd's descending inode counter
        mine = shard.partition_by_inode(inodes, world)[rank]
        return mine, "strong", ("config3: the %d x %d B corpus (%.1f GiB) partitioned over %d GPU(s) by inode mod n, zstd level %d, "
                                "reference-writer frames, device-resident" % (F, S, F * S / 2**30, world, args.level))
    first = rank * F
    if args.config == 4:
        return np.arange(first + 9000000, first + 9000000 + F), "weak", (
            "config4: %d x %d B single-frame files per GPU, windowLog 23 (8 MiB window), zstd level %d standing in for level 19, "
            "reference-writer frames, device-resident" % (F, S, args.level))
    return np.arange(first, first + F), "weak", (
        "config2: batched decode of %d x %d B synthetic JSON files per GPU, zstd level %d, reference-writer frames (FCS + XXH64 "
        "verified), device-resident" % (F, S, args.level))


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_ev, self.sm, self.reasons, self.max_mhz, self.err = index, threading.Event(), [], set(), None, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_ev.is_set():
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.stop_ev.wait(0.05)
        except Exception as e:      # pragma: no cover
            self.err = repr(e)

    def result(self):
        self.stop_ev.set(); self.join(2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm), **({"error": self.err} if self.err else {})}


def gpu_index_for_nvml(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference(w, threads, n_files, passes, oneshot=False):
    """copy_decode restated on libzstd (oracle/_ref/libfzref.so), `threads` host threads, one file per thread at a
    time.  Returns GB/s of uncompressed output over the timed passes (first pass untimed)."""
    n = min(n_files, w.n)
    out = np.empty((n, w.size), dtype=np.uint8)
    sp = (w.packed.ctypes.data + w.comp_off[:n]).astype(np.uint64)
    dp = (out.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(w.size)).astype(np.uint64)
    dc = np.full(n, w.size, dtype=np.uint64)
    times = []
    for p in range(passes + 1):
        t, ol, st = w.ref.batch(1 if oneshot else 0, threads, sp, w.comp_len[:n].copy(), dp, dc)
        assert not st.any() and (ol == w.size).all(), "libzstd decode failed"
        if p:
            times.append(t)
    k = min(n, len(w.plain_head))
    assert hashlib.sha256(out[:k].tobytes()).digest() == hashlib.sha256(w.plain_head[:k].tobytes()).digest()
    return n * w.size / 1e9 / (sum(times) / len(times)), n


def run_reference(args, rank):
    """The reference's own CPU implementation of the path on this box's host cores: copy_decode (configs 2-4) or the writer
    (config 5) restated call for call on libzstd (oracle/_ref), all host threads, the SAME files per step as one GPU of the
    b200 arm (--cpu-files bounds it)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.config == 5:
        return run_reference_encode(args, threads)
    mine, scaling, name = rank_files(args, 0, 1)
    if args.cpu_files:
        mine = mine[:args.cpu_files]
    w = Workload(mine, args.file_size, args.level, threads, window_log=23 if args.config == 4 else 0)
    if not w.ref.available:
        print(json.dumps({"impl": "reference", "unavailable": "libzstd.so.1 not present on this box"})); return
    n = w.n
    out = np.empty((n, w.size), dtype=np.uint8)
    sp = (w.packed.ctypes.data + w.comp_off).astype(np.uint64)
    dp = (out.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(w.size)).astype(np.uint64)
    dc = np.full(n, w.size, dtype=np.uint64)
    ts = []
    for k in range(args.warmup + args.steps):
        t, ol, st = w.ref.batch(0, threads, sp, w.comp_len.copy(), dp, dc)
        assert not st.any()
        if k >= args.warmup:
            ts.append(t)
    sec = sum(ts) / len(ts)
    val = n * w.size / 1e9 / sec
    sample = "all %d files of a step (%d B each, level %d), copy_decode restated on libzstd %d (oracle/_ref), %d threads" % (
        n, args.file_size, args.level, w.ref.version, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name.replace(", device-resident", ""), "files_per_gpu": n, "file_size": args.file_size, "level": args.level,
                   "ratio": round(w.plain_bytes / w.comp_bytes, 3)},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference_encode(args, threads):
    """config 5 on the CPU: Encoder::new(level) + set_pledged_src_size + include_checksum + finish (src/main.rs:781-791) on libzstd."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    corpus = importlib.import_module("fuse-zstd_b200.corpus")
    ref = pyoracle.Ref()
    if not ref.available:
        print(json.dumps({"impl": "reference", "unavailable": "libzstd.so.1 not present on this box"})); return
    F, S = args.files, args.file_size
    n = min(F, args.cpu_files or max(threads * 8, (8 << 30) // S))      # a step of <= 8 GiB of input: a few seconds of all cores
    plain = corpus.json_files(ENC_FIRST, n, S, threads=threads)
    bound = ref.bound(S) + 64
    comp = np.empty((n, bound), dtype=np.uint8)
    sp = (plain.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
    dp = (comp.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(bound)).astype(np.uint64)
    sl, dc = np.full(n, S, dtype=np.uint64), np.full(n, bound, dtype=np.uint64)
    ts = []
    for k in range(args.warmup + args.steps):
        t, ol, st = ref.batch(2, threads, sp, sl, dp, dc, args.level)
        assert not st.any()
        if k >= args.warmup:
            ts.append(t)
    sec = sum(ts) / len(ts)
    val = n * S / 1e9 / sec
    sample = "%d of the %d x %d B files per step, the reference's writer restated on libzstd %d level %d (oracle/_ref), %d threads" % (
        n, F, S, ref.version, args.level, threads)
    print(json.dumps({
        "impl": "reference", "metric": ENC_METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config5: level-%d encode of %d x %d B synthetic JSON files" % (args.level, F, S), "sample_files_per_step": n,
                   "ratio": round(n * S / float(ol.sum()), 3)},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa(index):
    """Pins this process to the CPUs of its GPU's NUMA node (NVML), so that the pinned host buffers of the e2e leg -- first
    touch -- and the threads that feed them are local to the GPU's PCIe root.  Returns the number of CPUs, or 0 when unknown."""
    if os.environ.get("FZG_NUMA_BIND", "1") == "0":
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 1) + 63) // 64)
        cpus = [64 * w_ + b for w_, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def host_link_ceiling(torch, shard, barrier, h_src, h_dst, in_bytes, out_bytes):
    """What the box's host link lets `e2e` reach at most, measured with the e2e leg's own pinned buffers on all ranks AT ONCE
    (tools/pcie_bw.py is the stand-alone version): H2D alone, D2H alone, both together -> GB/s summed over the ranks.  The
    ceiling models a step as: both directions busy until the compressed input (the smaller side) is in, then D2H alone."""
    n_in, n_out = min(in_bytes, 1 << 30), min(out_bytes, 1 << 30)
    d_src = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_dst = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_src.copy_(h_src[:n_in], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_dst[:n_out].copy_(d_dst, non_blocking=True)

    def timed(fs, reps=3):
        for f in fs:
            f()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fs:
                f()
        torch.cuda.synchronize()
        return shard.max_over_ranks((time.perf_counter() - t0) / reps, "cuda")
    try:
        t_h, t_d, t_b = timed([h2d]), timed([d2h]), timed([h2d, d2h])
        tot_in, tot_out = shard.sum_over_ranks(n_in, "cuda"), shard.sum_over_ranks(n_out, "cuda")
        r_h, r_d = tot_in / 1e9 / t_h, tot_out / 1e9 / t_d                      # alone
        r_b = (tot_in + tot_out) / 1e9 / t_b                                    # both directions at once, in + out
        step_in, step_out = shard.sum_over_ranks(in_bytes, "cuda") / 1e9, shard.sum_over_ranks(out_bytes, "cuda") / 1e9
        t_overlap = 2 * step_in / r_b                                           # each direction gets about half of r_b while both run
        t_step = t_overlap + max(0.0, step_out - step_in) / r_d
        return {"h2d_alone_GBps": round(r_h, 2), "d2h_alone_GBps": round(r_d, 2), "both_GBps": round(r_b, 2),
                "ceiling_GBps": round(step_out / t_step, 2), "how": "pinned copies of <= 1 GiB on every rank at once; ceiling = uncompressed "
                "bytes / (2 x in / both + (out - in) / d2h_alone)"}
    except Exception as e:       # pragma: no cover
        log("host link probe failed: %r" % (e,))
        return None


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, local_rank, world):
    numa_cpus = bind_to_gpu_numa(gpu_index_for_nvml(local_rank)) if world > 1 else 0      # before torch allocates anything
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    codec = importlib.import_module("fuse-zstd_b200.codec")
    shard = importlib.import_module("fuse-zstd_b200.shard")
    if not os.path.exists(codec.SO):
        codec.build()
    codec.init([local_rank])
    dev = local_rank
    threads = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    if numa_cpus:
        threads = max(1, min(threads, numa_cpus))
        log("[rank %d] bound to the %d CPUs of GPU %d's NUMA node" % (rank, numa_cpus, local_rank))
    if args.config == 5:
        return run_b200_encode(args, rank, local_rank, world, codec, shard, threads)
    S = args.file_size
    mine, scaling, workload_name = rank_files(args, rank, world)
    F = len(mine)                                                    # files of THIS rank
    w = Workload(mine, S, args.level, threads, window_log=23 if args.config == 4 else 0)
    log("[rank %d] corpus: %d files of %d B, ratio %.3f, gen %.1fs compress %.1fs (%d threads)" % (
        rank, F, S, w.plain_bytes / w.comp_bytes, w.gen_s, w.cmp_s, threads))

    d_src = torch.from_numpy(w.packed).cuda()
    d_dst = torch.zeros(F * S + 256, dtype=torch.uint8, device="cuda")
    sp = (d_src.data_ptr() + w.comp_off).astype(np.uint64)
    dp = (d_dst.data_ptr() + np.arange(F, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
    dc = np.full(F, S, dtype=np.uint64)
    flags = codec.SRC_DEVICE | codec.DST_DEVICE | codec.PROFILE
    stream = torch.cuda.ExternalStream(codec.stream_handle(dev), device=torch.device("cuda", dev))

    def step():
        dl, st = codec.decode_batch_ptrs(dev, sp, w.comp_len, dp, dc, flags)
        return dl, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        dl, st = step()
    assert not st.any(), "decode failed: statuses %s" % np.unique(st)
    assert (dl == S).all()
    k = len(w.plain_head)
    got = d_dst[:k * S].cpu().numpy()
    assert hashlib.sha256(got.tobytes()).digest() == hashlib.sha256(w.plain_head.tobytes()).digest(), "decoded bytes differ from the plain corpus"
    del got

    sampler = ClockSampler(gpu_index_for_nvml(local_rank)); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms, launches = {}, 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        t = codec.last_timing(dev)
        launches += t["launches"]
        for nm, ms in t["stages"].items():
            stage_ms[nm] = stage_ms.get(nm, 0.0) + ms
    ev1.record(stream)
    barrier()
    clocks = sampler.result()
    ms = ev0.elapsed_time(ev1) / args.steps
    ms = shard.max_over_ranks(ms, "cuda")
    total_plain = shard.sum_over_ranks(w.plain_bytes, "cuda")         # bytes all ranks produced per step

    # ---- e2e: pinned host buffers through the same C ABI call
    e2e = None
    if not args.no_e2e:
        if F * S > (40 << 30):          # the e2e leg stages the whole batch in HBM once more: the device-resident buffers of a batch this large must go first
            del d_dst, d_src
            torch.cuda.empty_cache()
        E = min(args.e2e_files, F) if args.e2e_files else F
        if not args.e2e_files:
            import psutil
            need = 1.5 * F * S * int(os.environ.get("LOCAL_WORLD_SIZE", world))       # pinned in + out buffers of every local rank
            if psutil.virtual_memory().available < 2 * need:
                E = F // 2
        src_bytes = int(w.comp_off[E - 1] + w.comp_len[E - 1])
        h_src = torch.empty(src_bytes + 64, dtype=torch.uint8, pin_memory=True)
        h_src.numpy()[:src_bytes] = w.packed[:src_bytes]
        h_dst = torch.empty(E * S, dtype=torch.uint8, pin_memory=True)
        hsp = (h_src.data_ptr() + w.comp_off[:E]).astype(np.uint64)
        hdp = (h_dst.data_ptr() + np.arange(E, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
        for _ in range(2):
            dl, st = codec.decode_batch_ptrs(dev, hsp, w.comp_len[:E], hdp, dc[:E], 0)
        assert not st.any() and (dl == S).all()
        kk = min(E, len(w.plain_head))
        assert hashlib.sha256(h_dst.numpy()[:kk * S].tobytes()).digest() == hashlib.sha256(w.plain_head[:kk].tobytes()).digest()
        barrier()
        reps = max(3, min(args.steps, 5))
        t_0 = time.perf_counter()
        for _ in range(reps):
            codec.decode_batch_ptrs(dev, hsp, w.comp_len[:E], hdp, dc[:E], 0)
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t_0) * 1e3 / reps
        e_ms = shard.max_over_ranks(e_ms, "cuda")
        link = host_link_ceiling(torch, shard, barrier, h_src, h_dst, src_bytes, E * S)
        e2e = {"value": round(shard.sum_over_ranks(E * S, "cuda") / 1e9 / (e_ms / 1e3), 3), "unit": UNIT,
               "h2d_bytes_per_step": int(w.comp_len[:E].sum()), "d2h_bytes_per_step": E * S,
               "files_per_step_per_gpu": E, "ms_per_step": round(e_ms, 3), "timer": "host wall clock around the blocking C-ABI call",
               "numa_bound_cpus": numa_cpus, "host_link": link,
               "frac_of_host_link_ceiling": round(shard.sum_over_ranks(E * S, "cuda") / 1e9 / (e_ms / 1e3) / link["ceiling_GBps"], 3) if link else None}
        del h_src, h_dst

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events recorded by the library on its stream, FZG_PROFILE)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = w.comp_bytes + w.plain_bytes
    top = max(stage_ms, key=stage_ms.get)
    top_ms = stage_ms[top] / args.steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel") == top and tj.get("files_per_gpu") == F:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(alg_bytes / 1e9 / (top_ms / 1e3), 2), "peak": peak, "unit": "GB/s",
                "frac": round(alg_bytes / 1e9 / (top_ms / 1e3) / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(top_ms, 4),
                "pipeline_frac": round(alg_bytes / 1e9 / (ms / 1e3) / peak, 4),
                "stage_ms": {k: round(v / args.steps, 4) for k, v in stage_ms.items()}}

    cpu = None
    if not args.no_cpu_baseline and world == 1 and w.ref.available:
        allc = os.cpu_count() or 1
        n_cpu = args.cpu_files or (4096 if S <= (4 << 20) else F)    # a bounded sample: about a second of all cores
        v_all, n_s = cpu_reference(w, allc, n_cpu, 1)
        v_one, n_1 = cpu_reference(w, 1, max(1, min(F, max(64 if S <= (4 << 20) else 1, n_cpu // 16))), 1)
        cpu = {"value": round(v_all, 4), "unit": UNIT, "cores": allc, "kind": "reference",
               "sample": "first %d of the %d files, copy_decode restated on libzstd %d (oracle/_ref: 131075-B reads, 8 KiB "
                         "writes), one file per thread; 1 warm + 1 timed pass" % (n_s, F, w.ref.version),
               "value_1_thread": round(v_one, 4), "sample_1_thread_files": n_1}

    # ---- BASELINE.json's second figure, "fio read MB/s via mount": the fzfs host with the GPU codec and with the
    # reference's libzstd calls, same data directory, parallel-files.fio shape (fio itself is not in this image).  N=1 only.
    mount = None
    if not args.no_mount and world == 1 and args.config == 2:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mount_bench.py"), "--jobs", "16", "--nrfiles", "625", "--fio-write",
                                "--arms", "reference,gpu:1,gpu"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)   # (a run of 250 files per job lasts 0.4 s: mostly the first windows)
            if r.returncode == 0:
                mj = json.loads(r.stdout.decode().strip().splitlines()[-1])
                mount = {"gpu_MBps": mj.get("gpu"), "gpu_1_serving_thread_MBps": mj.get("gpu_1_threads"), "reference_cpu_MBps": mj.get("reference"),
                         "gpu_write_verify_MBps": mj.get("gpu_write_verify"), "reference_cpu_write_verify_MBps": mj.get("reference_write_verify"),
                         "gpu_fio_write_and_verify_MBps": mj.get("gpu_fio_write_and_verify"), "reference_cpu_fio_write_and_verify_MBps": mj.get("reference_fio_write_and_verify"),
                         "workload": mj.get("workload"),
                         "jobs": mj.get("jobs"), "nrfiles": mj.get("nrfiles"), "filesize": mj.get("filesize"), "bs": mj.get("bs"),
                         "note": "reference = copy_decode restated on libzstd behind the same host (oracle/_ref/fzfs_ref), one serving thread as fuse-zstd / fuser; "
                                 "gpu = 8 serving threads (READs of cached files run side by side), gpu_1_serving_thread = the same host with one"}
            else:
                log("mount leg failed: " + r.stderr.decode()[-400:])
        except Exception as e:      # no /dev/fuse, no mount permission, timeout: the leg is reported as absent, the bench line stands
            log("mount leg unavailable: %r" % (e,))

    out = {"metric": METRIC, "value": round(total_plain / 1e9 / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": workload_name,
                      "files_per_gpu": F, "file_size": S, "level": args.level, "ratio": round(w.plain_bytes / w.comp_bytes, 3),
                      "compressed_bytes_per_gpu": w.comp_bytes, "sharding": "by inode (ino mod n), no collective" if args.config == 3 else "by inode, no collective",
                      "l2": "working set %.1f GB per step >> 126 MB L2, no flush needed" % (alg_bytes / 1e9)},
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mount": mount}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_b200_encode(args, rank, local_rank, world, codec, shard, threads):
    """config 5, the write path (store_to_source_file, /root/reference/src/main.rs:781-791): fzg_encode_batch over F plain files
    of S bytes resident in HBM; every step re-encodes all of them.  Round trip and ratio are checked against libzstd on a sample."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    corpus = importlib.import_module("fuse-zstd_b200.corpus")
    ref = pyoracle.Ref()
    dev = local_rank
    F, S = args.files, args.file_size
    first = ENC_FIRST + rank * F
    cap = codec.encode_bound(S)
    d_src = torch.empty(F * S, dtype=torch.uint8, device="cuda")
    d_dst = torch.zeros(F * cap, dtype=torch.uint8, device="cuda")
    chunk = max(1, (1 << 30) // S)
    head = None
    t0 = time.time()
    for c0 in range(0, F, chunk):                                    # generated and uploaded a GiB at a time
        m = min(chunk, F - c0)
        plain = corpus.json_files(first + c0, m, S, threads=threads)
        d_src[c0 * S:(c0 + m) * S] = torch.from_numpy(plain.reshape(-1)).cuda()
        if c0 == 0:
            head = plain[:min(m, 64)].copy()
    log("[rank %d] config 5 corpus: %d x %d B plain files in HBM, %.1fs" % (rank, F, S, time.time() - t0))
    sp = (d_src.data_ptr() + np.arange(F, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
    dp = (d_dst.data_ptr() + np.arange(F, dtype=np.uint64) * np.uint64(cap)).astype(np.uint64)
    sl, dc = np.full(F, S, dtype=np.uint64), np.full(F, cap, dtype=np.uint64)
    flags = codec.SRC_DEVICE | codec.DST_DEVICE | codec.PROFILE
    stream = torch.cuda.ExternalStream(codec.stream_handle(dev), device=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        dl, st = codec.encode_batch_ptrs(dev, sp, sl, dp, dc, args.level, 0, flags)
    assert not st.any(), "encode failed: statuses %s" % np.unique(st)
    out_bytes = int(dl.sum())
    # every emitted file of the sample round-trips through the reference's reader (libzstd copy_decode), byte for byte
    k = len(head)
    verify = None
    if ref.available:
        host = d_dst[:k * cap].cpu().numpy()
        l3 = 0
        for i in range(k):
            s_, plain_back = ref.copy_decode(host[i * cap:i * cap + int(dl[i])].tobytes(), S)
            assert s_ == 0 and plain_back == head[i].tobytes(), "libzstd does not reproduce file %d from the GPU encoder's output" % i
            l3 += len(ref.writer_encode(head[i].tobytes(), args.level))
        verify = {"files_round_tripped_through_libzstd": k, "bytes_vs_libzstd_same_level": round(float(dl[:k].sum()) / l3, 4),
                  "libzstd_ratio": round(k * S / l3, 3)}
    sampler = ClockSampler(gpu_index_for_nvml(local_rank)); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms, launches = {}, 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        codec.encode_batch_ptrs(dev, sp, sl, dp, dc, args.level, 0, flags)
        t = codec.last_timing(dev, encode=True)
        launches += t["launches"]
        for nm, ms_ in t["stages"].items():
            stage_ms[nm] = stage_ms.get(nm, 0.0) + ms_
    ev1.record(stream)
    barrier()
    clocks = sampler.result()
    ms = shard.max_over_ranks(ev0.elapsed_time(ev1) / args.steps, "cuda")

    # ---- e2e: pinned host buffers, H2D of the plain bytes and D2H of the frames inside the timed region
    e2e = None
    if not args.no_e2e:
        E = min(args.e2e_files or max(64, (8 << 30) // S), F)
        h_src = torch.empty(E * S, dtype=torch.uint8, pin_memory=True)
        h_src.copy_(d_src[:E * S])
        h_dst = torch.empty(E * cap, dtype=torch.uint8, pin_memory=True)
        hsp = (h_src.data_ptr() + np.arange(E, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
        hdp = (h_dst.data_ptr() + np.arange(E, dtype=np.uint64) * np.uint64(cap)).astype(np.uint64)
        for _ in range(2):
            edl, est = codec.encode_batch_ptrs(dev, hsp, sl[:E], hdp, dc[:E], args.level, 0, 0)
        assert not est.any()
        barrier()
        reps = max(3, min(args.steps, 5))
        t_0 = time.perf_counter()
        for _ in range(reps):
            codec.encode_batch_ptrs(dev, hsp, sl[:E], hdp, dc[:E], args.level, 0, 0)
        torch.cuda.synchronize()
        e_ms = shard.max_over_ranks((time.perf_counter() - t_0) * 1e3 / reps, "cuda")
        e2e = {"value": round(world * E * S / 1e9 / (e_ms / 1e3), 3), "unit": UNIT, "h2d_bytes_per_step": E * S,
               "d2h_bytes_per_step": int(edl.sum()), "files_per_step_per_gpu": E, "ms_per_step": round(e_ms, 3),
               "timer": "host wall clock around the blocking C-ABI call"}
        del h_src, h_dst
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = F * S + out_bytes                                     # uncompressed in + compressed out
    top = max(stage_ms, key=stage_ms.get)
    top_ms = stage_ms[top] / args.steps
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(alg_bytes / 1e9 / (top_ms / 1e3), 2), "peak": peak, "unit": "GB/s",
                "frac": round(alg_bytes / 1e9 / (top_ms / 1e3) / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(top_ms, 4),
                "pipeline_frac": round(alg_bytes / 1e9 / (ms / 1e3) / peak, 4),
                "stage_ms": {k_: round(v / args.steps, 4) for k_, v in stage_ms.items()}}
    cpu = None
    if not args.no_cpu_baseline and world == 1 and ref.available:
        allc = os.cpu_count() or 1
        n_cpu = min(F, args.cpu_files or max(allc * 8, (4 << 30) // S))
        plain = corpus.json_files(first, n_cpu, S, threads=allc)
        bound = ref.bound(S) + 64
        comp = np.empty((n_cpu, bound), dtype=np.uint8)
        csp = (plain.ctypes.data + np.arange(n_cpu, dtype=np.uint64) * np.uint64(S)).astype(np.uint64)
        cdp = (comp.ctypes.data + np.arange(n_cpu, dtype=np.uint64) * np.uint64(bound)).astype(np.uint64)
        csl, cdc = np.full(n_cpu, S, dtype=np.uint64), np.full(n_cpu, bound, dtype=np.uint64)
        ref.batch(2, allc, csp, csl, cdp, cdc, args.level)
        t_, ol_, st_ = ref.batch(2, allc, csp, csl, cdp, cdc, args.level)
        assert not st_.any()
        cpu = {"value": round(n_cpu * S / 1e9 / t_, 4), "unit": UNIT, "cores": allc, "kind": "reference",
               "sample": "first %d of the %d files, the reference's writer (Encoder::new(level) + pledged size + checksum, src/main.rs:781-791) "
                         "restated on libzstd %d level %d, one file per thread; 1 warm + 1 timed pass" % (n_cpu, F, ref.version, args.level),
               "ratio": round(n_cpu * S / float(ol_.sum()), 3)}
    out = {"metric": ENC_METRIC, "value": round(world * F * S / 1e9 / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": "config5: level-%d multi-frame encode (release / flush of the write path) of %d x %d B synthetic JSON files per "
                                  "GPU, FCS + XXH64 in every frame, device-resident" % (args.level, F, S),
                      "files_per_gpu": F, "file_size": S, "level": args.level, "ratio": round(F * S / out_bytes, 3),
                      "compressed_bytes_per_gpu": out_bytes, "sharding": "by inode, no collective", "verify": verify,
                      "l2": "working set %.1f GB per step >> 126 MB L2, no flush needed" % (alg_bytes / 1e9)},
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mount": None}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "b200":      # convenience: self-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
