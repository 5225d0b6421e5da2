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

--impl reference times only that CPU path (all host threads, bounded sample per step).
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
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=10000, help="files per GPU (config 2: 10 000)")
    ap.add_argument("--file-size", type=int, default=1 << 20)
    ap.add_argument("--level", type=int, default=3)
    ap.add_argument("--e2e-files", type=int, default=0, help="files per GPU in the host-buffer (e2e) leg (0: all, or half when host RAM is short)")
    ap.add_argument("--cpu-files", type=int, default=4096, help="bounded sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mount", action="store_true", help="skip the read-through-the-mount leg (tools/mount_bench.py)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- workload
class Workload:
    """F files of S bytes: plain -> level-3 frames by the reference-writer restatement (oracle/_ref, libzstd);
    packed compressed bytes stay on the host (numpy), offsets 16-byte aligned."""

    def __init__(self, first, n, size, level, threads, keep_plain=512):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        corpus = importlib.import_module("fuse-zstd_b200.corpus")
        pyoracle.build()
        self.ref = pyoracle.Ref()
        self.n, self.size = n, size
        chunk = 500
        self.comp_len = np.zeros(n, dtype=np.uint64)
        self.comp_off = np.zeros(n, dtype=np.uint64)
        parts, off = [], 0
        self.plain_head = None
        bound = (self.ref.bound(size) if self.ref.available else size + size // 128 + 512) + 64
        t_gen = t_cmp = 0.0
        for c0 in range(0, n, chunk):
            m = min(chunk, n - c0)
            t0 = time.time()
            plain = corpus.json_files(first + c0, m, size, threads=threads)
            t1 = time.time()
            comp = np.empty((m, bound), dtype=np.uint8)
            if self.ref.available:
                sp = plain.ctypes.data + np.arange(m, dtype=np.uint64) * np.uint64(size)
                dp = comp.ctypes.data + np.arange(m, dtype=np.uint64) * np.uint64(bound)
                _, ol, st = self.ref.batch(2, threads, sp.astype(np.uint64), np.full(m, size, dtype=np.uint64),
                                           dp.astype(np.uint64), np.full(m, bound, dtype=np.uint64), level)
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
            parts.append((comp, ol))
            if c0 == 0:
                k = min(keep_plain, m)
                self.plain_head = plain[:k].copy()
        self.packed = np.zeros(off + 64, dtype=np.uint8)
        i = 0
        for comp, ol in parts:
            for j in range(len(ol)):
                o, L = int(self.comp_off[i]), int(ol[j])
                self.packed[o:o + L] = comp[j, :L]
                i += 1
        del parts
        self.comp_bytes = int(self.comp_len.sum())
        self.plain_bytes = n * size
        self.gen_s, self.cmp_s = t_gen, t_cmp


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
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = min(args.cpu_files, args.files)
    w = Workload(0, n, args.file_size, args.level, threads)
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
    sample = "%d of the %d x %d B level-%d files per step, copy_decode restated on libzstd %d (oracle/_ref), %d threads" % (
        n, args.files, args.file_size, args.level, w.ref.version, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config2: batched decode of %d x %d B synthetic JSON files, zstd level %d" % (args.files, args.file_size, args.level),
                   "sample_files_per_step": n, "ratio": round(w.plain_bytes / w.comp_bytes, 3)},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, local_rank, world):
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
    F, S = args.files, args.file_size
    t0 = time.time()
    w = Workload(shard.files_for_rank(rank, F)[0], F, S, args.level, threads)
    log("[rank %d] corpus: %d files, ratio %.3f, gen %.1fs compress %.1fs (%d threads)" % (
        rank, F, w.plain_bytes / w.comp_bytes, w.gen_s, w.cmp_s, threads))

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

    # ---- e2e: pinned host buffers through the same C ABI call
    e2e = None
    if not args.no_e2e:
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
        e2e = {"value": round(world * E * S / 1e9 / (e_ms / 1e3), 3), "unit": UNIT,
               "h2d_bytes_per_step": int(w.comp_len[:E].sum()), "d2h_bytes_per_step": E * S,
               "files_per_step_per_gpu": E, "ms_per_step": round(e_ms, 3), "timer": "host wall clock around the blocking C-ABI call"}
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
        v_all, n_s = cpu_reference(w, allc, args.cpu_files, 1)
        v_one, n_1 = cpu_reference(w, 1, max(64, args.cpu_files // 16), 1)
        cpu = {"value": round(v_all, 4), "unit": UNIT, "cores": allc, "kind": "reference",
               "sample": "first %d of the %d files, copy_decode restated on libzstd %d (oracle/_ref: 131075-B reads, 8 KiB "
                         "writes), one file per thread; 1 warm + 1 timed pass" % (n_s, F, w.ref.version),
               "value_1_thread": round(v_one, 4), "sample_1_thread_files": n_1}

    # ---- BASELINE.json's second figure, "fio read MB/s via mount": the fzfs host with the GPU codec and with the
    # reference's libzstd calls, same data directory, parallel-files.fio shape (fio itself is not in this image).  N=1 only.
    mount = None
    if not args.no_mount and world == 1:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mount_bench.py"), "--jobs", "16", "--nrfiles", "125"],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
            if r.returncode == 0:
                mj = json.loads(r.stdout.decode().strip().splitlines()[-1])
                mount = {"gpu_MBps": mj.get("gpu"), "reference_cpu_MBps": mj.get("reference"),
                         "gpu_write_verify_MBps": mj.get("gpu_write_verify"), "reference_cpu_write_verify_MBps": mj.get("reference_write_verify"),
                         "workload": mj.get("workload"),
                         "jobs": mj.get("jobs"), "nrfiles": mj.get("nrfiles"), "filesize": mj.get("filesize"), "bs": mj.get("bs"),
                         "note": "one FUSE thread in both arms, as fuse-zstd; reference = copy_decode restated on libzstd (oracle/_ref/fzfs_ref)"}
            else:
                log("mount leg failed: " + r.stderr.decode()[-400:])
        except Exception as e:      # no /dev/fuse, no mount permission, timeout: the leg is reported as absent, the bench line stands
            log("mount leg unavailable: %r" % (e,))

    out = {"metric": METRIC, "value": round(world * w.plain_bytes / 1e9 / (ms / 1e3), 3), "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": "config2: batched decode of %d x %d B synthetic JSON files per GPU, zstd level %d, reference-writer "
                                  "frames (FCS + XXH64 verified), device-resident" % (F, S, args.level),
                      "files_per_gpu": F, "file_size": S, "level": args.level, "ratio": round(w.plain_bytes / w.comp_bytes, 3),
                      "compressed_bytes_per_gpu": w.comp_bytes, "sharding": "by inode, no collective",
                      "l2": "working set %.1f GB per step >> 126 MB L2, no flush needed" % (alg_bytes / 1e9)},
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "mount": mount}
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
