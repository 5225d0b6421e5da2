"""Execute stage, warps per frame (FZG_EXEC_W = 1: k_execute, 8 / 32: k_execute_cta) against batch shape: the same
level-3 reference-writer corpus decoded device-resident as batches of 1 .. N files of 1 MiB, and a few large
windowLog-23 frames (config 4 in small).  Every result is compared with the plain bytes.
usage: exec_width_probe.py [max files of 1 MiB] [large files] [MiB per large file] [batch sizes, comma] [widths, comma]"""
import importlib, os, sys, time, hashlib
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
codec = importlib.import_module("fuse-zstd_b200.codec"); corpus = importlib.import_module("fuse-zstd_b200.corpus")
import pyoracle
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nlarge = int(sys.argv[2]) if len(sys.argv) > 2 else 16
large_mib = int(sys.argv[3]) if len(sys.argv) > 3 else 64
sizes_arg = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else None
widths_arg = tuple(sys.argv[5].split(",")) if len(sys.argv) > 5 else None      # FZG_EXEC_W values: 1, 8, 32, s, t128, t256, t1024 ...
R = pyoracle.Ref(); assert R.available
codec.init([0])
thr = os.cpu_count() or 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def pack(blobs):
    off = np.zeros(len(blobs), dtype=np.int64); tot = 0
    for i, b in enumerate(blobs): off[i] = tot; tot += (len(b) + 15 & ~15) + 16
    packed = np.zeros(tot + 64, dtype=np.uint8)
    for i, b in enumerate(blobs): packed[off[i]:off[i] + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return packed, off


def run(tag, d_src, off, lens, d_dst, size, n, digest, widths, reps=5):
    reps = int(os.environ.get("PROBE_REPS", reps))      # 1 under ncu
    sp = (d_src.data_ptr() + off[:n]).astype(np.uint64)
    dp = (d_dst.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size)).astype(np.uint64)
    sl = np.asarray(lens[:n], dtype=np.uint64); dc = np.full(n, size, dtype=np.uint64)
    fl = codec.SRC_DEVICE | codec.DST_DEVICE | codec.PROFILE
    for w in widths:
        os.environ["FZG_EXEC_W"] = str(w)
        d_dst[:n * size].zero_()
        best = None
        for it in range(reps):
            dl, st = codec.decode_batch_ptrs(0, sp, sl, dp, dc, fl)
            t = codec.last_timing(0)
            if best is None or t["total_ms"] < best["total_ms"]: best = t
        assert not st.any() and (dl == size).all(), (tag, w)
        k = min(n, 4)
        ok = hashlib.sha256(d_dst[:k * size].cpu().numpy().tobytes()).digest() == digest(k)
        log("%-22s W=%-5s total %8.3f ms  execute %8.3f ms  %7.1f GB/s out  %s  %s" % (
            tag, w, best["total_ms"], best["stages"]["execute"], n * size / 1e9 / (best["total_ms"] / 1e3), "ok" if ok else "MISMATCH",
            " ".join("%s %.2f" % (k[:3], v) for k, v in best["stages"].items() if v >= 0.05 and k != "execute")))
        assert ok, (tag, w)
    os.environ.pop("FZG_EXEC_W", None)


# ---- 1 MiB files (config 2's shape)
size = 1 << 20
t0 = time.time()
plain = corpus.json_files(0, nmax, size, threads=thr)
cap = R.bound(size) + 64
comp = np.empty((nmax, cap), dtype=np.uint8)
sp = np.array([plain[i].ctypes.data for i in range(nmax)], dtype=np.uint64); sl = np.full(nmax, size, dtype=np.uint64)
dp = np.array([comp[i].ctypes.data for i in range(nmax)], dtype=np.uint64); dc = np.full(nmax, cap, dtype=np.uint64)
_, ol, st = R.batch(2, thr, sp, sl, dp, dc, 3)
assert not st.any()
blobs = [comp[i, :int(ol[i])] for i in range(nmax)]
log("corpus: %d x 1 MiB, ratio %.3f, %.1f s" % (nmax, nmax * size / float(ol.sum()), time.time() - t0))
packed, off = pack(blobs)
d_src = torch.from_numpy(packed).cuda(); d_dst = torch.zeros(nmax * size, dtype=torch.uint8, device="cuda")
dig = lambda k: hashlib.sha256(plain[:k].tobytes()).digest()
for n in sizes_arg or (1, 8, 64, 148, 296, 592, 1184, 2368, 4736, nmax):
    if n > nmax: continue
    run("%d x 1 MiB" % n, d_src, off, ol, d_dst, size, n, dig, widths_arg or ((1, 8, 32) if n <= 2368 else (1, 8)), reps=5 if n < 4000 else 3)
del d_src, d_dst, packed, comp, plain
torch.cuda.empty_cache()

# ---- large single frames, windowLog 23 (config 4 in small)
if nlarge:
    size = large_mib << 20
    t0 = time.time()
    plain = corpus.json_files(9000000, nlarge, size, threads=thr)
    with ThreadPoolExecutor(thr) as ex:
        blobs = list(ex.map(lambda i: R.writer_encode(plain[i].tobytes(), 3, window_log=23), range(nlarge)))
    log("corpus: %d x %d MiB windowLog 23, ratio %.3f, %.1f s" % (nlarge, large_mib, nlarge * size / sum(map(len, blobs)), time.time() - t0))
    packed, off = pack(blobs)
    d_src = torch.from_numpy(packed).cuda(); d_dst = torch.zeros(nlarge * size, dtype=torch.uint8, device="cuda")
    dig = lambda k: hashlib.sha256(plain[:k].tobytes()).digest()
    run("%d x %d MiB wlog23" % (nlarge, large_mib), d_src, off, [len(b) for b in blobs], d_dst, size, nlarge, dig, widths_arg or (1, 8, 32), reps=2)
