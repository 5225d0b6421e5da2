"""Config 4 in small: a few LARGE single-frame files (level 3, windowLog 23 = 8 MiB window, matches reach across dozens of
blocks), device-resident batched decode.  With so few frames each one is executed by a CTA of 32 warps with a checksum
CTA beside it (k_execute_cta, DESIGN.md section 2 "Few frames"); FZG_EXEC_W=1 gives the one-warp-per-frame figure.
usage: large_file_probe.py [files] [MiB per file] [host]      (host: also through pinned HOST buffers, the e2e path)"""
import importlib, os, sys, time, hashlib
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
codec = importlib.import_module("fuse-zstd_b200.codec"); corpus = importlib.import_module("fuse-zstd_b200.corpus")
import pyoracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = (int(sys.argv[2]) if len(sys.argv) > 2 else 128) << 20
R = pyoracle.Ref(); assert R.available
codec.init([0])
t0 = time.time()
plain = corpus.json_files(9000000, n, size, threads=os.cpu_count())
with ThreadPoolExecutor(os.cpu_count()) as ex:
    blobs = list(ex.map(lambda i: R.writer_encode(plain[i].tobytes(), 3, window_log=23), range(n)))
print("corpus: %d x %d MiB, ratio %.3f, %.1f s" % (n, size >> 20, n * size / sum(map(len, blobs)), time.time() - t0), file=sys.stderr)
off = np.zeros(n, dtype=np.int64); tot = 0
for i, b in enumerate(blobs): off[i] = tot; tot += (len(b) + 15 & ~15) + 16
packed = np.zeros(tot + 64, dtype=np.uint8)
for i, b in enumerate(blobs): packed[off[i]:off[i] + len(b)] = np.frombuffer(b, dtype=np.uint8)
d_src = torch.from_numpy(packed).cuda(); d_dst = torch.zeros(n * size, dtype=torch.uint8, device="cuda")
sp = (d_src.data_ptr() + off).astype(np.uint64); dp = (d_dst.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size)).astype(np.uint64)
sl = np.array([len(b) for b in blobs], dtype=np.uint64); dc = np.full(n, size, dtype=np.uint64)
fl = codec.SRC_DEVICE | codec.DST_DEVICE | codec.PROFILE
for it in range(3):
    dl, st = codec.decode_batch_ptrs(0, sp, sl, dp, dc, fl)
    t = codec.last_timing(0)
    print("decode %d x %d MiB (windowLog 23): gpu %.1f ms -> %.1f GB/s out, stages %s" % (n, size >> 20, t["total_ms"], n * size / 1e9 / (t["total_ms"] / 1e3), {k: round(v, 1) for k, v in t["stages"].items() if v > 0.5}), file=sys.stderr)
assert not st.any() and (dl == size).all()
k = min(n, 2)
assert hashlib.sha256(d_dst[:k * size].cpu().numpy().tobytes()).digest() == hashlib.sha256(plain[:k].tobytes()).digest()
print("bytes verified", file=sys.stderr)
if len(sys.argv) > 3 and sys.argv[3] == "host":
    h_src = torch.from_numpy(packed).pin_memory(); h_dst = torch.empty(n * size, dtype=torch.uint8, pin_memory=True)
    sp = (h_src.data_ptr() + off).astype(np.uint64); dp = (h_dst.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(size)).astype(np.uint64)
    for it in range(3):
        t0 = time.perf_counter(); dl, st = codec.decode_batch_ptrs(0, sp, sl, dp, dc, 0); dt = time.perf_counter() - t0
        print("host buffers: %.1f ms -> %.1f GB/s e2e" % (dt * 1e3, n * size / 1e9 / dt), file=sys.stderr)
    assert not st.any() and hashlib.sha256(h_dst[:k * size].numpy().tobytes()).digest() == hashlib.sha256(plain[:k].tobytes()).digest()
