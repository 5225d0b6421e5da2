"""Encode throughput / ratio probe (config 5 shape: N x 4 MiB JSON files, device-resident)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
codec = importlib.import_module("fuse-zstd_b200.codec"); corpus = importlib.import_module("fuse-zstd_b200.corpus")
import pyoracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 3      # 1-2: single-table matcher, else the two-table one
size = 4 << 20
codec.init([0])
plain = corpus.json_files(7000000, n, size, threads=os.cpu_count())
d_src = torch.from_numpy(plain).cuda()
cap = codec.encode_bound(size)
d_dst = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
sp = [d_src.data_ptr() + i * size for i in range(n)]; dp = [d_dst.data_ptr() + i * cap for i in range(n)]
fl = codec.SRC_DEVICE | codec.DST_DEVICE | codec.PROFILE
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dl, st = codec.encode_batch_ptrs(0, sp, [size] * n, dp, [cap] * n, level, 0, fl)
    dt = time.perf_counter() - t0
    t = codec.last_timing(0, encode=True)
    print("level %d: encode %d x 4 MiB: %.1f ms wall, gpu %.1f ms -> %.1f GB/s in, ratio %.3f, stages %s" % (level, n, dt * 1e3, t["total_ms"], n * size / 1e9 / (t["total_ms"] / 1e3), n * size / dl.sum(), {k: round(v, 2) for k, v in list(t["stages"].items())[:5]}), file=sys.stderr)
assert not st.any()
R = pyoracle.Ref()
if R.available:
    host = d_dst.cpu().numpy(); k = min(n, 32); ok = 0; l3 = 0
    for i in range(k):
        s, out = R.copy_decode(host[i * cap:i * cap + int(dl[i])].tobytes(), size)
        ok += (s == 0 and out == plain[i].tobytes()); l3 += len(R.writer_encode(plain[i].tobytes(), 3))
    print("libzstd round trip: %d / %d files ok; bytes vs libzstd L3: x%.3f" % (ok, k, dl[:k].sum() / l3), file=sys.stderr)
    k2 = min(n, 64)
    sp = np.array([plain[i].ctypes.data for i in range(k2)], dtype=np.uint64); sl = np.full(k2, size, dtype=np.uint64)
    bound = R.bound(size) + 64; comp = np.empty((k2, bound), dtype=np.uint8)
    dp2 = np.array([comp[i].ctypes.data for i in range(k2)], dtype=np.uint64); dc = np.full(k2, bound, dtype=np.uint64)
    R.batch(2, os.cpu_count(), sp, sl, dp2, dc, 3)
    t, ol, st2 = R.batch(2, os.cpu_count(), sp, sl, dp2, dc, 3)
    print("CPU reference writer (libzstd %d level 3, %d threads): %.2f GB/s in, ratio %.3f" % (R.version, os.cpu_count(), k2 * size / 1e9 / t, k2 * size / ol.sum()), file=sys.stderr)
