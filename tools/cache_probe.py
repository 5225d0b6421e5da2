"""The mount's access pattern: N files of a directory opened one after the other (fuse-zstd: one decode per open, one
thread).  (a) fzg_decode_fd per file; (b) one fzg_cache_prefetch of the directory, then fzg_cache_open per file;
(c) the reference's CPU path per file (copy_decode restated on libzstd, one thread).   usage: cache_probe.py [files] [KiB]"""
import importlib, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
codec = importlib.import_module("fuse-zstd_b200.codec"); corpus = importlib.import_module("fuse-zstd_b200.corpus")
import pyoracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
size = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024) << 10
R = pyoracle.Ref(); assert R.available
codec.init([0])
codec.cache_configure(1 << 30); codec.cache_reserve()
plain = corpus.json_files(8800000, n, size, threads=os.cpu_count())
with tempfile.TemporaryDirectory() as d:
    paths = []
    for i in range(n):
        p = os.path.join(d, "f%05d.zst" % i)
        with open(p, "wb") as fh: fh.write(R.writer_encode(plain[i].tobytes(), 3))
        paths.append(p)
    keys = list(range(n))
    def each(fn):
        t0 = time.perf_counter()
        for i in range(n):
            with open(paths[i], "rb") as src, tempfile.TemporaryFile() as tmp: fn(src, tmp, i)
        return time.perf_counter() - t0
    each(lambda s, t, i: codec.decode_fd(s.fileno(), t.fileno(), i) if i < 4 else None)      # warm up
    ta = each(lambda s, t, i: codec.decode_fd(s.fileno(), t.fileno(), i))
    t0 = time.perf_counter(); codec.cache_prefetch(paths, [k + (1 << 20) for k in keys]); tcold = time.perf_counter() - t0    # first batch of the process: staging buffers are allocated
    for k in keys: codec.cache_invalidate(k + (1 << 20))
    t0 = time.perf_counter(); added = codec.cache_prefetch(paths, keys); tp = time.perf_counter() - t0
    tb = each(lambda s, t, i: codec.cache_open(s.fileno(), t.fileno(), i))
    def cpu(s, t, i):
        st, out = R.copy_decode(s.read(), size); t.write(out)
    tc = each(cpu)
    gb = n * size / 1e9
    print("%d x %d KiB opened one by one: fzg_decode_fd %.1f ms/file (%.2f GB/s); prefetch of the directory %.1f ms (%d files; first batch of the process %.1f ms) + fzg_cache_open %.2f ms/file -> %.2f GB/s overall; libzstd on one thread %.2f ms/file (%.2f GB/s)"
          % (n, size >> 10, ta / n * 1e3, gb / ta, tp * 1e3, added, tcold * 1e3, tb / n * 1e3, gb / (tp + tb), tc / n * 1e3, gb / tc), file=sys.stderr)
