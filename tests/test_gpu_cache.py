"""Batch formation + decoded-file cache, through the C ABI: the mount's one-file-per-open pattern
(/root/reference/src/main.rs:451-493) served from one batched decode of a directory's .zst files."""
import hashlib
import importlib
import os
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

codec = importlib.import_module("fuse-zstd_b200.codec")


@pytest.fixture(scope="module", autouse=True)
def _init():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    codec.build()
    codec.init([0])
    yield


def _open_through_cache(path, key):
    with open(path, "rb") as src, tempfile.TemporaryFile() as tmp:
        st, size, hit = codec.cache_open(src.fileno(), tmp.fileno(), key)
        at_end = src.tell() if False else os.lseek(src.fileno(), 0, os.SEEK_CUR)
        tmp.seek(0)
        return st, tmp.read(), hit, at_end


def test_prefetch_then_open(ref, corpus):
    if not ref.available:
        pytest.skip("system libzstd absent")
    n, size = 48, 1 << 20
    plain = corpus.json_files(5550000, n, size, threads=os.cpu_count())
    with tempfile.TemporaryDirectory() as d:
        paths, keys = [], []
        for i in range(n):
            p = os.path.join(d, "f%03d.zst" % i)
            with open(p, "wb") as fh:
                fh.write(ref.writer_encode(plain[i].tobytes(), 3))
            paths.append(p); keys.append(1000 + i)
        with open(os.path.join(d, "junk.zst"), "wb") as fh:
            fh.write(b"not a zstd file")
        paths.append(os.path.join(d, "junk.zst")); keys.append(2000)
        paths.append(os.path.join(d, "missing.zst")); keys.append(2001)
        codec.cache_configure(1 << 30)
        s0 = codec.cache_stats()
        assert codec.cache_prefetch(paths, keys) == n                     # junk and missing files are skipped, not fatal
        assert codec.cache_prefetch(paths, keys) == 0                     # nothing left to do
        for i in range(n):
            st, out, hit, at_end = _open_through_cache(paths[i], keys[i])
            assert st == 0 and hit and out == plain[i].tobytes()
            assert at_end == os.path.getsize(paths[i])                    # copy_decode leaves the source at its end
        s1 = codec.cache_stats()
        assert s1["hits"] - s0["hits"] == n and s1["files"] >= n
        # the junk file fails the ordinary way (src/main.rs:467: any error -> EFAULT in the shim)
        with open(paths[n], "rb") as src, tempfile.TemporaryFile() as tmp:
            st, _, hit = codec.cache_open(src.fileno(), tmp.fileno(), keys[n])
            assert st != 0 and not hit
        # a rewritten source is never served from the cache (size / mtime stamp)
        new = corpus.json_file(99, 700000).tobytes()
        time.sleep(0.01)
        with open(paths[3], "wb") as fh:
            fh.write(ref.writer_encode(new, 3))
        st, out, hit, _ = _open_through_cache(paths[3], keys[3])
        assert st == 0 and not hit and out == new
        # explicit invalidation (store_to_source_file, rename, unlink)
        assert codec.cache_invalidate(keys[5]) == 1 and codec.cache_invalidate(keys[5]) == 0
        st, out, hit, _ = _open_through_cache(paths[5], keys[5])
        assert st == 0 and not hit and out == plain[5].tobytes()
        # a new capacity starts empty; a prefetch never takes more than the cache holds; slabs are reused oldest-first
        codec.cache_configure(8 << 20)
        assert codec.cache_stats()["files"] == 0
        assert codec.cache_prefetch(paths[:n], keys[:n]) == 8                # 8 x 1 MiB fit
        for i in range(8):
            assert _open_through_cache(paths[i], keys[i])[2]
        assert codec.cache_prefetch(paths[8:20], keys[8:20]) == 8             # the slab is reused: the first eight leave
        st, out, hit, _ = _open_through_cache(paths[0], keys[0])
        assert st == 0 and not hit and out == plain[0].tobytes()
        st, out, hit, _ = _open_through_cache(paths[9], keys[9])
        assert st == 0 and hit and out == plain[9].tobytes()
        codec.cache_configure(0)
        assert codec.cache_stats()["files"] == 0 and codec.cache_prefetch(paths, keys) == 0
        codec.cache_configure(1 << 30)


def test_background_prefetch(ref, corpus):
    if not ref.available:
        pytest.skip("system libzstd absent")
    n, size = 16, 1 << 19
    plain = corpus.json_files(5560000, n, size, threads=os.cpu_count())
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(n):
            p = os.path.join(d, "g%02d.zst" % i)
            with open(p, "wb") as fh:
                fh.write(ref.writer_encode(plain[i].tobytes(), 3))
            paths.append(p)
        keys = list(range(7000, 7000 + n))
        codec.cache_prefetch(paths, keys, background=True)
        for _ in range(400):
            if codec.cache_stats()["files"] >= n:
                break
            time.sleep(0.01)
        for i in range(n):                                                 # correct whether or not the batch has landed
            st, out, hit, _ = _open_through_cache(paths[i], keys[i])
            assert st == 0 and hashlib.sha256(out).digest() == hashlib.sha256(plain[i].tobytes()).digest()
        for k in keys:
            codec.cache_invalidate(k)


def test_directory_larger_than_a_slab(ref, corpus):
    """slabs reserved at mount time hold 256 MiB each: a directory of 300 MiB is decoded as two batches, not dropped"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    n, size = 300, 1 << 20
    plain = corpus.json_files(5570000, 4, size, threads=os.cpu_count())
    comps = [ref.writer_encode(plain[i].tobytes(), 3) for i in range(4)]
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i in range(n):
            p = os.path.join(d, "h%03d.zst" % i)
            with open(p, "wb") as fh:
                fh.write(comps[i % 4])
            paths.append(p)
        keys = list(range(9000, 9000 + n))
        codec.cache_configure(1 << 30)
        assert codec.cache_reserve() >= 2
        assert codec.cache_prefetch(paths, keys) == n
        for i in (0, 1, 254, 255, 256, 257, n - 1):
            st, out, hit, _ = _open_through_cache(paths[i], keys[i])
            assert st == 0 and hit and hashlib.sha256(out).digest() == hashlib.sha256(plain[i % 4].tobytes()).digest()
        for k in keys:
            codec.cache_invalidate(k)
