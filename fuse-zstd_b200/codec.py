"""ctypes bindings of libfzgpu.so (include/fzgpu.h).  No torch types cross this boundary."""
import ctypes as C
import errno
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("FZG_LIB") or os.path.join(HERE, "libfzgpu.so")   # FZG_LIB: a differently tuned build (kernel experiments)

OK, E_MAGIC, E_TRUNCATED, E_UNSUPPORTED, E_CORRUPT, E_DSTSIZE, E_CHECKSUM, E_FCS = range(8)
SRC_DEVICE, DST_DEVICE, NO_VERIFY_CHECKSUM, PROFILE, SEEK_TABLE = 1, 2, 4, 8, 16

EXPORTS = ["fzg_init", "fzg_shutdown", "fzg_device_count", "fzg_decode_fd", "fzg_encode_fd", "fzg_decode_batch",
           "fzg_encode_batch", "fzg_encode_bound", "fzg_frame_info", "fzg_strerror", "fzg_last_timing", "fzg_streamed_copies", "fzg_reserve_staging",
           "fzg_stage_name", "fzg_stream", "fzg_decode_range", "fzg_decode_range_fd", "fzg_seek_footer",
           "fzg_cache_configure", "fzg_cache_reserve", "fzg_cache_prefetch", "fzg_cache_prefetch_async", "fzg_cache_open", "fzg_cache_invalidate",
           "fzg_cache_stats", "fzg_cache_drain", "fzg_cache_view", "fzg_cache_unview", "fzg_cache_wait", "fzg_cache_pending"]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("kernel_ms", C.c_float * 16), ("launches", C.c_int),
                ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64)]


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); cross-compiles without a GPU."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "csrc"), "all"] + (["-B"] if force else []))


_lib = None


def lib():
    """The loaded C ABI.  Raises if the CUDA library has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError("libfzgpu.so is not built (run __graft_entry__.build()); fuse-zstd_b200 has no CPU path")
        L = C.CDLL(SO)
        L.fzg_init.restype = C.c_int; L.fzg_init.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.fzg_shutdown.restype = None
        L.fzg_device_count.restype = C.c_int
        L.fzg_decode_fd.restype = C.c_int
        L.fzg_decode_fd.argtypes = [C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
        L.fzg_encode_fd.restype = C.c_int
        L.fzg_encode_fd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.fzg_decode_batch.restype = C.c_int
        L.fzg_decode_batch.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        L.fzg_encode_batch.restype = C.c_int
        L.fzg_encode_batch.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        L.fzg_encode_bound.restype = C.c_size_t; L.fzg_encode_bound.argtypes = [C.c_size_t, C.c_size_t]
        L.fzg_frame_info.restype = C.c_int
        L.fzg_frame_info.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.fzg_strerror.restype = C.c_char_p; L.fzg_strerror.argtypes = [C.c_int]
        L.fzg_last_timing.restype = C.c_int; L.fzg_last_timing.argtypes = [C.c_int, C.POINTER(Timing)]
        L.fzg_streamed_copies.restype = C.c_uint64; L.fzg_streamed_copies.argtypes = [C.c_int]
        L.fzg_stage_name.restype = C.c_char_p; L.fzg_stage_name.argtypes = [C.c_int]
        L.fzg_stream.restype = C.c_void_p; L.fzg_stream.argtypes = [C.c_int]
        if not hasattr(L, "fzg_decode_range"):            # an older tuned build loaded through FZG_LIB (kernel experiments)
            _lib = L
            return _lib
        L.fzg_decode_range.restype = C.c_int
        L.fzg_decode_range.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        L.fzg_decode_range_fd.restype = C.c_int
        L.fzg_decode_range_fd.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        L.fzg_cache_configure.restype = C.c_int; L.fzg_cache_configure.argtypes = [C.c_size_t]
        L.fzg_cache_reserve.restype = C.c_int; L.fzg_cache_reserve.argtypes = []
        L.fzg_cache_prefetch.restype = C.c_int
        L.fzg_cache_prefetch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_size_t]
        L.fzg_cache_prefetch_async.restype = C.c_int
        L.fzg_cache_prefetch_async.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_size_t]
        L.fzg_cache_open.restype = C.c_int
        L.fzg_cache_open.argtypes = [C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.fzg_cache_invalidate.restype = C.c_int; L.fzg_cache_invalidate.argtypes = [C.c_uint64]
        L.fzg_cache_stats.restype = None
        L.fzg_cache_stats.argtypes = [C.POINTER(C.c_uint64)] * 4
        L.fzg_seek_footer.restype = C.c_int
        L.fzg_seek_footer.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _check(rc, what):
    if rc < 0:
        raise OSError(-rc, "%s: %s" % (what, os.strerror(-rc)))
    return rc


def init(devices=None):
    if devices is None:
        return _check(lib().fzg_init(None, 0), "fzg_init")
    arr = (C.c_int * len(devices))(*devices)
    return _check(lib().fzg_init(arr, len(devices)), "fzg_init")


def shutdown():
    lib().fzg_shutdown()


def strerror(code):
    return lib().fzg_strerror(code).decode()


def frame_info(data):
    """-> (status, content_size or None, compressed_size).  Host-side header walk, no GPU."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    cs, zs = C.c_uint64(0), C.c_uint64(0)
    st = lib().fzg_frame_info(a.ctypes.data if a.size else None, a.size, C.byref(cs), C.byref(zs))
    return st, (None if cs.value == 2**64 - 1 else cs.value), zs.value


def _u64(seq):
    return np.ascontiguousarray(np.asarray(seq, dtype=np.uint64))


def decode_batch_ptrs(device, src_ptrs, src_lens, dst_ptrs, dst_caps, flags=0):
    """Raw form: arrays of pointers/sizes (host or device memory per `flags`) -> (dst_len[], status[])."""
    sp, sl, dp, dc = _u64(src_ptrs), _u64(src_lens), _u64(dst_ptrs), _u64(dst_caps)
    n = len(sp)
    dl = np.zeros(n, dtype=np.uint64); st = np.zeros(n, dtype=np.int32)
    _check(lib().fzg_decode_batch(device, n, sp.ctypes.data, sl.ctypes.data, dp.ctypes.data, dc.ctypes.data,
                                  dl.ctypes.data, st.ctypes.data, flags), "fzg_decode_batch")
    return dl, st


def encode_batch_ptrs(device, src_ptrs, src_lens, dst_ptrs, dst_caps, level=3, chunk_size=0, flags=0):
    sp, sl, dp, dc = _u64(src_ptrs), _u64(src_lens), _u64(dst_ptrs), _u64(dst_caps)
    n = len(sp)
    dl = np.zeros(n, dtype=np.uint64); st = np.zeros(n, dtype=np.int32)
    _check(lib().fzg_encode_batch(device, n, sp.ctypes.data, sl.ctypes.data, dp.ctypes.data, dc.ctypes.data,
                                  dl.ctypes.data, st.ctypes.data, level, chunk_size, flags), "fzg_encode_batch")
    return dl, st


def decode_batch(blobs, caps=None, device=0, flags=0):
    """Host convenience: list of bytes-like .zst files -> list of (status, bytes)."""
    arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in blobs]
    if caps is None:
        caps = []
        for a in arrs:
            st, cs, _ = frame_info(a)
            caps.append(cs if (st == 0 and cs is not None) else max(1 << 16, 64 * a.size))
    outs = [np.empty(max(int(c), 1), dtype=np.uint8) for c in caps]
    dl, st = decode_batch_ptrs(device, [a.ctypes.data if a.size else 0 for a in arrs], [a.size for a in arrs],
                               [o.ctypes.data for o in outs], [int(c) for c in caps], flags & ~(SRC_DEVICE | DST_DEVICE))
    return [(int(s), o[:int(l)].tobytes()) for s, l, o in zip(st, dl, outs)]


def encode_bound(n, chunk_size=0):
    return lib().fzg_encode_bound(n, chunk_size)


def encode_batch(blobs, level=3, chunk_size=0, device=0, flags=0):
    """Host convenience: list of bytes-like plain files -> list of (status, zstd bytes)."""
    arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in blobs]
    caps = [encode_bound(a.size, chunk_size) for a in arrs]
    outs = [np.empty(max(c, 1), dtype=np.uint8) for c in caps]
    dl, st = encode_batch_ptrs(device, [a.ctypes.data if a.size else 0 for a in arrs], [a.size for a in arrs],
                               [o.ctypes.data for o in outs], caps, level, chunk_size, flags & ~(SRC_DEVICE | DST_DEVICE))
    return [(int(s), o[:int(l)].tobytes()) for s, l, o in zip(st, dl, outs)]


def last_timing(device=0, encode=False):
    """Stage times of the last call on `device` (FZG_PROFILE); encode=True names the stages of fzg_encode_batch."""
    t = Timing()
    _check(lib().fzg_last_timing(device, C.byref(t)), "fzg_last_timing")
    stages = {}
    for k in range(16):
        nm = lib().fzg_stage_name(k + (16 if encode else 0)).decode()
        if nm:
            stages[nm] = t.kernel_ms[k]
    return dict(total_ms=t.total_ms, launches=t.launches, bytes_in=t.bytes_in, bytes_out=t.bytes_out, stages=stages)


def streamed_copies(device=0):
    """device -> host copies queued behind a running execute stage since init (few large frames into host buffers)"""
    return int(lib().fzg_streamed_copies(device))


def stream_handle(device=0):
    return lib().fzg_stream(device)


def decode_fd(src_fd, dst_fd, shard_key=0):
    """fzg_decode_fd -> plain size; raises OSError(EFAULT) on any codec failure (src/main.rs:467)."""
    out = C.c_uint64(0)
    rc = lib().fzg_decode_fd(src_fd, dst_fd, shard_key, C.byref(out))
    if rc < 0:
        raise OSError(-rc, os.strerror(-rc))
    if rc > 0:
        raise OSError(errno.EFAULT, "zstd decode failed: " + strerror(rc))
    return out.value


def encode_fd(src_fd, dst_fd, level, src_size, shard_key=0):
    """fzg_encode_fd -> compressed size; failures surface as EIO (src/errors.rs:4-10)."""
    out = C.c_uint64(0)
    rc = lib().fzg_encode_fd(src_fd, dst_fd, level, src_size, shard_key, C.byref(out))
    if rc != 0:
        raise OSError(-rc if rc < 0 else errno.EIO, "zstd encode failed")
    return out.value


def seek_footer(tail, file_size):
    """Host-only: (rc, n_frames, table_bytes) of the seek table whose footer ends `tail` (the last >= 9 bytes of a file)."""
    a = np.frombuffer(tail, dtype=np.uint8)
    nf, tb = C.c_uint32(0), C.c_uint64(0)
    rc = lib().fzg_seek_footer(a.ctypes.data if a.size else None, a.size, file_size, C.byref(nf), C.byref(tb))
    return rc, nf.value, tb.value


def decode_range(data, offset, size, device=0):
    """Plain bytes [offset, offset + size) of a .zst image that ends in a seek table; only the frames touched are decoded.
    -> (rc, bytes); rc == -errno.ENOENT when there is no seek table."""
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(max(size, 1), dtype=np.uint8); got = C.c_size_t(0)
    rc = lib().fzg_decode_range(device, a.ctypes.data, a.size, offset, size, out.ctypes.data, C.byref(got))
    return rc, out[:got.value].tobytes()


def decode_range_fd(fd, offset, size, shard_key=0):
    out = np.empty(max(size, 1), dtype=np.uint8); got = C.c_size_t(0)
    rc = lib().fzg_decode_range_fd(fd, shard_key, offset, size, out.ctypes.data, C.byref(got))
    return rc, out[:got.value].tobytes()


# ---- batch formation + decoded-file cache
def cache_configure(capacity_bytes):
    return _check(lib().fzg_cache_configure(capacity_bytes), "fzg_cache_configure")


def cache_reserve():
    return _check(lib().fzg_cache_reserve(), "fzg_cache_reserve")


def cache_prefetch(paths, keys, device=0, background=False):
    """Decode the listed .zst files as one batch and keep the plain bytes, keyed by inode -> files added."""
    n = len(paths)
    p = (C.c_char_p * n)(*[os.fsencode(x) for x in paths]); k = (C.c_uint64 * n)(*keys)
    f = lib().fzg_cache_prefetch_async if background else lib().fzg_cache_prefetch
    return _check(f(device, p, k, n), "fzg_cache_prefetch")


def cache_open(src_fd, dst_fd, key):
    """open_wrapper's codec call with the cache in front -> (status, plain size, hit)."""
    size, hit = C.c_uint64(0), C.c_int(0)
    rc = lib().fzg_cache_open(src_fd, dst_fd, key, C.byref(size), C.byref(hit))
    if rc < 0:
        raise OSError(-rc, "fzg_cache_open: %s" % os.strerror(-rc))
    return rc, size.value, bool(hit.value)


def cache_invalidate(key):
    return lib().fzg_cache_invalidate(key)


def cache_stats():
    v = [C.c_uint64(0) for _ in range(4)]
    lib().fzg_cache_stats(*[C.byref(x) for x in v])
    return dict(hits=v[0].value, misses=v[1].value, bytes=v[2].value, files=v[3].value)
