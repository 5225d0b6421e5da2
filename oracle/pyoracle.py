"""ctypes bindings for the TEST-ONLY checkers in oracle/.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

  Oracle  -- the plain-C RFC 8878 restatement (oracle/zstd_oracle.c), restating what
             /root/reference/src/main.rs:463-467 computes through libzstd.
  Ref     -- oracle/_ref/libfzref.so: the reference's two zstd-rs call sites
             (src/main.rs:463-467, :781-791) replayed against the system libzstd.so.1.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfzoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfzref.so")

STATUS_NAMES = {0: "OK", 1: "MAGIC", 2: "TRUNCATED", 3: "UNSUPPORTED", 4: "CORRUPT", 5: "DSTSIZE",
                6: "CHECKSUM", 7: "FCS"}


def build(force=False):
    """Compile the checkers (gcc only).  Building the checker is not using it."""
    if force or not (os.path.exists(ORACLE_SO) and os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))


class _Seq(C.Structure):
    _fields_ = [("lit_len", C.c_uint32), ("match_len", C.c_uint32), ("offset", C.c_uint32), ("of_value", C.c_uint32)]


class _Trace(C.Structure):
    _fields_ = [("literals", C.c_void_p), ("literals_cap", C.c_size_t), ("literals_len", C.c_size_t),
                ("seqs", C.c_void_p), ("seqs_cap", C.c_size_t), ("seqs_len", C.c_size_t),
                ("block_nseq", C.c_void_p), ("blocks_cap", C.c_size_t), ("blocks_len", C.c_size_t),
                ("block_litsize", C.c_void_p), ("block_type", C.c_void_p),
                ("block_littype", C.c_void_p), ("block_modes", C.c_void_p)]


def _buf(b):
    """bytes / bytearray / numpy uint8 -> (pointer, length, keepalive)"""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, (bytes, bytearray)) else np.frombuffer(b, dtype=np.uint8)
    return (a.ctypes.data if a.size else 0), a.size, a


class Oracle:
    def __init__(self):
        build()
        L = C.CDLL(ORACLE_SO)
        L.fzo_decode.restype = C.c_int
        L.fzo_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.fzo_decode_trace.restype = C.c_int
        L.fzo_decode_trace.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(_Trace)]
        L.fzo_frame_info.restype = C.c_int
        L.fzo_frame_info.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.fzo_xxh64.restype = C.c_uint64
        L.fzo_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.fzo_decode_batch.restype = C.c_double
        L.fzo_decode_batch.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L = L

    def decode(self, src, cap=None):
        """-> (status, bytes).  cap defaults to frame_info's content size (or 64x input)."""
        p, n, keep = _buf(src)
        if cap is None:
            st, size, _ = self.frame_info(src)
            cap = size if (st == 0 and size != 2**64 - 1) else max(1 << 16, 64 * n)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        got = C.c_size_t(0)
        st = self.L.fzo_decode(p, n, out.ctypes.data, cap, C.byref(got))
        return st, out[:got.value].tobytes()

    def decode_trace(self, src, cap, max_seqs=1 << 22, max_blocks=1 << 16):
        p, n, keep = _buf(src)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        lits = np.empty(max(cap, 1), dtype=np.uint8)
        seqs = np.empty((max_seqs, 4), dtype=np.uint32)
        bn = np.empty(max_blocks, dtype=np.uint32); bl = np.empty(max_blocks, dtype=np.uint32)
        bt = np.empty(max_blocks, dtype=np.uint8); blt = np.zeros(max_blocks, dtype=np.uint8)
        bm = np.zeros(max_blocks, dtype=np.uint8)
        tr = _Trace(lits.ctypes.data, lits.size, 0, seqs.ctypes.data, max_seqs, 0,
                    bn.ctypes.data, max_blocks, 0, bl.ctypes.data, bt.ctypes.data,
                    blt.ctypes.data, bm.ctypes.data)
        got = C.c_size_t(0)
        st = self.L.fzo_decode_trace(p, n, out.ctypes.data, cap, C.byref(got), C.byref(tr))
        nb = min(tr.blocks_len, max_blocks)
        return dict(status=st, out=out[:got.value].tobytes(), literals=lits[:min(tr.literals_len, lits.size)].copy(),
                    seqs=seqs[:min(tr.seqs_len, max_seqs)].copy(), block_nseq=bn[:nb].copy(),
                    block_litsize=bl[:nb].copy(), block_type=bt[:nb].copy(),
                    block_littype=blt[:nb].copy(), block_modes=bm[:nb].copy())

    def frame_info(self, src):
        p, n, keep = _buf(src)
        size = C.c_uint64(0); frames = C.c_uint64(0)
        st = self.L.fzo_frame_info(p, n, C.byref(size), C.byref(frames))
        return st, size.value, frames.value

    def xxh64(self, data, seed=0):
        p, n, keep = _buf(data)
        return self.L.fzo_xxh64(p, n, seed)

    def decode_batch(self, threads, src_ptrs, src_lens, dst_ptrs, dst_caps):
        """arrays of uint64 pointers / sizes -> (seconds, out_len[], status[])"""
        n = len(src_ptrs)
        out_len = np.zeros(n, dtype=np.uint64); status = np.zeros(n, dtype=np.int32)
        t = self.L.fzo_decode_batch(threads, n, src_ptrs.ctypes.data, src_lens.ctypes.data, dst_ptrs.ctypes.data,
                                    dst_caps.ctypes.data, out_len.ctypes.data, status.ctypes.data)
        return t, out_len, status


class Ref:
    """The reference's call sites replayed on the system libzstd (if present)."""

    def __init__(self):
        build()
        L = C.CDLL(REF_SO)
        L.fzr_available.restype = C.c_int
        L.fzr_version.restype = C.c_uint
        L.fzr_compress_bound.restype = C.c_size_t
        L.fzr_compress_bound.argtypes = [C.c_size_t]
        for name in ("fzr_copy_decode", "fzr_decode_oneshot"):
            f = getattr(L, name); f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.fzr_writer_encode.restype = C.c_int
        L.fzr_writer_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_size_t)]
        L.fzr_bulk_compress.restype = C.c_int
        L.fzr_bulk_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.fzr_batch.restype = C.c_double
        L.fzr_batch.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int]
        self.L = L

    @property
    def available(self):
        return bool(self.L.fzr_available())

    @property
    def version(self):
        return self.L.fzr_version()

    def bound(self, n):
        return self.L.fzr_compress_bound(n)

    def copy_decode(self, src, cap, oneshot=False):
        p, n, keep = _buf(src)
        out = np.empty(max(cap, 1), dtype=np.uint8); got = C.c_size_t(0)
        f = self.L.fzr_decode_oneshot if oneshot else self.L.fzr_copy_decode
        st = f(p, n, out.ctypes.data, cap, C.byref(got))
        return st, out[:got.value].tobytes()

    def writer_encode(self, data, level=0, pledge=True, checksum=True, window_log=0):
        """The reference's writer (src/main.rs:781-791) by default."""
        p, n, keep = _buf(data)
        cap = self.bound(n) + 64
        out = np.empty(cap, dtype=np.uint8); got = C.c_size_t(0)
        st = self.L.fzr_writer_encode(p, n, out.ctypes.data, cap, level, int(pledge), int(checksum), window_log, C.byref(got))
        if st != 0:
            raise RuntimeError("libzstd encode failed: %d" % st)
        return out[:got.value].tobytes()

    def bulk_compress(self, data, level=0):
        """zstd::bulk::compress(data, level) -- tests/convert.rs:18"""
        p, n, keep = _buf(data)
        cap = self.bound(n) + 64
        out = np.empty(cap, dtype=np.uint8); got = C.c_size_t(0)
        st = self.L.fzr_bulk_compress(p, n, out.ctypes.data, cap, level, C.byref(got))
        if st != 0:
            raise RuntimeError("libzstd compress failed: %d" % st)
        return out[:got.value].tobytes()

    def batch(self, mode, threads, src_ptrs, src_lens, dst_ptrs, dst_caps, level=3):
        """mode 0 copy_decode / 1 one-shot decode / 2 writer encode -> (seconds, out_len[], status[])"""
        n = len(src_ptrs)
        out_len = np.zeros(n, dtype=np.uint64); status = np.zeros(n, dtype=np.int32)
        t = self.L.fzr_batch(mode, threads, n, src_ptrs.ctypes.data, src_lens.ctypes.data, dst_ptrs.ctypes.data,
                             dst_caps.ctypes.data, out_len.ctypes.data, status.ctypes.data, level)
        return t, out_len, status
