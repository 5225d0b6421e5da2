"""TEST ONLY: loads tests/emul/libfzemul.so -- the decoder's per-thread device code compiled for the
host, replaying the launch sequence serially (see tests/emul/emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "libfzemul.so")
SRC = os.path.join(HERE, "emul", "emul.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "fuse-zstd_b200", "csrc")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in ("fz_core.cuh", "fz_kernels.cuh", "fz_enc_core.cuh")]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-fsanitize=undefined",
                                   "-fno-sanitize-recover=undefined", "-o", SO, SRC])
        _lib = C.CDLL(SO)
    return _lib


def decode_batch(blobs, caps, flags=0):
    n = len(blobs)
    arrs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
    outs = [np.zeros(max(c, 1), dtype=np.uint8) for c in caps]
    sp = (C.c_void_p * n)(*[a.ctypes.data if a.size else 0 for a in arrs]); sl = (C.c_size_t * n)(*[a.size for a in arrs])
    dp = (C.c_void_p * n)(*[o.ctypes.data for o in outs]); dc = (C.c_size_t * n)(*caps)
    dl = (C.c_size_t * n)(); st = (C.c_int * n)()
    rc = lib().fze_decode_batch(C.c_size_t(n), sp, sl, dp, dc, dl, st, flags)
    assert rc == 0, rc
    return [(st[i], outs[i][:dl[i]].tobytes()) for i in range(n)]


def trace(blob, max_seq=1 << 21, max_lit=1 << 22):
    a = np.frombuffer(blob, dtype=np.uint8)
    seqs = np.zeros(max_seq, dtype=np.uint64); lits = np.zeros(max_lit, dtype=np.uint8)
    ns, nl = C.c_size_t(0), C.c_size_t(0)
    st = lib().fze_trace(C.c_void_p(a.ctypes.data), C.c_size_t(a.size), C.c_void_p(seqs.ctypes.data), C.c_size_t(max_seq),
                         C.byref(ns), C.c_void_p(lits.ctypes.data), C.c_size_t(max_lit), C.byref(nl))
    return st, seqs[:ns.value], lits[:nl.value]


def fse_roundtrip(syms, n_sym, max_log):
    a = np.ascontiguousarray(syms, dtype=np.uint8)
    return lib().fze_fse_roundtrip(C.c_void_p(a.ctypes.data), C.c_size_t(a.size), n_sym, max_log)


def far_records():
    """RAW records in the far form (extra bits of one sequence > 32: stage A stores the bit cursor) seen so far"""
    L = lib(); L.fze_far_records.restype = C.c_uint64
    return L.fze_far_records()


def far_offset_long_length_plain(seed=11, base=200000, pieces=40):
    """plain bytes whose level-19 parse has sequences with a long literal run AND a far offset AND a long match:
    low-entropy noise (Huffman-compressible, hardly matchable) with slices of the first `base` bytes planted far later"""
    rs = np.random.RandomState(seed)

    def noise(n):
        return (rs.randint(0, 12, n) * rs.randint(1, 12, n)).astype(np.uint8).tobytes()
    head = noise(base)
    parts = [head]
    for j in range(pieces):
        parts.append(noise(4200 + j * 900))
        src = 3000 + j * 2000
        parts.append(head[src:src + 140 + j * 25])
    return b"".join(parts)
