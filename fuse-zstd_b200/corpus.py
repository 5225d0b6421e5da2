"""Deterministic synthetic JSON-lines corpus -- bench/test support, host only.

File i is generated from seed 20261018 + i by csrc/corpus_gen.c (built into libfzcorpus.so).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfzcorpus.so")
SRC = os.path.join(HERE, "csrc", "corpus_gen.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", SO, SRC, "-lm", "-lpthread"])


def _get():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.fzc_generate.argtypes = [C.c_uint64, C.c_void_p, C.c_size_t]
        _lib.fzc_generate_many.argtypes = [C.c_uint64, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
    return _lib


def json_file(index, size):
    """-> numpy uint8[size]: synthetic JSON-lines file number `index`."""
    out = np.empty(size, dtype=np.uint8)
    if size:
        _get().fzc_generate(index, out.ctypes.data, size)
    return out


def json_files(first, n, size, threads=None, out=None):
    """-> numpy uint8[n, size]: files first..first+n-1 generated on `threads` host threads."""
    if out is None:
        out = np.empty((n, size), dtype=np.uint8)
    if n and size:
        _get().fzc_generate_many(first, n, out.ctypes.data, size, out.strides[0], threads or os.cpu_count() or 1)
    return out


def json_files_idx(indices, size, threads=None, out=None):
    """-> numpy uint8[len(indices), size]: the files with the given indices (any order, e.g. one rank's shard by inode)."""
    from concurrent.futures import ThreadPoolExecutor
    idx = [int(i) for i in indices]
    if out is None:
        out = np.empty((len(idx), size), dtype=np.uint8)
    L = _get()
    if idx and size:
        def one(k):
            L.fzc_generate(idx[k], out[k].ctypes.data, size)            # ctypes releases the GIL
        with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
            list(ex.map(one, range(len(idx))))
    return out
