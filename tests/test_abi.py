"""The C-ABI shared library loads and exports every symbol include/fzgpu.h declares (CPU, no GPU)."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def codec():
    m = importlib.import_module("fuse-zstd_b200.codec")
    m.build()
    return m


def test_every_declared_symbol_is_exported(codec):
    hdr = open(os.path.join(ROOT, "include", "fzgpu.h")).read()
    declared = set(re.findall(r"\b(fzg_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(codec.EXPORTS), declared ^ set(codec.EXPORTS)
    L = C.CDLL(codec.SO)
    for name in declared:
        assert hasattr(L, name), name


def test_no_oracle_linkage(codec):
    """the product library must not link, load or reference anything under oracle/"""
    blob = open(codec.SO, "rb").read()
    for needle in (b"fzo_", b"fzr_", b"libfzoracle", b"libfzref", b"libzstd"):
        assert needle not in blob, needle


def test_frame_info_host_walk(codec, golden):
    for name, (comp, meta) in golden.items():
        st, size, csize = codec.frame_info(comp)
        assert st == 0, name
        if "nopledge" in name:
            assert size is None
        else:
            assert size == meta["plain_len"], name
        assert csize == len(comp)
    assert codec.frame_info(b"junkjunk")[0] == codec.E_MAGIC
    assert codec.frame_info(golden["json_2000_L3_writer"][0][:100])[0] == codec.E_TRUNCATED


def test_strerror(codec):
    assert b"checksum" in codec.lib().fzg_strerror(codec.E_CHECKSUM)


def test_compute_fails_loudly_without_gpu(codec, golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(OSError) as e:
        codec.decode_batch([golden["json_2000_L3_writer"][0]])
    assert e.value.errno == 19   # ENODEV: no CPU fallback


def test_seek_footer_parser_host_only():
    """fzg_seek_footer (no GPU): the footer of the zstd seekable format -- Number_Of_Frames | descriptor | 0x8F92EAB1"""
    import struct
    import errno
    import importlib
    codec = importlib.import_module("fuse-zstd_b200.codec")
    foot = struct.pack("<IBI", 5, 0, 0x8F92EAB1)
    assert codec.seek_footer(foot, 1000) == (0, 5, 8 + 8 * 5 + 9)
    assert codec.seek_footer(b"xxxx" + foot, 1000) == (0, 5, 57)                        # a longer tail is fine
    assert codec.seek_footer(struct.pack("<IBI", 5, 0x80, 0x8F92EAB1), 1000) == (0, 5, 8 + 12 * 5 + 9)   # entries with checksums
    assert codec.seek_footer(struct.pack("<IBI", 5, 0, 0x8F92EAB2), 1000)[0] == -errno.ENOENT          # wrong magic
    assert codec.seek_footer(struct.pack("<IBI", 5, 0x04, 0x8F92EAB1), 1000)[0] == -errno.ENOENT       # reserved bits
    assert codec.seek_footer(foot, 40)[0] == -errno.ENOENT                                # table larger than the file
    assert codec.seek_footer(foot[:8], 1000)[0] == -errno.ENOENT


def test_frame_info_sums_do_not_wrap():
    """two frames whose Frame_Content_Size fields add up past 2^64: refused, not wrapped (host-side header walk, no GPU)"""
    import struct
    codec = importlib.import_module("fuse-zstd_b200.codec")
    one = struct.pack("<IB", 0xFD2FB528, 0xC0) + bytes([0x00]) + struct.pack("<Q", (1 << 64) - 16) + bytes([1, 0, 0])   # windowed, empty last Raw block
    st, content, csize = codec.frame_info(one)
    assert st == 0 and content == (1 << 64) - 16 and csize == len(one)
    st, content, _ = codec.frame_info(one + one)
    assert st == codec.E_UNSUPPORTED
