"""Pins the CPU oracle (oracle/zstd_oracle.c) -- CPU only, no GPU.

(1) golden vectors the reference's own tests hold for the path (Raw-block frames);
(2) committed libzstd-made fixtures covering every block / literals / sequence mode;
(3) live differential runs against the system libzstd when present.
"""
import hashlib

import numpy as np
import pytest

REF_HEX = {  # SURVEY.md §8c; /root/reference/tests/cmdline.rs:34-43,160-178, tests/convert.rs:16-43
    "ref_touch_empty_writer": ("28b52ffd240001000099e9d851", b""),
    "ref_compressed_data_bulk": ("28b52ffd200f790000636f6d707265737365642064617461", b"compressed data"),
    "ref_compressed_data_writer": ("28b52ffd240f790000636f6d707265737365642064617461ca9d2b2c", b"compressed data"),
    "ref_overlap_compressed_bulk": ("28b52ffd20129100006f7665726c617020636f6d70726573736564", b"overlap compressed"),
    "ref_truncated_and_appended_writer": ("28b52ffd2416b100007472756e636174656420616e6420617070656e6465643d98a66b",
                                          b"truncated and appended"),
    "ref_empty_bulk": ("28b52ffd2000010000", b""),
}


def test_xxh64_known_answers(oracle):
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    # trailers of the reference-writer vectors are the low 32 bits of XXH64(content)
    assert oracle.xxh64(b"compressed data") & 0xFFFFFFFF == 0x2C2B9DCA
    assert oracle.xxh64(b"truncated and appended") & 0xFFFFFFFF == 0x6BA6983D


@pytest.mark.parametrize("name", sorted(REF_HEX))
def test_reference_test_vectors(oracle, golden, name):
    hexs, plain = REF_HEX[name]
    comp, meta = golden[name]
    assert comp.hex() == hexs                      # fixture file == bytes the reference's tests pin
    st, out = oracle.decode(bytes.fromhex(hexs))
    assert st == 0 and out == plain
    st, size, frames = oracle.frame_info(comp)
    assert (st, size, frames) == (0, len(plain), 1)  # user.real_size the reference records after open


def test_golden_all(oracle, golden):
    for name, (comp, meta) in golden.items():
        st, out = oracle.decode(comp, cap=meta["plain_len"])
        assert st == 0, name
        assert len(out) == meta["plain_len"], name
        assert hashlib.sha256(out).hexdigest() == meta["plain_sha256"], name


def test_golden_feature_coverage(golden):
    feats = set()
    for comp, meta in golden.values():
        feats.update(meta["features"])
    need = {"block_raw", "block_rle", "block_compressed", "lit_raw", "lit_rle", "lit_huffman", "lit_treeless",
            "huf_1stream", "huf_4streams", "huf_weights_direct", "huf_weights_fse", "nseq_0", "overlap_match",
            "repcode", "repcode_rep0_minus_1"}
    for t in ("ll", "of", "ml"):
        need |= {"%s_%s" % (t, m) for m in ("predefined", "rle", "fse", "repeat")}
    assert need <= feats, need - feats


def test_error_paths(oracle, golden):
    comp, meta = golden["json_20000_L3_writer"]
    n = meta["plain_len"]
    assert oracle.decode(b"")[0] == 0                                  # empty input: success, empty output
    assert oracle.decode(b"not a zstd file at all")[0] == 1            # bad magic  -> EFAULT in the reference
    assert oracle.decode(comp[:-1], cap=n)[0] == 2                     # truncated
    assert oracle.decode(comp[: len(comp) // 2], cap=n)[0] == 2
    assert oracle.decode(comp + b"\x00", cap=n)[0] == 2                # trailing garbage
    assert oracle.decode(comp + b"garbage!", cap=n)[0] == 1
    bad = bytearray(comp); bad[-1] ^= 0x55
    assert oracle.decode(bytes(bad), cap=n)[0] == 6                    # checksum mismatch
    bad = bytearray(comp); bad[4] |= 0x08
    assert oracle.decode(bytes(bad), cap=n)[0] == 3                    # reserved FHD bit
    bad = bytearray(comp); bad[len(comp) // 2] ^= 0xFF
    assert oracle.decode(bytes(bad), cap=n)[0] != 0                    # corruption somewhere in a block
    assert oracle.decode(comp, cap=n - 1)[0] in (5, 4)                 # destination too small
    # reserved block type 3
    frame = bytes.fromhex("28b52ffd2000") + bytes([0x07, 0, 0])
    assert oracle.decode(frame)[0] == 4
    # FCS mismatch: header says 16 bytes, raw block carries 15
    frame = bytearray(golden["ref_compressed_data_bulk"][0]); frame[5] = 16
    assert oracle.decode(bytes(frame), cap=64)[0] in (7, 4)
    # dictionary id present and non-zero
    frame = bytes.fromhex("28b52ffd21") + b"\x05" + b"\x00" + bytes([0x01, 0, 0])
    assert oracle.decode(frame)[0] == 3
    # window descriptor above 2^27
    frame = bytes.fromhex("28b52ffd00") + bytes([(18 << 3)]) + bytes([0x01, 0, 0])
    assert oracle.decode(frame)[0] == 3


def _variants(ref, data):
    yield ref.writer_encode(data, 0)
    yield ref.writer_encode(data, 1, pledge=False, checksum=False)
    yield ref.bulk_compress(data, 3)
    yield ref.writer_encode(data, 9, window_log=12)


def test_differential_vs_libzstd(oracle, ref, corpus):
    if not ref.available:
        pytest.skip("system libzstd not present")
    rs = np.random.RandomState(20261018)
    cases = [corpus.json_file(i, s).tobytes() for i, s in enumerate((0, 1, 7, 64, 1000, 5000, 131072, 131073, 400000))]
    cases += [bytes(rs.randint(0, 256, 70000, dtype=np.uint8)), bytes(rs.randint(0, 4, 70000, dtype=np.uint8)),
              b"z" * 262144, (b"0123456789abcdef" * 9000)[:140001]]
    for data in cases:
        for comp in _variants(ref, data):
            st_r, out_r = ref.copy_decode(comp, len(data) + 8)
            st_o, out_o = oracle.decode(comp, cap=len(data))
            assert st_r == 0 and st_o == 0
            assert out_o == out_r == data


def test_differential_corruption_vs_libzstd(oracle, ref, corpus):
    """Mutated frames (no checksum, so mutations reach the codec): every frame the oracle accepts must be
    accepted by libzstd with identical bytes.  The converse is not required: libzstd >= 1.5.4's fast
    4-stream Huffman loop skips the end-of-bitstream check the format demands, so it "decodes" some corrupt
    literal streams that the oracle (and libzstd's own non-fast path) reject as corrupt."""
    if not ref.available:
        pytest.skip("system libzstd not present")
    rs = np.random.RandomState(7)
    data = corpus.json_file(3, 30000).tobytes()
    comp = ref.writer_encode(data, 3, checksum=False)
    both = stricter = 0
    for _ in range(300):
        bad = bytearray(comp)
        for _ in range(rs.randint(1, 3)):
            bad[rs.randint(6, len(bad))] ^= 1 << rs.randint(0, 8)
        st_r, out_r = ref.copy_decode(bytes(bad), 1 << 20)
        st_o, out_o = oracle.decode(bytes(bad), cap=1 << 20)
        if st_o == 0:
            assert st_r == 0 and out_r == out_o
            both += 1
        elif st_r == 0:
            assert st_o == 4
            stricter += 1
    assert both > 50 and stricter < both
