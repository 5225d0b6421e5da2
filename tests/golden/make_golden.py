#!/usr/bin/env python
"""Generates tests/golden/*.zst + manifest.json.  Run in the build container:

    python tests/golden/make_golden.py

Sources of truth:
  * "ref-tests": the byte strings the reference's own tests pin for this path
    (/root/reference/tests/cmdline.rs:34-43,160-178, tests/convert.rs:16-43,54-98) -- produced here
    by replaying the reference's two writers against the system libzstd (oracle/ref_libzstd.c):
    zstd::bulk::compress(data, 0) and the Encoder{level, pledged size, checksum} flow of
    /root/reference/src/main.rs:781-791 -- and checked against the hex recorded in SURVEY.md §8c.
  * "modes": frames made by libzstd 1.5.5 (levels 1/3/19, several window/flag settings) from inputs
    chosen to reach every block type, literals type, Huffman header form and sequence table mode
    (SURVEY.md Appendix A "inputs that trigger the rare paths").

For every fixture the manifest stores the sha256 + length of the plain bytes as libzstd decodes
them (never as the oracle does), plus which format features the frame exercises.
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402

corpus = importlib.import_module("fuse-zstd_b200.corpus")

R = pyoracle.Ref()
O = pyoracle.Oracle()
assert R.available, "system libzstd needed to (re)generate fixtures"

SURVEY_HEX = {
    "ref_touch_empty_writer": "28b52ffd2400010000" "99e9d851",
    "ref_compressed_data_bulk": "28b52ffd200f790000636f6d707265737365642064617461",
    "ref_compressed_data_writer": "28b52ffd240f790000636f6d707265737365642064617461ca9d2b2c",
    "ref_overlap_compressed_bulk": "28b52ffd20129100006f7665726c617020636f6d70726573736564",
    "ref_truncated_and_appended_writer": "28b52ffd2416b100007472756e636174656420616e6420617070656e6465643d98a66b",
    "ref_empty_bulk": "28b52ffd2000010000",
}


def rng(seed):
    return np.random.RandomState(seed)


def json_bytes(index, size):
    return corpus.json_file(index, size).tobytes()


def build_inputs():
    """name -> (plain bytes, [(variant, compressed bytes)])"""
    fx = {}

    def add(name, plain, comp, note=""):
        fx[name] = (bytes(plain), bytes(comp), note)

    # ---- the reference's own test payloads
    add("ref_touch_empty_writer", b"", R.writer_encode(b"", 0), "tests/cmdline.rs:34-43")
    add("ref_compressed_data_bulk", b"compressed data", R.bulk_compress(b"compressed data", 0), "tests/convert.rs:16-25")
    add("ref_compressed_data_writer", b"compressed data", R.writer_encode(b"compressed data", 0), "tests/convert.rs:189-212")
    add("ref_overlap_compressed_bulk", b"overlap compressed", R.bulk_compress(b"overlap compressed", 0), "tests/convert.rs:33-43")
    add("ref_overlap_compressed_writer", b"overlap compressed", R.writer_encode(b"overlap compressed", 0), "tests/convert.rs:222-263")
    add("ref_truncated_and_appended_writer", b"truncated and appended", R.writer_encode(b"truncated and appended", 0),
        "tests/cmdline.rs:160-178")
    add("ref_empty_bulk", b"", R.bulk_compress(b"", 0), "zstd::bulk::compress(b\"\", 0)")
    for s in (b"1st file in third", b"truncated", b"KEEPIT", b"OVERRIDE", b"TOO CLOSE"):
        add("ref_roundtrip_" + s.decode().replace(" ", "_"), s, R.writer_encode(s, 0), "tests/cmdline.rs:96-179, tests/glitches.rs:93-262")

    # ---- mode coverage
    for size in (300, 900, 2000, 20000, 150000, 300000):
        d = json_bytes(1000 + size, size)
        add("json_%d_L3_writer" % size, d, R.writer_encode(d, 3))
    d = json_bytes(7, 200000)
    add("json_200k_L1_bulk", d, R.bulk_compress(d, 1))
    add("json_200k_L19_writer", d, R.writer_encode(d, 19))
    add("json_200k_L3_nopledge_nochk", d, R.writer_encode(d, 3, pledge=False, checksum=False))
    add("json_200k_L3_wlog10", d, R.writer_encode(d, 3, window_log=10))
    add("json_200k_L5_wlog17", d, R.writer_encode(d, 5, window_log=17))
    d = json_bytes(8, 70000)
    add("json_70k_L3_fcs2", d[:40000], R.bulk_compress(d[:40000], 3))          # 2-byte FCS (+256 rule)
    add("json_200_L3_fcs1", d[:200], R.bulk_compress(d[:200], 3))              # 1-byte FCS
    add("rle_block_a400k", b"a" * 400000, R.writer_encode(b"a" * 400000, 3))   # RLE blocks
    r = rng(7)
    d = b"".join(b'{"k":%d,"v":"%s"}\n' % (i % 97, bytes(r.randint(97, 123, 8, dtype=np.uint8))) for i in range(6000))
    add("rle_mode_of_ml", d, R.writer_encode(d, 19, window_log=14))          # RLE-mode OF / ML, Repeat LL / ML
    r = rng(12); pool = bytes(r.randint(0, 128, 4096, dtype=np.uint8))
    first = (pool * 32)[:131072]
    second = b"".join(bytes([r.randint(128, 256)]) + pool[a:a + 32] for a in r.randint(0, 4000, 500))
    add("rle_mode_ll_ml", first + second, R.writer_encode(first + second, 19))  # RLE-mode LL / ML
    r = rng(2); pool = bytes(r.randint(0, 256, 4096, dtype=np.uint8))
    first = (pool * 32)[:131072]
    for k in (1, 2):
        second = b"".join(b"Q" * k + pool[a:a + 40] for a in r.randint(0, 4000, 2000))
    add("rle_literals", first + second, R.writer_encode(first + second, 19))   # RLE literals section
    d = json_bytes(13, 120000)
    add("json_120k_L19_wlog11_repeat", d, R.writer_encode(d, 19, window_log=11))  # Repeat modes, many small blocks
    r = rng(3); d = bytes(r.randint(0, 16, 3000, dtype=np.uint8))
    add("direct_weights_huffman", d, R.writer_encode(d, 3))                   # h >= 128 header
    r = rng(4); d = bytes(r.randint(0, 256, 5000, dtype=np.uint8))
    add("incompressible_5000_raw", d, R.writer_encode(d, 3))                  # Raw block
    r = rng(5); d = bytes(r.randint(0, 256, 140000, dtype=np.uint8))
    add("incompressible_140k_raw", d, R.writer_encode(d, 1))                  # multi Raw blocks
    r = rng(6); d = bytes(r.randint(97, 101, 100000, dtype=np.uint8))
    add("lowentropy_100k", d, R.writer_encode(d, 3))                          # literal-heavy, few sequences
    d = (b"abcdefghij" * 30000)
    add("long_repeat_300k", d, R.writer_encode(d, 3))                         # overlapping matches, huge ML
    d = b"x" + b"ab" * 70000 + b"y" * 5 + b"abc" * 33333
    add("overlap_small_offsets", d, R.writer_encode(d, 19))
    d = json_bytes(9, 300000)
    add("json_300k_L19_wlog", d, R.writer_encode(d, 19, window_log=18))        # Repeat modes, block splitting
    # concatenated + skippable frames
    a = json_bytes(10, 50000); b = json_bytes(11, 3000)
    skip = bytes.fromhex("502a4d18") + (11).to_bytes(4, "little") + b"hello world"
    skip2 = bytes.fromhex("5f2a4d18") + (0).to_bytes(4, "little")
    add("multi_frame_skippable", a + b + b"" + a[:100],
        R.writer_encode(a, 3) + skip + R.bulk_compress(b, 1) + skip2 + R.writer_encode(b"", 0) + R.writer_encode(a[:100], 19))
    # multi-frame file of independent chunks (the layout our encoder emits)
    d = json_bytes(12, 300000)
    add("multi_frame_chunks", d, b"".join(R.writer_encode(d[i:i + 65536], 3) for i in range(0, len(d), 65536)))
    return fx


def features(comp, plain_len):
    tr = O.decode_trace(comp, max(plain_len, 1), max_seqs=1 << 20)
    f = set()
    for t, lt, md, ns in zip(tr["block_type"], tr["block_littype"], tr["block_modes"], tr["block_nseq"]):
        f.add(("block_raw", "block_rle", "block_compressed")[t])
        if t == 2:
            f.add(("lit_raw", "lit_rle", "lit_huffman", "lit_treeless")[lt & 3])
            if (lt & 3) >= 2:
                f.add("huf_4streams" if lt & 4 else "huf_1stream")
            if (lt & 3) == 2:
                f.add("huf_weights_fse" if lt & 8 else "huf_weights_direct")
            if ns:
                for nm, sh in (("ll", 6), ("of", 4), ("ml", 2)):
                    f.add("%s_%s" % (nm, ("predefined", "rle", "fse", "repeat")[(md >> sh) & 3]))
            else:
                f.add("nseq_0")
    s = tr["seqs"]
    if len(s):
        if (s[:, 2] < s[:, 1]).any():
            f.add("overlap_match")
        rep = s[:, 3] <= 3
        if rep.any():
            f.add("repcode")
        if ((s[:, 3] == 3) & (s[:, 0] == 0)).any():
            f.add("repcode_rep0_minus_1")
    return tr, sorted(f)


def main():
    fx = build_inputs()
    manifest = {}
    allf = set()
    for name, (plain, comp, note) in sorted(fx.items()):
        st, dec = R.copy_decode(comp, len(plain) + 16)
        assert st == 0 and dec == plain, name
        st1, dec1 = R.copy_decode(comp, len(plain) + 16, oneshot=True)
        assert st1 == 0 and dec1 == plain, name
        if name in SURVEY_HEX:
            assert comp.hex() == SURVEY_HEX[name], (name, comp.hex())
        tr, feats = features(comp, len(plain))
        assert tr["status"] == 0 and tr["out"] == plain, ("oracle disagrees with libzstd", name, tr["status"])
        allf.update(feats)
        with open(os.path.join(HERE, name + ".zst"), "wb") as fh:
            fh.write(comp)
        manifest[name] = dict(plain_len=len(plain), plain_sha256=hashlib.sha256(plain).hexdigest(),
                              comp_len=len(comp), features=feats, note=note,
                              hex=comp.hex() if len(comp) <= 64 else None)
        print("%-36s plain %7d comp %7d  %s" % (name, len(plain), len(comp), " ".join(feats)))
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(dict(libzstd_version=R.version, vectors=manifest, all_features=sorted(allf)), fh, indent=1, sort_keys=True)
    print("features covered:", sorted(allf))
    print("total bytes:", sum(len(c) for _, c, _ in fx.values()))


if __name__ == "__main__":
    main()
