"""Host logic of the decoder (CPU, no GPU): the per-thread device code of fz_core.cuh /
fz_kernels.cuh compiled with g++ and run through the same stage order as fz_decode.cu, checked
against the oracle.  This validates descriptor / pipeline logic; the CUDA parity tests proper
are in test_gpu_decode.py."""
import hashlib

import numpy as np
import pytest

import emul_util


def test_golden_through_emulated_pipeline(golden):
    names = sorted(golden)
    res = emul_util.decode_batch([golden[n][0] for n in names], [golden[n][1]["plain_len"] for n in names])
    for n, (st, out) in zip(names, res):
        meta = golden[n][1]
        assert st == 0, n
        assert len(out) == meta["plain_len"] and hashlib.sha256(out).hexdigest() == meta["plain_sha256"], n


def test_stage_level_records_match_oracle(golden, oracle):
    """literals and (ll, ml, resolved distance) per sequence before execution == the oracle's trace; this covers the
    cumulative-position records and the symbolic repeat-offset history that is resolved across blocks"""
    for n in ("json_150000_L3_writer", "json_200k_L19_writer", "rle_mode_of_ml", "rle_mode_ll_ml", "rle_literals",
              "json_120k_L19_wlog11_repeat", "direct_weights_huffman", "long_repeat_300k"):
        comp, meta = golden[n]
        tr = oracle.decode_trace(comp, meta["plain_len"])
        st, seqs, lits = emul_util.trace(comp)
        assert st == 0, n
        assert np.array_equal(lits, tr["literals"]), n
        s = tr["seqs"]
        assert len(seqs) == len(s), n
        assert np.array_equal(seqs & 0x1FFFF, s[:, 0]), n
        assert np.array_equal((seqs >> 17) & 0x3FFFF, s[:, 1]), n
        assert np.array_equal(seqs >> 35, s[:, 2]), n


def test_error_statuses_match_oracle(golden, oracle):
    comp, meta = golden["json_20000_L3_writer"]
    n = meta["plain_len"]
    cases = [b"", b"not a zstd file at all", comp[:-1], comp[: len(comp) // 2], comp + b"\x00", comp + b"garbage!",
             bytes.fromhex("28b52ffd2000") + bytes([0x07, 0, 0]), bytes.fromhex("28b52ffd21") + b"\x05\x00" + bytes([1, 0, 0]),
             bytes.fromhex("28b52ffd00") + bytes([18 << 3]) + bytes([1, 0, 0])]
    for pos, mask in ((len(comp) - 1, 0x55), (4, 0x08), (len(comp) // 2, 0xFF), (40, 0x01), (300, 0x80)):
        bad = bytearray(comp); bad[pos] ^= mask; cases.append(bytes(bad))
    frame = bytearray(golden["ref_compressed_data_bulk"][0]); frame[5] = 16; cases.append(bytes(frame))
    res = emul_util.decode_batch(cases, [n] * len(cases))
    for c, (st, out) in zip(cases, res):
        st_o, out_o = oracle.decode(c, cap=n)
        assert (st == 0) == (st_o == 0), (st, st_o, c[:16])
        if st_o == 0:
            assert out == out_o
        else:
            assert st == st_o, (st, st_o, c[:16].hex())
    # destination too small
    (st, _), = emul_util.decode_batch([comp], [n - 1])
    assert st in (5, 4)


def test_mutation_fuzz_matches_oracle(golden, oracle):
    rs = np.random.RandomState(99)
    for name in ("json_20000_L3_writer", "json_200k_L3_nopledge_nochk", "json_120k_L19_wlog11_repeat"):
        comp, meta = golden[name]
        cases = []
        for _ in range(60):
            bad = bytearray(comp)
            for _ in range(rs.randint(1, 4)):
                bad[rs.randint(0, len(bad))] ^= 1 << rs.randint(0, 8)
            cases.append(bytes(bad))
        res = emul_util.decode_batch(cases, [meta["plain_len"] + 64] * len(cases), flags=0)
        for c, (st, out) in zip(cases, res):
            st_o, out_o = oracle.decode(c, cap=meta["plain_len"] + 64)
            assert (st == 0) == (st_o == 0), (name, st, st_o)
            if st == 0:
                assert out == out_o


def test_far_form_records(ref, oracle):
    """a sequence whose extra bits exceed the 32-bit window of a RAW record travels as a bit cursor (stage A) and is
    read back from the stream by stage B: long literal run + far offset + long match in one sequence"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    plain = emul_util.far_offset_long_length_plain()
    comp = ref.writer_encode(plain, 19)
    before = emul_util.far_records()
    (st, out), = emul_util.decode_batch([comp], [len(plain)])
    assert st == 0 and out == plain
    assert emul_util.far_records() - before >= 10
    st_o, out_o = oracle.decode(comp, cap=len(plain))
    assert st_o == 0 and out_o == plain


def test_batch_of_mixed_items(golden):
    good, meta = golden["json_2000_L3_writer"]
    blobs = [good, b"junk", golden["ref_touch_empty_writer"][0], good[:50], golden["multi_frame_skippable"][0]]
    caps = [meta["plain_len"], 10, 0, meta["plain_len"], golden["multi_frame_skippable"][1]["plain_len"]]
    res = emul_util.decode_batch(blobs, caps)
    assert [r[0] for r in res] == [0, 1, 0, 2, 0]
    assert hashlib.sha256(res[4][1]).hexdigest() == golden["multi_frame_skippable"][1]["plain_sha256"]
    assert res[2][1] == b""


def test_encoder_fse_tables_round_trip_through_the_decoder_tables():
    """normalisation + table description writer + encoding table vs read_ncount + decoding table (CPU, no GPU)"""
    rs = np.random.RandomState(5)
    cases = []
    for n_sym, max_log in ((36, 9), (32, 8), (53, 9)):
        for n in (70, 300, 5000, 14000):
            cases.append((rs.randint(0, n_sym, n), n_sym, max_log))                                  # flat
            cases.append((np.minimum(rs.geometric(0.35, n) - 1, n_sym - 1), n_sym, max_log))         # skewed, like LL / ML codes
            cases.append((rs.choice([0, 1, n_sym - 1], n, p=[0.9, 0.09, 0.01]), n_sym, max_log))     # gaps of absent symbols
            z = rs.randint(0, 3, n); z[0] = n_sym - 1; cases.append((z, n_sym, max_log))             # one rare high symbol
            cases.append((rs.choice([3, 30 if n_sym > 30 else 7], n), n_sym, max_log))               # two symbols, long zero runs
    for syms, n_sym, max_log in cases:
        assert emul_util.fse_roundtrip(syms, n_sym, max_log) == 0, (n_sym, len(syms), np.bincount(syms)[:8])
