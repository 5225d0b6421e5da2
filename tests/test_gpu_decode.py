"""Parity tests proper (B200): the CUDA decoder, called through the C ABI (include/fzgpu.h), against

  * the committed golden fixtures (tests/golden: frames libzstd 1.5.5 produced for every block /
    literals / sequence mode + the byte strings the reference's own tests pin,
    /root/reference/tests/cmdline.rs:34-43,160-178, tests/convert.rs:16-43,54-98),
  * the plain-C oracle (oracle/zstd_oracle.c) on seeded inputs, bit-exact, including error statuses,
  * size-independent properties at larger sizes (stored XXH64 verified on device, FCS == produced).

Nothing here reads /root/reference.  Every test needs a GPU.
"""
import hashlib
import importlib
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

codec = importlib.import_module("fuse-zstd_b200.codec")
stream = importlib.import_module("fuse-zstd_b200.stream")


@pytest.fixture(scope="module", autouse=True)
def _init():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    codec.build()
    codec.init([0])
    yield


@pytest.fixture(params=["1", "8", "s", "d", "p4", "t128", "t1024"])
def exec_w(request):
    """the execute kernel (FZG_EXEC_W): 1 = k_execute<false> (warp per frame, a lane loops over its sequence: the default for
    large batches), 8 = k_execute_cta<8> (bitmap dataflow: small batches), and round 2's alternatives -- s = k_execute<true>
    (steps dealt out over the lanes), d = k_execute2 (two sequences per lane), p4 = k_execute_pass<4> (four warps per frame,
    barrier passes), t128 / t1024 = k_execute_tile (output-centric, shared-memory window, TMA record ring); without the
    fixture the library chooses by batch shape"""
    old = os.environ.get("FZG_EXEC_W")
    os.environ["FZG_EXEC_W"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("FZG_EXEC_W", None)
    else:
        os.environ["FZG_EXEC_W"] = old


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_golden_vectors_bit_exact(golden, oracle, exec_w):
    names = sorted(golden)
    res = codec.decode_batch([golden[n][0] for n in names], [golden[n][1]["plain_len"] for n in names])
    for n, (st, out) in zip(names, res):
        meta = golden[n][1]
        assert st == 0, (n, codec.strerror(st))
        assert len(out) == meta["plain_len"] and sha(out) == meta["plain_sha256"], n
        st_o, out_o = oracle.decode(golden[n][0], cap=meta["plain_len"])
        assert st_o == 0 and out_o == out, n
    t = codec.last_timing(0)
    assert t["launches"] >= 5          # the CUDA pipeline ran (no fallback exists)


def test_golden_one_item_per_call(golden):
    """the open() flow: one file per call (src/main.rs:463), incl. the 13-byte empty frame"""
    for n in sorted(golden):
        comp, meta = golden[n]
        (st, out), = codec.decode_batch([comp], [meta["plain_len"]])
        assert st == 0 and sha(out) == meta["plain_sha256"], n


def test_error_statuses_match_oracle(golden, oracle):
    comp, meta = golden["json_20000_L3_writer"]
    n = meta["plain_len"]
    cases = [b"", b"not a zstd file at all", comp[:-1], comp[: len(comp) // 2], comp + b"\x00", comp + b"garbage!",
             bytes.fromhex("28b52ffd2000") + bytes([0x07, 0, 0]), bytes.fromhex("28b52ffd21") + b"\x05\x00" + bytes([1, 0, 0]),
             bytes.fromhex("28b52ffd00") + bytes([18 << 3]) + bytes([1, 0, 0])]
    for pos, mask in ((len(comp) - 1, 0x55), (4, 0x08), (len(comp) // 2, 0xFF), (40, 0x01), (300, 0x80)):
        bad = bytearray(comp); bad[pos] ^= mask; cases.append(bytes(bad))
    frame = bytearray(golden["ref_compressed_data_bulk"][0]); frame[5] = 16; cases.append(bytes(frame))
    res = codec.decode_batch(cases, [n] * len(cases))
    for c, (st, out) in zip(cases, res):
        st_o, out_o = oracle.decode(c, cap=n)
        assert (st == 0) == (st_o == 0), (st, st_o, c[:16])
        if st_o == 0:
            assert out == out_o
        else:
            assert st == st_o, (st, st_o, c[:16].hex())
    (st, _), = codec.decode_batch([comp], [n - 1])
    assert st in (codec.E_DSTSIZE, codec.E_CORRUPT)


def test_mutation_fuzz_matches_oracle(golden, oracle, exec_w):
    rs = np.random.RandomState(1234)
    for name in ("json_20000_L3_writer", "json_200k_L3_nopledge_nochk", "json_120k_L19_wlog11_repeat", "rle_mode_of_ml",
                 "direct_weights_huffman"):
        comp, meta = golden[name]
        cases = []
        for _ in range(150):
            bad = bytearray(comp)
            for _ in range(rs.randint(1, 4)):
                bad[rs.randint(0, len(bad))] ^= 1 << rs.randint(0, 8)
            cases.append(bytes(bad))
        cap = meta["plain_len"] + 64
        res = codec.decode_batch(cases, [cap] * len(cases))
        for c, (st, out) in zip(cases, res):
            st_o, out_o = oracle.decode(c, cap=cap)
            assert (st == 0) == (st_o == 0), (name, st, st_o)
            if st == 0:
                assert out == out_o, name


def _mixed_plain(rs, corpus, idx, size):
    kind = idx % 6
    if kind == 0:
        return corpus.json_file(1000 + idx, size).tobytes()
    if kind == 1:
        return rs.randint(0, 256, size, dtype=np.uint8).tobytes()                     # incompressible -> Raw blocks
    if kind == 2:
        return bytes(rs.randint(0, 16, size, dtype=np.uint8))                         # low entropy
    if kind == 3:
        pat = rs.randint(0, 256, 23, dtype=np.uint8).tobytes()
        return (pat * (size // 23 + 1))[:size]                                        # long repeats / RLE-ish
    if kind == 4:
        return (b"a" * size)                                                          # RLE blocks
    j = corpus.json_file(2000 + idx, size).tobytes()
    return j[: size // 2] + j[: size - size // 2]                                     # long-distance match


def test_seeded_corpus_vs_oracle(ref, oracle, corpus, exec_w):
    """levels 1/3/19 x sizes 0..600k x 6 input kinds; reference-writer and bulk framing"""
    if not ref.available:
        pytest.skip("system libzstd absent: cannot produce fresh frames")
    rs = np.random.RandomState(7)
    blobs, plains = [], []
    sizes = [0, 1, 7, 100, 1000, 4096, 65536, 131072, 131073, 200000, 600000]
    for i, size in enumerate(sizes * 3):
        level = (1, 3, 19)[i % 3] if size <= 200000 else (1, 3)[i % 2]
        plain = _mixed_plain(rs, corpus, i, size)
        comp = ref.writer_encode(plain, level) if i % 2 == 0 else ref.bulk_compress(plain, level)
        blobs.append(comp); plains.append(plain)
    res = codec.decode_batch(blobs, [len(p) for p in plains])
    for i, ((st, out), plain, comp) in enumerate(zip(res, plains, blobs)):
        assert st == 0, (i, codec.strerror(st))
        assert out == plain, i
        st_o, out_o = oracle.decode(comp, cap=len(plain))
        assert st_o == 0 and out_o == out


def test_window_and_multiframe(ref, corpus, exec_w):
    if not ref.available:
        pytest.skip("system libzstd absent")
    plain = corpus.json_file(77, 3 << 20).tobytes()
    blobs = [ref.writer_encode(plain, 3, window_log=wl) for wl in (10, 14, 17, 20, 22)]
    blobs.append(ref.writer_encode(plain, 5, pledge=False, checksum=False))
    multi = b"".join(ref.writer_encode(plain[o:o + 300000], 3) for o in range(0, len(plain), 300000))
    skippable = bytes.fromhex("502a4d18") + (5).to_bytes(4, "little") + b"hello"
    blobs.append(skippable + multi + skippable)
    res = codec.decode_batch(blobs, [len(plain)] * len(blobs))
    for i, (st, out) in enumerate(res):
        assert st == 0 and out == plain, i


def test_unaligned_frames_and_their_checksums(ref, corpus, exec_w):
    """concatenated frames of odd sizes: every frame after the first starts at an odd address of the output, which is the
    unaligned path of the checksum that runs beside the execution (widths 8 and 32) and of k_checksum (1, 2); a flipped
    trailer bit in the LAST frame must be reported as a checksum error, nothing else"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    j = corpus.json_file(991, 1 << 20).tobytes()
    sizes = [1001, 77777, 131073, 33, 262147, 5]
    parts, plain, o = [], b"", 0
    for sz in sizes:
        parts.append(ref.writer_encode(j[o:o + sz], 3)); plain += j[o:o + sz]; o += sz
    good = b"".join(parts)
    bad = bytearray(good); bad[-1] ^= 0x10
    res = codec.decode_batch([good, bytes(bad), good], [len(plain)] * 3)
    assert res[0][0] == 0 and res[0][1] == plain
    assert res[1][0] == codec.E_CHECKSUM, codec.strerror(res[1][0])
    assert res[2][0] == 0 and res[2][1] == plain


def test_many_frames_split_between_the_two_executors(ref, corpus):
    """more than 32 x SMs frames: one warp per frame for most of the batch, the last 1/16 on the side stream with 8 warps per
    frame and the checksum beside them (the drain of k_execute); every output compared, one bad trailer in each part found"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    n, size = 5200, 24 * 1024
    plain = corpus.json_files(700000, n, size)
    cap = ref.bound(size) + 64
    comp = np.empty((n, cap), dtype=np.uint8)
    sp = np.array([plain[i].ctypes.data for i in range(n)], dtype=np.uint64); sl = np.full(n, size, dtype=np.uint64)
    dp = np.array([comp[i].ctypes.data for i in range(n)], dtype=np.uint64); dc = np.full(n, cap, dtype=np.uint64)
    _, ol, st = ref.batch(2, os.cpu_count() or 1, sp, sl, dp, dc, 3)
    assert not st.any()
    for bad in (17, n - 5):                                   # one in k_execute's part, one in the tail
        comp[bad, int(ol[bad]) - 1] ^= 0x01
    out = np.zeros((n, size), dtype=np.uint8)
    dl, st = codec.decode_batch_ptrs(0, dp, ol, [out[i].ctypes.data for i in range(n)], [size] * n, 0)
    want = np.zeros(n, dtype=np.int32); want[17] = want[n - 5] = codec.E_CHECKSUM
    assert (st == want).all(), np.nonzero(st != want)[0][:10]
    ok = np.ones(n, dtype=bool); ok[17] = ok[n - 5] = False
    assert (dl[ok] == size).all()
    assert hashlib.sha256(out[ok].tobytes()).digest() == hashlib.sha256(plain[ok].tobytes()).digest()


def test_device_resident_batch(golden):
    """FZG_SRC_DEVICE | FZG_DST_DEVICE: what bench.py's `value` times"""
    import torch
    names = [n for n in sorted(golden) if golden[n][1]["plain_len"] > 0]
    srcs = [torch.frombuffer(bytearray(golden[n][0]), dtype=torch.uint8).cuda() for n in names]
    dsts = [torch.zeros(golden[n][1]["plain_len"] + 32, dtype=torch.uint8, device="cuda") for n in names]
    torch.cuda.synchronize()
    dl, st = codec.decode_batch_ptrs(0, [s.data_ptr() for s in srcs], [s.numel() for s in srcs],
                                     [d.data_ptr() for d in dsts], [golden[n][1]["plain_len"] for n in names],
                                     codec.SRC_DEVICE | codec.DST_DEVICE)
    for n, d, l, s in zip(names, dsts, dl, st):
        assert s == 0 and l == golden[n][1]["plain_len"], n
        host = d.cpu().numpy()
        assert sha(host[:int(l)].tobytes()) == golden[n][1]["plain_sha256"], n
        assert not host[int(l):].any(), "wrote past dst_len: " + n


def test_large_batch_properties(ref, corpus):
    """1024 x 1 MiB level-3 reference-writer files: FCS == produced, stored XXH64 verified on device,
    and a checksum of checksums over the outputs equals the one over the plain inputs."""
    if not ref.available:
        pytest.skip("system libzstd absent")
    n, size = 1024, 1 << 20
    plain = corpus.json_files(500000, n, size)
    cap = ref.bound(size) + 64
    comp = np.empty((n, cap), dtype=np.uint8)
    sp = np.array([plain[i].ctypes.data for i in range(n)], dtype=np.uint64); sl = np.full(n, size, dtype=np.uint64)
    dp = np.array([comp[i].ctypes.data for i in range(n)], dtype=np.uint64); dc = np.full(n, cap, dtype=np.uint64)
    _, ol, st = ref.batch(2, os.cpu_count() or 1, sp, sl, dp, dc, 3)
    assert not st.any()
    out = np.zeros((n, size), dtype=np.uint8)
    dl, st = codec.decode_batch_ptrs(0, dp, ol, [out[i].ctypes.data for i in range(n)], [size] * n, 0)
    assert not st.any() and (dl == size).all()
    assert hashlib.sha256(out.tobytes()).digest() == hashlib.sha256(plain.tobytes()).digest()


def test_fd_entry_points_follow_the_reference_flow(golden):
    """copy_decode(src_file, tmp_file) as open_wrapper drives it (src/main.rs:461-470)"""
    comp, meta = golden["json_300000_L3_writer"]
    with tempfile.TemporaryFile() as src, tempfile.TemporaryFile() as dst:
        src.write(comp); src.flush(); src.seek(0)
        n = stream.copy_decode(src, dst, inode=2**64 - 5)
        assert n == meta["plain_len"]
        assert os.lseek(dst.fileno(), 0, os.SEEK_CUR) == n        # offset left at the end, like io::copy
        dst.seek(0)
        assert sha(dst.read()) == meta["plain_sha256"]
    with tempfile.TemporaryFile() as src:                          # any failure -> EFAULT (src/main.rs:467)
        src.write(b"this is not zstd"); src.flush(); src.seek(0)
        with pytest.raises(OSError) as e:
            stream.decode_all(src)
        assert e.value.errno == 14
    with tempfile.TemporaryFile() as src:                          # empty input decodes to empty output
        assert stream.decode_all(src) == b""


def test_large_window_long_distance_and_level19(ref, corpus, exec_w):
    """config 4 in miniature: windowLog 23 (8 MiB window, offsets across dozens of blocks), a level-19 frame (block
    splitting, Repeat modes, Treeless literals) and a long-distance repeat; one frame each, so one LZ77 chain each"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    j = corpus.json_file(4242, 24 << 20).tobytes()
    far = j[: 6 << 20] + j[: 6 << 20]                                   # second half matches 6 MiB back
    blobs = [ref.writer_encode(j, 3, window_log=23), ref.writer_encode(far, 3, window_log=23), ref.writer_encode(j[: 3 << 20], 19)]
    plains = [j, far, j[: 3 << 20]]
    assert blobs[0][4] & 0x20 == 0                                      # not Single_Segment: a Window_Descriptor is present
    res = codec.decode_batch(blobs, [len(p) for p in plains])
    for i, ((st, out), plain) in enumerate(zip(res, plains)):
        assert st == 0, (i, codec.strerror(st))
        assert hashlib.sha256(out).digest() == hashlib.sha256(plain).digest(), i


def test_true_level19_with_an_8_mib_window(ref, corpus):
    """config 4's real settings on a file large enough to use them: libzstd level 19 (btultra2: block splitting, Repeat modes,
    Treeless literals, long matches) with windowLog 23 on 10 MiB, so offsets reach 8 MiB back across ~130 blocks"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    j = corpus.json_file(191919, 10 << 20).tobytes()
    j = j[: 9 << 20] + j[: 1 << 20]                                    # the last MiB repeats the first: matches 9 MiB -> capped at the window
    blob = ref.writer_encode(j, 19, window_log=23)
    assert blob[4] & 0x20 == 0 and blob[5] == ((23 - 10) << 3)           # Window_Descriptor: 8 MiB
    (st, out), = codec.decode_batch([blob], [len(j)])
    assert st == 0, codec.strerror(st)
    assert hashlib.sha256(out).digest() == hashlib.sha256(j).digest()


def test_far_form_records(ref, exec_w):
    """sequences with more than 32 extra bits (long literal run + far offset + long match): stage A hands stage B the bit
    cursor instead of the bits (tests/test_emul.py::test_far_form_records checks that the vector really has such sequences)"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    import emul_util
    plains = [emul_util.far_offset_long_length_plain(seed) for seed in (11, 12, 13)]
    blobs = [ref.writer_encode(p, 19) for p in plains] + [ref.writer_encode(plains[0], 3)]
    plains.append(plains[0])
    res = codec.decode_batch(blobs, [len(p) for p in plains])
    for i, ((st, out), plain) in enumerate(zip(res, plains)):
        assert st == 0, (i, codec.strerror(st))
        assert out == plain, i


def test_untrusted_sizes_in_the_fd_entry_point(ref):
    """fzg_decode_fd buffers whole files, so nothing in the input may size an allocation beyond what its block headers can
    regenerate: a frame that declares 2^62 bytes fails that one open (src/main.rs:467: EFAULT) instead of taking the process
    down, and a highly compressible file WITHOUT Frame_Content_Size decodes in one attempt (bound = the block walk)."""
    import errno
    import struct
    import tempfile
    # FCS_flag 3 (8-byte field), Single_Segment, no checksum; one last Raw block of 5 bytes
    lying = struct.pack("<IB", 0xFD2FB528, 0xE0) + struct.pack("<Q", 1 << 62) + bytes([5 << 3 | 1, 0, 0]) + b"hello"
    st, content, _ = codec.frame_info(lying)
    assert st == codec.E_UNSUPPORTED or content == 1 << 62        # single segment: window = FCS > 2^27 is already refused by the header walk
    lying2 = struct.pack("<IB", 0xFD2FB528, 0xC0) + bytes([0x00]) + struct.pack("<Q", 1 << 62) + bytes([5 << 3 | 1, 0, 0]) + b"hello"   # windowed: 1 KiB window
    st, content, _ = codec.frame_info(lying2)
    assert st == 0 and content == 1 << 62
    for blob in (lying, lying2):
        with tempfile.TemporaryFile() as src, tempfile.TemporaryFile() as dst:
            src.write(blob); src.seek(0)
            with pytest.raises(OSError) as ei:
                codec.decode_fd(src.fileno(), dst.fileno())
            assert ei.value.errno == errno.EFAULT
            assert dst.seek(0, 2) == 0
    if not ref.available:
        pytest.skip("system libzstd absent")
    plain = bytes(48 << 20)                                       # zeros: ratio far beyond any fixed guess
    comp = ref.writer_encode(plain, 3, pledge=False)
    assert len(comp) * 1024 < len(plain) and codec.frame_info(comp)[1] is None
    with tempfile.TemporaryFile() as src, tempfile.TemporaryFile() as dst:
        src.write(comp); src.seek(0)
        assert codec.decode_fd(src.fileno(), dst.fileno()) == len(plain)
        dst.seek(0)
        assert hashlib.sha256(dst.read()).digest() == hashlib.sha256(plain).digest()


def test_guard_bands_around_every_output(golden, exec_w):
    """compute-sanitizer is closed on this pool (profiles/r02_compute_sanitizer_closed.txt), so the bounds are checked the
    plain way: every output of a device-resident batch (all golden frames + mutated copies of them, which fail half way
    through) sits between two 256-byte guard bands and is given exactly its capacity; no kernel may touch a guard."""
    import torch
    rng = np.random.default_rng(77)
    names = [n for n in sorted(golden) if golden[n][1]["plain_len"] > 0]
    blobs, caps = [], []
    for n in names:
        comp, meta = golden[n]
        blobs.append(bytes(comp)); caps.append(meta["plain_len"])
        for _ in range(3):                                        # corrupted twins: wrong sizes, bad offsets, broken streams
            b = bytearray(comp)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(4, len(b)))] ^= 1 << int(rng.integers(0, 8))
            blobs.append(bytes(b)); caps.append(meta["plain_len"])
    G = 256
    offs, tot = [], 0
    for c in caps:
        offs.append(tot + G); tot += G + ((c + 255) & ~255) + G
    d_dst = torch.full((tot,), 0xA5, dtype=torch.uint8, device="cuda")
    srcs = [torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda() for b in blobs]
    torch.cuda.synchronize()
    dl, st = codec.decode_batch_ptrs(0, [s.data_ptr() for s in srcs], [s.numel() for s in srcs],
                                     [d_dst.data_ptr() + o for o in offs], caps, codec.SRC_DEVICE | codec.DST_DEVICE)
    host = d_dst.cpu().numpy()
    for i, (o, c) in enumerate(zip(offs, caps)):
        assert (host[o - G:o] == 0xA5).all(), "wrote before output %d" % i
        assert (host[o + c:o + ((c + 255) & ~255) + G] == 0xA5).all(), "wrote past the capacity of output %d" % i
    for k, n in enumerate(names):                                 # the intact ones are still right
        i = 4 * k
        assert st[i] == 0 and dl[i] == caps[i] and sha(host[offs[i]:offs[i] + caps[i]].tobytes()) == golden[n][1]["plain_sha256"], n


def test_large_frames_go_home_behind_the_execute_stage(ref, corpus):
    """few large frames into host buffers (SURVEY 8 config 4 in small): the device -> host copy follows the frames' progress
    words while they are still executed; every byte compared, and what the streaming must not break: an item of several
    frames, a bad trailer (reported, the other items intact), a small item, a capacity larger than the content"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    big = [corpus.json_file(5100 + i, (20 << 20) + 4097 * i).tobytes() for i in range(3)]
    blobs = [ref.writer_encode(b, 3, window_log=23) for b in big]
    multi_plain = corpus.json_file(5200, 9 << 20).tobytes()
    blobs.append(b"".join(ref.writer_encode(multi_plain[o:o + (3 << 20) + 5], 3) for o in range(0, len(multi_plain), (3 << 20) + 5)))
    bad = bytearray(blobs[1]); bad[-2] ^= 0x40
    blobs.append(bytes(bad))
    small = corpus.json_file(5300, 1000).tobytes()
    blobs.append(ref.writer_encode(small, 3))
    plains = big + [multi_plain, None, small]
    caps = [len(big[0]), len(big[1]) + 12345, len(big[2]), len(multi_plain), len(big[1]), 4096]
    old = os.environ.get("FZG_STREAM_OUT_MB")
    try:
        os.environ["FZG_STREAM_OUT_MB"] = "16"
        before = codec.streamed_copies(0)
        res = codec.decode_batch(blobs, caps)
        streamed = codec.streamed_copies(0) - before
        os.environ["FZG_STREAM_OUT_MB"] = "1000000"
        res_plain = codec.decode_batch(blobs, caps)
        assert codec.streamed_copies(0) - before == streamed
    finally:
        if old is None:
            os.environ.pop("FZG_STREAM_OUT_MB", None)
        else:
            os.environ["FZG_STREAM_OUT_MB"] = old
    assert streamed >= 1, streamed                    # pieces of at least 4 MiB of the 20 MiB frames, queued while the chains ran (how many: timing)
    for r in (res, res_plain):
        for i, (st, out) in enumerate(r):
            if plains[i] is None:
                assert st == codec.E_CHECKSUM, (i, codec.strerror(st))
            else:
                assert st == 0 and out == plains[i], i


def test_pinned_sources_in_several_allocations(ref, corpus):
    """host-resident batches whose items live in pinned memory: one span is copied as it is, items scattered over several pinned
    allocations (the cache's merged readahead batches) are copied one by one, and items that look like one span but are two
    allocations must not be taken for one (a span ends at a gap of more than 256 KiB; a failed span copy falls back to copies
    per item).  Pinned and pageable destinations, results compared byte for byte."""
    if not ref.available:
        pytest.skip("system libzstd absent")
    import torch
    n = 24
    plains = [corpus.json_file(6100 + i, 150000 + 1237 * i).tobytes() for i in range(n)]
    blobs = [ref.writer_encode(b, 3) for b in plains]
    bufs = [torch.empty(8 << 20, dtype=torch.uint8).pin_memory() for _ in range(3)]
    fill = [0, 0, 0]
    sp, sl = [], []
    for i, b in enumerate(blobs):                       # round-robin over the three allocations: never one increasing span
        k = i % 3
        a = bufs[k].numpy()
        a[fill[k]:fill[k] + len(b)] = np.frombuffer(b, dtype=np.uint8)
        sp.append(a.ctypes.data + fill[k]); sl.append(len(b))
        fill[k] += (len(b) + 63) & ~63
    caps = [len(p_) for p_ in plains]
    out_pinned = torch.empty(sum(caps) + 64 * n, dtype=torch.uint8).pin_memory().numpy()
    out_paged = np.zeros(sum(caps) + 64 * n, dtype=np.uint8)
    for out in (out_pinned, out_paged):
        dp, o = [], 0
        for c in caps:
            dp.append(out.ctypes.data + o); o += c + 64
        dl, st = codec.decode_batch_ptrs(0, sp, sl, dp, caps, 0)
        assert not st.any() and (dl == np.array(caps, dtype=np.uint64)).all()
        o = 0
        for i, c in enumerate(caps):
            assert out[o:o + c].tobytes() == plains[i], i
            o += c + 64
    # increasing addresses across two allocations: sorted by address, the items of buffer A then those of buffer B
    order = sorted(range(n), key=lambda i: sp[i])
    dp, o = [], 0
    for i in order:
        dp.append(out_paged.ctypes.data + o); o += caps[i] + 64
    out_paged[:] = 0
    dl, st = codec.decode_batch_ptrs(0, [sp[i] for i in order], [sl[i] for i in order], dp, [caps[i] for i in order], 0)
    assert not st.any()
    o = 0
    for i in order:
        assert out_paged[o:o + caps[i]].tobytes() == plains[i], i
        o += caps[i] + 64
