"""The reference's mount-level flow tests (/root/reference/tests/cmdline.rs, glitches.rs; convert mode excluded), restated
against the fzfs host (fuse-zstd_b200/csrc/fzfs.cpp).

Two hosts are built from the same source:
  * fuse-zstd_b200/fzfs      the product: libfzgpu.so behind the codec boundary          -> `-m gpu`
  * oracle/_ref/fzfs_ref     TEST ONLY: the reference's libzstd calls behind the boundary -> runs here, on the CPU, and pins
                             the HOST LOGIC (namespace, handle table, sync flow, xattrs) to the byte strings the reference's
                             own tests pin (tests/cmdline.rs:34-43, 160-178)
Needs /dev/fuse and mount(2) (root); skipped otherwise.
"""
import importlib
import os
import subprocess
import sys
import tempfile
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_HOST = os.path.join(ROOT, "fuse-zstd_b200", "fzfs")
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "fzfs_ref")


def _can_mount():
    return os.path.exists("/dev/fuse") and os.geteuid() == 0


class Mount:
    def __init__(self, host, *extra):
        self.data = tempfile.mkdtemp(prefix="fzdata")
        self.mp = tempfile.mkdtemp(prefix="fzmnt")
        host, *opts = host.split(" ")                    # "path --threads 4": a host with options (see _hosts)
        self.proc = subprocess.Popen([host, "--data-dir", self.data, "--mount-point", self.mp, *opts, *extra], stderr=subprocess.PIPE)
        for _ in range(400):
            if os.path.ismount(self.mp) or self.proc.poll() is not None:
                break
            time.sleep(0.025)
        if not os.path.ismount(self.mp):
            err = self.proc.stderr.read().decode() if self.proc.poll() is not None else ""
            self.close()
            raise RuntimeError("mount failed: " + err)

    def close(self):
        if self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=10)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        subprocess.call(["umount", "-l", self.mp], stderr=subprocess.DEVNULL)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _hosts():
    # the serving threads: the reference's codec defaults to 1 (fuser's loop), the GPU's to 8; both hosts are run both ways so that the
    # locking of the multi-threaded loop is exercised here on the CPU too
    out = [pytest.param(REF_HOST, id="reference-codec"), pytest.param(REF_HOST + " --threads 4", id="reference-codec-4-threads")]
    out.append(pytest.param(GPU_HOST, id="gpu-codec", marks=pytest.mark.gpu))
    out.append(pytest.param(GPU_HOST + " --threads 1", id="gpu-codec-1-thread", marks=pytest.mark.gpu))
    return out


@pytest.fixture(scope="module", autouse=True)
def _build():
    if not _can_mount():
        pytest.skip("needs /dev/fuse and root")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "fuse-zstd_b200", "csrc"), "all"])


def _populate(mp):
    os.makedirs(os.path.join(mp, "first/second/third"))
    os.makedirs(os.path.join(mp, "first/second/empty"))
    for rel, text in (("file1.txt", b"1st file in root"), ("first/file1.txt", b"1st file in first"), ("first/file2.txt", b"2nd file in first"),
                      ("first/second/file1.txt", b"1st file in second"), ("first/second/file2.txt", b"2nd file in second"),
                      ("first/second/file3.txt", b"3rd file in second"), ("first/second/third/file1.txt", b"1st file in third")):
        with open(os.path.join(mp, rel), "wb") as fh:
            fh.write(text)


def _decoded(path, oracle):
    with open(path, "rb") as fh:
        blob = fh.read()
    st, out = oracle.decode(blob, cap=1 << 22)
    assert st == 0, path
    return out


@pytest.mark.parametrize("host", _hosts())
def test_touch_mkdir_ls_cat(host, oracle):
    """tests/cmdline.rs: touch (:33-43, the 13-byte frame), mkdir (:45-54), ls (:56-93), cat (:95-115)"""
    with Mount(host) as m:
        open(os.path.join(m.mp, "file.txt"), "wb").close()
        assert open(os.path.join(m.data, "file.txt.zst"), "rb").read() == bytes.fromhex("28b52ffd240001000099e9d851")
        os.mkdir(os.path.join(m.mp, "directory"))
        assert os.path.isdir(os.path.join(m.data, "directory"))
        os.unlink(os.path.join(m.mp, "file.txt")); os.rmdir(os.path.join(m.mp, "directory"))
        _populate(m.mp)
        assert sorted(os.listdir(m.mp)) == ["file1.txt", "first"]
        assert sorted(os.listdir(os.path.join(m.mp, "first"))) == ["file1.txt", "file2.txt", "second"]
        assert sorted(os.listdir(os.path.join(m.mp, "first/second"))) == ["empty", "file1.txt", "file2.txt", "file3.txt", "third"]
        assert os.listdir(os.path.join(m.mp, "first/second/empty")) == []
        assert sorted(os.listdir(os.path.join(m.data, "first"))) == ["file1.txt.zst", "file2.txt.zst", "second"]
        assert open(os.path.join(m.mp, "first/second/third/file1.txt"), "rb").read() == b"1st file in third"
        assert subprocess.check_output(["cat", os.path.join(m.mp, "first/file2.txt")]) == b"2nd file in first"
        assert _decoded(os.path.join(m.data, "first/second/file3.txt.zst"), oracle) == b"3rd file in second"
        with open(os.path.join(m.data, "plain.bin"), "wb") as fh:       # not a .zst file: hidden (src/main.rs:338-344)
            fh.write(b"x")
        assert "plain.bin" not in os.listdir(m.mp)
        assert os.stat(os.path.join(m.mp, "file1.txt")).st_size == len(b"1st file in root")


@pytest.mark.parametrize("host", _hosts())
def test_tee_truncate_append(host, oracle):
    """tests/cmdline.rs:117-179: overwrite, truncate and append through the mount; the last state is the 35-byte frame the
    reference writer produces for "truncated and appended" """
    with Mount(host) as m:
        p = os.path.join(m.mp, "file.txt")
        with open(p, "wb") as fh:
            fh.write(b"first version of the file")
        assert _decoded(os.path.join(m.data, "file.txt.zst"), oracle) == b"first version of the file"
        with open(p, "wb") as fh:                                         # O_TRUNC
            fh.write(b"truncated")
        assert open(p, "rb").read() == b"truncated"
        with open(p, "ab") as fh:
            fh.write(b" and appended")
        assert open(p, "rb").read() == b"truncated and appended"
        blob = open(os.path.join(m.data, "file.txt.zst"), "rb").read()
        assert _decoded(os.path.join(m.data, "file.txt.zst"), oracle) == b"truncated and appended"
        if host == REF_HOST:
            assert blob == bytes.fromhex("28b52ffd2416b100007472756e636174656420616e6420617070656e6465643d98a66b")
        xs = os.getxattr(os.path.join(m.data, "file.txt.zst"), "user.real_size") if _xattr_ok(m.data) else None
        assert xs in (None, (22).to_bytes(8, "big"))                      # src/main.rs:821: 8-byte big-endian


def _xattr_ok(d):
    try:
        os.setxattr(d, "user.fz_probe", b"1"); os.removexattr(d, "user.fz_probe")
        return True
    except OSError:
        return False


@pytest.mark.parametrize("host", _hosts())
def test_mv_rm_rmdir(host, oracle):
    """tests/cmdline.rs:181-290"""
    with Mount(host) as m:
        _populate(m.mp)
        mp, dd = m.mp, m.data
        os.rename(os.path.join(mp, "first/second/file1.txt"), os.path.join(mp, "first/second/fileI.txt"))
        assert os.path.exists(os.path.join(dd, "first/second/fileI.txt.zst")) and not os.path.exists(os.path.join(dd, "first/second/file1.txt.zst"))
        os.rename(os.path.join(mp, "first/second/file2.txt"), os.path.join(mp, "first/file3.txt"))
        assert _decoded(os.path.join(dd, "first/file3.txt.zst"), oracle) == b"2nd file in second"
        os.rename(os.path.join(mp, "first/file1.txt"), os.path.join(mp, "first/second/third/file1.txt"))       # onto an existing file
        assert not os.path.exists(os.path.join(dd, "first/file1.txt.zst"))
        assert _decoded(os.path.join(dd, "first/second/third/file1.txt.zst"), oracle) == b"1st file in first"
        assert open(os.path.join(mp, "first/second/third/file1.txt"), "rb").read() == b"1st file in first"
        os.rename(os.path.join(mp, "first/second/empty"), os.path.join(mp, "first/second/void"))
        assert os.path.isdir(os.path.join(dd, "first/second/void")) and not os.path.exists(os.path.join(dd, "first/second/empty"))
        subprocess.check_call(["mv", os.path.join(mp, "first/second"), mp])
        assert os.path.isdir(os.path.join(dd, "second/void")) and not os.path.exists(os.path.join(dd, "first/second"))
        assert subprocess.call(["mv", "-T", os.path.join(mp, "second/third"), os.path.join(mp, "second/file3.txt")], stderr=subprocess.DEVNULL) != 0
        assert os.path.isdir(os.path.join(dd, "second/third"))
        os.unlink(os.path.join(mp, "second/file3.txt"))
        assert not os.path.exists(os.path.join(dd, "second/file3.txt.zst"))
        with pytest.raises(OSError):
            os.rmdir(os.path.join(mp, "second/third"))                    # not empty
        os.unlink(os.path.join(mp, "second/third/file1.txt")); os.rmdir(os.path.join(mp, "second/third"))
        assert not os.path.exists(os.path.join(dd, "second/third"))


@pytest.mark.parametrize("host", _hosts())
def test_parallel_write_append_flush(host, oracle):
    """tests/glitches.rs:19-91, 196-234: several handles on one file behave as on a plain directory; every close of a
    written handle re-encodes the whole file"""
    def parallel_write(d):
        p = os.path.join(d, "file.txt")
        f1 = os.open(p, os.O_WRONLY | os.O_CREAT, 0o644); f2 = os.open(p, os.O_WRONLY | os.O_CREAT, 0o644); f3 = os.open(p, os.O_WRONLY | os.O_CREAT, 0o644)
        os.write(f2, b"SECOND"); os.write(f1, b"FIRST"); os.fsync(f1); os.close(f1)
        os.write(f3, b"THIRD"); os.fsync(f2); os.close(f2); os.fsync(f3); os.close(f3)
        return open(p, "rb").read()

    def append(d):
        p = os.path.join(d, "app.txt")
        with open(p, "wb") as fh:
            fh.write(b"BASIC")
        with open(p, "ab") as fh:
            fh.write(b"APPENDED"); fh.flush(); os.fsync(fh.fileno())
        return open(p, "rb").read()

    with Mount(host) as m, tempfile.TemporaryDirectory() as plain:
        assert parallel_write(m.mp) == parallel_write(plain)
        assert append(m.mp) == append(plain) == b"BASICAPPENDED"
        assert _decoded(os.path.join(m.data, "app.txt.zst"), oracle) == b"BASICAPPENDED"
        p = os.path.join(m.mp, "flush.txt")                               # glitches.rs:196-234
        fd = os.open(p, os.O_WRONLY | os.O_CREAT, 0o644)
        os.write(fd, b"KEEPIT"); d2 = os.dup(fd); os.close(d2)            # closing a dup flushes
        assert _decoded(os.path.join(m.data, "flush.txt.zst"), oracle) == b"KEEPIT"
        os.pwrite(fd, b"OVERRIDE", 0); os.close(fd)
        assert _decoded(os.path.join(m.data, "flush.txt.zst"), oracle) == b"OVERRIDE"


@pytest.mark.parametrize("host", _hosts())
def test_files_written_by_stock_zstd_and_unlink_while_open(host, ref, corpus, oracle):
    """a frame written by stock zstd (no checksum, tests/convert.rs:16-25) is served; a larger level-3 file reads back
    bit-exactly in 128 KiB reads; a file unlinked while open is not written back (src/file.rs:119-127); a corrupt file
    fails open() with EFAULT (src/main.rs:467)"""
    if not ref.available:
        pytest.skip("system libzstd absent")
    big = corpus.json_file(2024, 3 << 20).tobytes()
    with Mount(host) as m:
        with open(os.path.join(m.data, "stock.zst"), "wb") as fh:
            fh.write(bytes.fromhex("28b52ffd200f790000636f6d707265737365642064617461"))
        with open(os.path.join(m.data, "big.json.zst"), "wb") as fh:
            fh.write(ref.writer_encode(big, 3))
        with open(os.path.join(m.data, "bad.zst"), "wb") as fh:
            fh.write(b"this is not zstd")
        assert open(os.path.join(m.mp, "stock"), "rb").read() == b"compressed data"
        assert os.stat(os.path.join(m.mp, "big.json")).st_size == len(big)   # from the frame header: no xattr yet
        got = bytearray()
        with open(os.path.join(m.mp, "big.json"), "rb", buffering=0) as fh:
            while True:
                b = fh.read(128 << 10)
                if not b:
                    break
                got += b
        assert bytes(got) == big
        with pytest.raises(OSError) as e:
            open(os.path.join(m.mp, "bad"), "rb")
        assert e.value.errno == 14
        p = os.path.join(m.mp, "gone.txt")
        fd = os.open(p, os.O_WRONLY | os.O_CREAT, 0o644)
        os.write(fd, b"never stored"); os.unlink(p); os.close(fd)
        assert not os.path.exists(os.path.join(m.data, "gone.txt.zst"))


@pytest.mark.parametrize("host", _hosts())
def test_readers_served_in_place_then_a_writer_joins(host, ref, corpus, oracle):
    """A read-only open of a file the readahead has decoded is served from the cache in place (GPU host; the CPU host takes the
    tmpfile path for the same calls).  The reference's handle semantics must survive it: every handle of an inode sees one
    file (src/file.rs:67-102) -- a second reader shares the bytes, a writer that joins moves all handles to one tmpfile, its
    writes are visible to the readers, fsync of a read-only handle re-encodes (src/main.rs:709-712), truncation reaches them."""
    if not ref.available:
        pytest.skip("system libzstd absent")
    files = {("d/f%03d" % i): corpus.json_file(880000 + i, 300000 + 1000 * i).tobytes() for i in range(6)}
    with Mount(host) as m:
        os.makedirs(os.path.join(m.data, "d"))
        for rel, plain in files.items():
            with open(os.path.join(m.data, rel + ".zst"), "wb") as fh:
                fh.write(ref.writer_encode(plain, 3))
        p = os.path.join(m.mp, "d/f002"); plain = files["d/f002"]
        r1 = os.open(p, os.O_RDONLY)
        assert os.pread(r1, 4096, 1000) == plain[1000:5096]
        r2 = os.open(p, os.O_RDONLY)                                       # shares the bytes
        assert os.pread(r2, 100, len(plain) - 50) == plain[-50:]
        assert os.fstat(r1).st_size == len(plain)
        w = os.open(p, os.O_RDWR)                                          # a writer joins
        os.pwrite(w, b"PATCHED!", 2000)
        assert os.pread(r1, 8, 2000) == b"PATCHED!" and os.pread(r2, 8, 2000) == b"PATCHED!"
        os.close(w)                                                        # release of a written handle: whole-file encode
        want = plain[:2000] + b"PATCHED!" + plain[2008:]
        st, back = oracle.decode(open(os.path.join(m.data, "d/f002.zst"), "rb").read(), cap=len(want))
        assert st == 0 and back == want
        os.close(r1); os.close(r2)
        assert open(p, "rb").read() == want
        # fsync of a read-only handle that is served in place: forced re-encode, same content afterwards
        q = os.path.join(m.mp, "d/f004")
        r = os.open(q, os.O_RDONLY)
        before = os.stat(os.path.join(m.data, "d/f004.zst")).st_ino
        os.fsync(r)
        assert os.stat(os.path.join(m.data, "d/f004.zst")).st_ino != before     # atomic rename: a new backing inode
        assert os.pread(r, 1 << 20, 0) == files["d/f004"]
        os.close(r)
        st, back = oracle.decode(open(os.path.join(m.data, "d/f004.zst"), "rb").read(), cap=len(files["d/f004"]))
        assert st == 0 and back == files["d/f004"]
        # truncation while a reader holds the file
        t = os.path.join(m.mp, "d/f005")
        r = os.open(t, os.O_RDONLY)
        assert os.pread(r, 10, 0) == files["d/f005"][:10]
        os.truncate(t, 1234)
        assert os.fstat(r).st_size == 1234 and os.pread(r, 5000, 0) == files["d/f005"][:1234]
        os.close(r)


@pytest.mark.parametrize("host", _hosts())
def test_many_readers_beside_writers(host, ref, corpus, oracle):
    """the serving threads (Fs::loop): eight readers walk a directory of pre-compressed files with random preads while two
    writers create, append to, rewrite, rename and delete files of their own in the same directory; every byte read is compared,
    every stored file is decoded by the oracle at the end.  With one serving thread this is the reference's behaviour, with
    several the READs overlap each other and wait for everything else."""
    if not ref.available:
        pytest.skip("system libzstd absent")
    import random
    import threading
    files = {("f%03d" % i): corpus.json_file(990000 + i, 200000 + 4097 * i).tobytes() for i in range(12)}
    with Mount(host) as m:
        os.makedirs(os.path.join(m.data, "d"))
        for rel, plain in files.items():
            with open(os.path.join(m.data, "d", rel + ".zst"), "wb") as fh:
                fh.write(ref.writer_encode(plain, 3))
        errors = []

        def reader(seed):
            try:
                rnd = random.Random(seed)
                for _ in range(12):
                    name = rnd.choice(sorted(files)); plain = files[name]
                    fd = os.open(os.path.join(m.mp, "d", name), os.O_RDONLY)
                    try:
                        assert os.fstat(fd).st_size == len(plain)
                        for _ in range(20):
                            o = rnd.randrange(len(plain)); n = rnd.choice((1, 4096, 65536, 131072))
                            assert os.pread(fd, n, o) == plain[o:o + n], (name, o, n)
                    finally:
                        os.close(fd)
            except Exception as e:                                        # noqa: BLE001 -- reported by the main thread
                errors.append(repr(e))

        stored = {}

        def writer(k):
            try:
                body = corpus.json_file(995000 + k, 150000).tobytes()
                for r in range(4):
                    p = os.path.join(m.mp, "d", "w%d_%d" % (k, r))
                    with open(p, "wb") as fh:
                        fh.write(body[:100000])
                    with open(p, "ab") as fh:
                        fh.write(body[100000:])
                    with open(p, "r+b") as fh:
                        fh.seek(5000); fh.write(b"REWRITTEN")
                    want = body[:5000] + b"REWRITTEN" + body[5009:]
                    assert open(p, "rb").read() == want
                    if r % 2:
                        q = p + "_moved"; os.rename(p, q); stored["w%d_%d_moved" % (k, r)] = want
                    elif r == 2:
                        os.unlink(p)
                    else:
                        stored["w%d_%d" % (k, r)] = want
            except Exception as e:                                        # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=reader, args=(s,)) for s in range(8)] + [threading.Thread(target=writer, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors[:3]
        for name, want in stored.items():
            st, back = oracle.decode(open(os.path.join(m.data, "d", name + ".zst"), "rb").read(), cap=len(want))
            assert st == 0 and back == want, name
        assert sorted(n for n in os.listdir(os.path.join(m.mp, "d"))) == sorted(list(files) + list(stored))


def test_product_host_fails_loudly_without_a_gpu():
    """no CPU fallback: the product host refuses to mount when the CUDA codec cannot start"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with tempfile.TemporaryDirectory() as d, tempfile.TemporaryDirectory() as mp:
        r = subprocess.run([GPU_HOST, "--data-dir", d, "--mount-point", mp], stderr=subprocess.PIPE, timeout=60)
        assert r.returncode != 0 and b"codec unavailable" in r.stderr
        assert not os.path.ismount(mp)
