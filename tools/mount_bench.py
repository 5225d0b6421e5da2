"""Read throughput THROUGH THE MOUNT (BASELINE.json: "fio read MB/s via mount"): the job shape of
/root/reference/benchmarks/parallel-files.fio (J jobs, each its own directory of nrfiles files of 1 MiB, opened one at a time
and read with 128 KiB reads, psync) without fio, which this image does not have.  Two arms, same data directory, same
reader processes, same fzfs host source (fuse-zstd_b200/csrc/fzfs.cpp):
  gpu        fuse-zstd_b200/fzfs      GPU codec, directory readahead + decoded-file cache
  reference  oracle/_ref/fzfs_ref     the reference's libzstd calls on the host's one FUSE thread (the restated CPU path; the
                                      unmodified fuse-zstd binary cannot be built here: no Rust toolchain)
An arm may name its serving threads: gpu:1 (one request at a time, as fuser), gpu:16; plain `gpu` / `reference` use the hosts' defaults (8 / 1).
usage: mount_bench.py [--jobs 16] [--nrfiles 125] [--filesize-kib 1024] [--arms reference,gpu:1,gpu]"""
import argparse, importlib, json, multiprocessing as mp, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def natural_key(name):
    import re
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def reader(args):
    d, bs = args
    n = 0
    for name in sorted(os.listdir(d), key=natural_key):          # file index order, as a fio job walks its nrfiles
        fd = os.open(os.path.join(d, name), os.O_RDONLY)
        while True:
            b = os.read(fd, bs)
            if not b:
                break
            n += len(b)
        os.close(fd)
    return n


def writer(args):
    """write-and-verify.fio in small: every file written sequentially with `bs` writes and closed (close = whole-file encode,
    /root/reference/src/main.rs:595-599), then read back and compared"""
    d, bs, nfiles, size, seed = args
    import importlib
    corpus = importlib.import_module("fuse-zstd_b200.corpus")
    plain = corpus.json_files(seed, nfiles, size, threads=1)
    n = 0
    for i in range(nfiles):
        body = plain[i].tobytes()
        p = os.path.join(d, "w%04d" % i)
        fd = os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        for o in range(0, size, bs):
            os.write(fd, body[o:o + bs])
        os.close(fd)
        with open(p, "rb") as fh:
            assert fh.read() == body, p
        n += size
    return n


def writer_fio(args):
    """/root/reference/benchmarks/write-and-verify.fio as written: rw=randwrite bs=4k size=100m nrfiles=5 (20 MB per file), every
    4 KiB block of a file written exactly once in random order, the file closed (= whole-file encode on release), then read back
    and compared (fio: verify=crc32c).  iodepth / direct / libaio have no meaning through a one-thread FUSE loop."""
    d, nfiles, size, seed = args
    import importlib
    import numpy as np
    corpus = importlib.import_module("fuse-zstd_b200.corpus")
    plain = corpus.json_files(seed, nfiles, size, threads=1)
    rs = np.random.RandomState(seed)
    n = 0
    for i in range(nfiles):
        body = plain[i].tobytes()
        p = os.path.join(d, "v%04d" % i)
        fd = os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        for blk in rs.permutation((size + 4095) // 4096):
            os.pwrite(fd, body[blk * 4096:(blk + 1) * 4096], int(blk) * 4096)
        os.close(fd)
        with open(p, "rb") as fh:
            assert fh.read() == body, p
        n += size
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=16); ap.add_argument("--nrfiles", type=int, default=125)
    ap.add_argument("--filesize-kib", type=int, default=1024); ap.add_argument("--arms", default="reference,gpu")
    ap.add_argument("--cache-mb", type=int, default=3072)
    ap.add_argument("--write-files", type=int, default=8, help="files per job in the write-and-verify phase (0: skip)")
    ap.add_argument("--write-mib", type=int, default=20)
    ap.add_argument("--fio-write", action="store_true", help="also run write-and-verify.fio's own shape: 5 jobs x 5 files x 20 MB, 4 KiB random writes")
    a = ap.parse_args()
    import pyoracle
    corpus = importlib.import_module("fuse-zstd_b200.corpus")
    pyoracle.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "fuse-zstd_b200", "csrc"), "all"])
    R = pyoracle.Ref(); assert R.available
    size = a.filesize_kib << 10
    data = tempfile.mkdtemp(prefix="fzbench_data")
    t0 = time.time()
    for j in range(a.jobs):
        os.mkdir(os.path.join(data, "job%02d" % j))
        plain = corpus.json_files(3300000 + j * a.nrfiles, a.nrfiles, size, threads=os.cpu_count())
        for i in range(a.nrfiles):
            with open(os.path.join(data, "job%02d" % j, "f%05d.zst" % i), "wb") as fh:
                fh.write(R.writer_encode(plain[i].tobytes(), 3))
    total = a.jobs * a.nrfiles * size
    print("corpus: %d jobs x %d files x %d KiB = %.2f GB plain, built in %.1f s" % (a.jobs, a.nrfiles, a.filesize_kib, total / 1e9, time.time() - t0), file=sys.stderr)
    hosts = {"gpu": [os.path.join(ROOT, "fuse-zstd_b200", "fzfs"), "--cache-mb", str(a.cache_mb)], "reference": [os.path.join(ROOT, "oracle", "_ref", "fzfs_ref")]}
    out = {"metric": "mount_read_MBps", "unit": "MB/s", "jobs": a.jobs, "nrfiles": a.nrfiles, "filesize": size, "bs": 131072,
           "workload": "benchmarks/parallel-files.fio shape (jobs x nrfiles x filesize, one open file per job) + sequential 128 KiB reads of every file"}
    for arm_spec in a.arms.split(","):
        arm, _, nthr = arm_spec.partition(":")
        mpnt = tempfile.mkdtemp(prefix="fzbench_mnt")
        proc = subprocess.Popen(hosts[arm] + (["--threads", nthr] if nthr else []) + os.environ.get("FZFS_EXTRA_ARGS", "").split() + ["--data-dir", data, "--mount-point", mpnt])
        arm = arm if not nthr else "%s_%s_threads" % (arm, nthr)
        for _ in range(2400):                         # the GPU host allocates its pinned cache before it mounts
            if os.path.ismount(mpnt) or proc.poll() is not None:
                break
            time.sleep(0.025)
        assert os.path.ismount(mpnt), arm
        try:
            dirs = [(os.path.join(mpnt, "job%02d" % j), 131072) for j in range(a.jobs)]
            with mp.Pool(a.jobs) as pool:
                t0 = time.perf_counter()
                got = sum(pool.map(reader, dirs))
                dt = time.perf_counter() - t0
            assert got == total, (arm, got, total)
            out[arm] = round(total / 1e6 / dt, 1)
            print("%s: %.2f GB through the mount in %.2f s -> %.1f MB/s" % (arm, total / 1e9, dt, total / 1e6 / dt), file=sys.stderr)
            if a.write_files:
                wjobs = min(a.jobs, 5)                                # write-and-verify.fio: 5 workers
                for j in range(wjobs):
                    os.mkdir(os.path.join(mpnt, "wjob%02d" % j))
                wargs = [(os.path.join(mpnt, "wjob%02d" % j), 131072, a.write_files, a.write_mib << 20, 4400000 + 1000 * j) for j in range(wjobs)]
                with mp.Pool(wjobs) as pool:
                    t0 = time.perf_counter()
                    wrote = sum(pool.map(writer, wargs))
                    dt = time.perf_counter() - t0
                out[arm + "_write_verify"] = round(wrote / 1e6 / dt, 1)
                print("%s: wrote + verified %.2f GB through the mount in %.2f s -> %.1f MB/s" % (arm, wrote / 1e9, dt, wrote / 1e6 / dt), file=sys.stderr)
                for j in range(wjobs):
                    shutil.rmtree(os.path.join(data, "wjob%02d" % j), ignore_errors=True)
            if a.fio_write:
                for j in range(5):
                    os.mkdir(os.path.join(mpnt, "vjob%02d" % j))
                vargs = [(os.path.join(mpnt, "vjob%02d" % j), 5, 20 * 1000 * 1000, 4500000 + 1000 * j) for j in range(5)]
                with mp.Pool(5) as pool:
                    t0 = time.perf_counter()
                    wrote = sum(pool.map(writer_fio, vargs))
                    dt = time.perf_counter() - t0
                out[arm + "_fio_write_and_verify"] = round(wrote / 1e6 / dt, 1)
                print("%s: write-and-verify.fio shape (5 x 5 x 20 MB, 4 KiB random writes): %.2f GB in %.2f s -> %.1f MB/s" % (arm, wrote / 1e9, dt, wrote / 1e6 / dt), file=sys.stderr)
                for j in range(5):
                    shutil.rmtree(os.path.join(data, "vjob%02d" % j), ignore_errors=True)
        finally:
            proc.terminate(); proc.wait(timeout=20)
            subprocess.call(["umount", "-l", mpnt], stderr=subprocess.DEVNULL)
    shutil.rmtree(data, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
