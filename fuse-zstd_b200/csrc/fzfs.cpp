/*
 * fzfs.cpp -- a raw-/dev/fuse host that restates fuse-zstd's filesystem surface in C++ so that the
 * GPU codec can be exercised and measured THROUGH A MOUNT in an environment without Rust, libfuse or fio.
 *
 * What it restates (paths relative to /root/reference):
 *   the Filesystem callbacks            src/main.rs:835-1207    -> Fs::dispatch (--threads 1: one request at a time, as fuser does;
 *                                                               more: READs of different requests run side by side, see Fs::loop)
 *   lookup / readdir / getattr          src/main.rs:215-405     `.zst` suffix added / stripped, other regular files hidden,
 *                                                               size = xattr user.real_size (8-byte BE), perms 0666 / 0777
 *   open (decode on first open)         src/main.rs:451-493     unlinked tmpfile, codec call, user.real_size written, fsync;
 *                                                               a file that is already open is dup'ed, not decoded again
 *   read / write / truncate             src/main.rs:495-513, 558-593, 408-449
 *   flush / fsync / release (encode)    src/main.rs:174-213, 755-832   temp file in the target directory, codec call,
 *                                                               user.ino kept, atomic rename, user.real_size, fsync
 *   create / mkdir / unlink / rmdir / rename   src/main.rs:515-555, 601-707
 *   handle table                        src/file.rs             fh -> {flags, needs_sync, tmpfile, refs}; unlink clears refs
 *   inode numbers                       src/main.rs:719-753     own counter counting down from 2^64-1, kept in the xattr user.ino
 *                                                               of every entry (an in-memory map when the data directory's
 *                                                               filesystem has no user xattrs)
 * What it does not restate: the sled inode cache (a hash map here), --convert mode, logging / Sentry, the CLI beyond the
 * three options the tests use.
 *
 * The codec sits behind four C functions (fzfs_codec.h): the product binary links fzfs_codec_gpu.cpp (libfzgpu.so: cached /
 * batched decode, GPU encode, directory readahead), the measurement baseline links oracle/fzfs_codec_ref.c (the reference's
 * libzstd calls, test infrastructure).  There is no fallback from one to the other.
 */
#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <linux/fuse.h>
#include <sched.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mount.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <sys/uio.h>
#include <sys/xattr.h>
#include <unistd.h>

#include <time.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "fzfs_codec.h"

namespace {

constexpr uint64_t kTtlSec = 1;                 // dcache lifetime (src/main.rs:25)
bool g_verbose = false;
// --stats: per opcode, how many requests, the time they took in the host and (of that) the time spent waiting for the host's lock
// or for a readahead batch; printed when the mount goes away
bool g_stats = false;
struct OpStat { std::atomic<uint64_t> n{ 0 }, ns{ 0 }, wait_ns{ 0 }, max_ns{ 0 }; };
OpStat g_op[64];
thread_local uint64_t t_wait_ns = 0;
inline uint64_t now_ns() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }
void logf(const char* fmt, ...)
{
    if (!g_verbose) return;
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr);
}

bool ends_with(const std::string& s, const char* suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
std::string dir_of(const std::string& p) { const size_t k = p.rfind('/'); return k == std::string::npos ? "." : p.substr(0, k); }

uint64_t be64(const uint8_t* b) { uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | b[i]; return v; }
void put_be64(uint8_t* b, uint64_t v) { for (int i = 7; i >= 0; i--) { b[i] = (uint8_t)v; v >>= 8; } }

// Sum of Frame_Content_Size over the frames of a .zst file, by walking frame and block headers with pread (no decoding):
// used when user.real_size is absent, so that a file does not look empty before its first open (README.md:20-23 of the
// reference lists that as a limitation).  Returns false when a frame omits the field or the file is malformed.
bool content_size_from_headers(int fd, uint64_t* out)
{
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    const uint64_t n = (uint64_t)st.st_size;
    uint64_t ip = 0, total = 0;
    uint8_t h[18];
    while (ip < n) {
        const ssize_t got = pread(fd, h, sizeof h, (off_t)ip);
        if (got < 8) return false;
        const uint32_t magic = h[0] | (h[1] << 8) | (h[2] << 16) | ((uint32_t)h[3] << 24);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) { ip += 8 + (uint64_t)(h[4] | (h[5] << 8) | (h[6] << 16) | ((uint32_t)h[7] << 24)); continue; }
        if (magic != 0xFD2FB528u) return false;
        const uint32_t fhd = h[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
        uint32_t pos = 5 + (single ? 0 : 1) + (did == 3 ? 4 : did);
        const uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
        if (fb == 0 || (ssize_t)(pos + fb) > got) return false;
        uint64_t fcs = 0;
        for (uint32_t i = 0; i < fb; i++) fcs |= (uint64_t)h[pos + i] << (8 * i);
        if (fb == 2) fcs += 256;
        total += fcs; ip += pos + fb;
        for (;;) {                                               // block headers
            uint8_t b[3];
            if (pread(fd, b, 3, (off_t)ip) != 3) return false;
            const uint32_t bh = b[0] | (b[1] << 8) | (b[2] << 16), type = (bh >> 1) & 3, size = bh >> 3;
            ip += 3 + (type == 1 ? 1 : size);
            if (bh & 1) break;
        }
        if (fhd & 4) ip += 4;
    }
    *out = total;
    return ip == n;
}

struct Handle {
    int flags; bool needs_sync; int fd; bool has_refs; uint64_t ino; std::string path;
    // READs in flight on this handle: they copy out of the backing WITHOUT the host's lock, so whoever replaces or closes the backing
    // waits for this to reach zero first (holding the lock exclusively, which keeps new READs from starting)
    std::shared_ptr<std::atomic<int>> busy = std::make_shared<std::atomic<int>>(0);
    // a read-only open served from the decoded-file cache IN PLACE (fd == -1): no tmpfile until somebody writes, truncates or fsyncs
    const uint8_t* view = nullptr; uint64_t view_size = 0; void* pin = nullptr;
};

class Fs {
public:
    Fs(std::string data_dir, int level, bool readahead) : data_(std::move(data_dir)), level_(level), readahead_(readahead)
    {
        uint8_t b[8];
        if (getxattr(data_.c_str(), "user.ino_idx", b, 8) == 8) ino_idx_ = be64(b);
    }
    int dev = -1;
    int n_threads = 1;
    void loop(int threads);

private:
    std::string data_; int level_; bool readahead_;
    uint64_t ino_idx_ = UINT64_MAX;
    bool xattr_ok_ = true;
    std::unordered_map<uint64_t, std::string> paths_;            // ino -> path in the data directory (the reference: sled)
    std::unordered_map<std::string, uint64_t> mem_ino_;          // path -> ino when user xattrs are not available
    std::unordered_map<std::string, uint64_t> mem_size_;         // path -> real size, same case
    std::unordered_map<uint64_t, Handle> handles_;
    std::unordered_map<uint64_t, std::unordered_set<uint64_t>> by_ino_;
    // readahead window per directory: the .zst entries in natural order, which of them were handed to the codec
    struct DirRa { std::vector<std::string> names; std::vector<uint8_t> asked; std::unordered_map<std::string, size_t> index; int64_t mtime_ns = -1; };
    std::unordered_map<std::string, DirRa> ra_;
    // Several threads may serve requests (Fs::loop).  Everything that changes the tables above, touches the data directory or calls
    // the codec takes mu_ exclusively -- the reference's one-request-at-a-time semantics, unchanged; a READ only looks a handle up and
    // copies bytes out of its tmpfile or its in-place view, so READs take mu_ shared and run side by side.
    std::shared_mutex mu_;
    std::atomic<int> parked_{ 0 };                               // OPEN requests waiting for their readahead batch on threads of their own
    void serve();

    // ---- inode numbers (src/main.rs:719-753)
    uint64_t next_ino()
    {
        const uint64_t r = ino_idx_;
        if (ino_idx_ - 1 <= FUSE_ROOT_ID) ino_idx_ = UINT64_MAX;
        ino_idx_--;
        uint8_t b[8]; put_be64(b, ino_idx_);
        if (xattr_ok_ && setxattr(data_.c_str(), "user.ino_idx", b, 8, 0) != 0 && (errno == ENOTSUP || errno == EPERM)) xattr_ok_ = false;
        return r;
    }
    uint64_t ino_of(const std::string& path)
    {
        uint8_t b[8];
        if (xattr_ok_) {
            if (getxattr(path.c_str(), "user.ino", b, 8) == 8) return be64(b);
            if (errno == ENOTSUP) xattr_ok_ = false;
        }
        if (!xattr_ok_) { auto it = mem_ino_.find(path); if (it != mem_ino_.end()) return it->second; }
        const uint64_t ino = next_ino();
        put_be64(b, ino);
        if (!xattr_ok_ || setxattr(path.c_str(), "user.ino", b, 8, 0) != 0) { if (xattr_ok_ && errno == ENOTSUP) xattr_ok_ = false; mem_ino_[path] = ino; }
        return ino;
    }
    uint64_t real_size(const std::string& path)                   // src/main.rs:40-47, plus the header walk when the xattr is absent
    {
        uint8_t b[8];
        if (xattr_ok_ && getxattr(path.c_str(), "user.real_size", b, 8) == 8) return be64(b);
        auto it = mem_size_.find(path);
        if (it != mem_size_.end()) return it->second;
        uint64_t sz = 0;
        const int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
        if (fd >= 0) { if (!content_size_from_headers(fd, &sz)) sz = 0; close(fd); }
        return sz;
    }
    void set_real_size(const std::string& path, int fd, uint64_t size)
    {
        uint8_t b[8]; put_be64(b, size);
        if (!xattr_ok_ || fsetxattr(fd, "user.real_size", b, 8, 0) != 0) mem_size_[path] = size;
    }
    int path_of(uint64_t ino, std::string& out)                   // get_path, src/main.rs:147-172
    {
        if (ino == FUSE_ROOT_ID) { out = data_; return 0; }
        auto it = paths_.find(ino);
        if (it != paths_.end()) { out = it->second; return 0; }
        auto h = by_ino_.find(ino);
        if (h != by_ino_.end()) for (uint64_t fh : h->second) { const Handle& x = handles_[fh]; if (x.has_refs) { out = x.path; return 0; } }
        return ENOENT;
    }
    void fill_attr(struct fuse_attr& a, const struct stat& st, uint64_t ino, uint64_t size)
    {
        memset(&a, 0, sizeof a);
        a.ino = ino; a.size = size; a.blocks = (size + 511) / 512;
        a.atime = (uint64_t)st.st_atime; a.mtime = (uint64_t)st.st_mtime; a.ctime = (uint64_t)st.st_ctime;
        a.mode = S_ISDIR(st.st_mode) ? (S_IFDIR | 0777) : (S_IFREG | 0666);     // access_all, src/main.rs:61-71
        a.nlink = (uint32_t)st.st_nlink; a.uid = st.st_uid; a.gid = st.st_gid; a.blksize = (uint32_t)st.st_blksize;
    }
    int attr_of_path(const std::string& path, uint64_t ino, struct fuse_attr& a)
    {
        struct stat st;
        if (stat(path.c_str(), &st) != 0) return errno;
        fill_attr(a, st, ino, S_ISDIR(st.st_mode) ? (uint64_t)st.st_size : real_size(path));
        auto h = by_ino_.find(ino);                                // an open file: the size of its tmpfile is the truth
        if (h != by_ino_.end() && !h->second.empty()) {
            const Handle& h0 = handles_[*h->second.begin()];
            struct stat ts;
            if (h0.view) { a.size = h0.view_size; a.blocks = (a.size + 511) / 512; }
            else if (fstat(h0.fd, &ts) == 0) { a.size = (uint64_t)ts.st_size; a.blocks = (a.size + 511) / 512; }
        }
        return 0;
    }

    // ---- handle table (src/file.rs)
    uint64_t new_fh() { for (uint64_t i = 0;; i++) if (!handles_.count(i)) return i; }
    uint64_t insert_handle(uint64_t ino, int flags, int fd, const std::string& path)
    {
        const uint64_t fh = new_fh();
        Handle nh; nh.flags = flags; nh.needs_sync = false; nh.fd = fd; nh.has_refs = true; nh.ino = ino; nh.path = path;
        handles_[fh] = nh;
        by_ino_[ino].insert(fh);
        return fh;
    }
    static void quiesce(const Handle& h) { while (h.busy->load(std::memory_order_acquire) > 0) sched_yield(); }
    void drop_handle_backing(Handle& h) { quiesce(h); if (h.view) { fzfs_unview(h.pin); h.view = nullptr; h.pin = nullptr; } else if (h.fd >= 0) close(h.fd); h.fd = -1; }
    // every in-place handle of `ino` gets the ordinary backing after all: ONE tmpfile holding the plain bytes, dup'ed per handle
    int materialize(uint64_t ino)
    {
        auto m = by_ino_.find(ino);
        if (m == by_ino_.end()) return 0;
        int tmp = -1;
        for (uint64_t fh : m->second) {
            Handle& h = handles_[fh];
            if (!h.view) continue;
            int fd;
            if (tmp < 0) {
                char tmpl[] = "/tmp/fzfs-XXXXXX";
                tmp = mkstemp(tmpl);
                if (tmp < 0) return errno;
                unlink(tmpl);
                for (uint64_t o = 0; o < h.view_size;) { const ssize_t w = pwrite(tmp, h.view + o, h.view_size - o, (off_t)o); if (w < 0) { if (errno == EINTR) continue; const int e = errno; close(tmp); return e; } o += (uint64_t)w; }
                fd = tmp;
            } else fd = dup(tmp);
            if (fd < 0) return errno;
            quiesce(h);
            fzfs_unview(h.pin); h.view = nullptr; h.pin = nullptr; h.fd = fd;
        }
        return 0;
    }
    void forget_ino(uint64_t ino)                                 // OpenedFiles::unlink: later syncs of these handles are no-ops
    {
        auto it = by_ino_.find(ino);
        if (it == by_ino_.end()) return;
        for (uint64_t fh : it->second) handles_[fh].has_refs = false;
        by_ino_.erase(it);
    }

    // ---- the two codec call sites
    int store_to_source_file(int plain_fd, const std::string& dir, const std::string& name, uint64_t* ino_out);   // src/main.rs:755-832
    int sync_to_fs(uint64_t fh, bool close_it, bool force);                                                        // src/main.rs:174-213
    int do_open(uint64_t ino, int flags, uint64_t* fh_out);                                                        // src/main.rs:451-493
    void readahead_dir(const std::string& dir, const std::string& path);     // path empty: the directory was listed, its first window

    // ---- protocol
    void reply(uint64_t unique, int err, const void* p = nullptr, size_t n = 0) { (void)send(unique, err, p, n); }
    bool send(uint64_t unique, int err, const void* p = nullptr, size_t n = 0)      // false: the kernel no longer wants the answer (interrupted)
    {
        struct fuse_out_header oh; oh.len = (uint32_t)(sizeof oh + (err ? 0 : n)); oh.error = -err; oh.unique = unique;
        struct iovec iov[2] = { { &oh, sizeof oh }, { const_cast<void*>(p), err ? 0 : n } };
        if (writev(dev, iov, err || n == 0 ? 1 : 2) >= 0) return true;
        if (errno != ENOENT) logf("fzfs: reply failed: %s", strerror(errno));
        return false;
    }
    void entry_out(struct fuse_entry_out& e, const struct fuse_attr& a) { memset(&e, 0, sizeof e); e.nodeid = a.ino; e.entry_valid = kTtlSec; e.attr_valid = kTtlSec; e.attr = a; }
    int lookup(uint64_t parent, const char* name, struct fuse_attr& a);
    void dispatch(const struct fuse_in_header* in, const uint8_t* arg, size_t arglen);
};

int Fs::lookup(uint64_t parent, const char* name, struct fuse_attr& a)      // lookup_wrapper, src/main.rs:215-305 (without --convert)
{
    std::string dir;
    if (int e = path_of(parent, dir)) return e;
    if (parent == FUSE_ROOT_ID && strcmp(name, ".fuse-zstd-inode_cache") == 0) return ENOENT;
    struct stat st;
    std::string p = dir + "/" + name;
    if (stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) {
        const uint64_t ino = ino_of(p);
        paths_[ino] = p;
        fill_attr(a, st, ino, (uint64_t)st.st_size);
        return 0;
    }
    p += ".zst";
    if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return ENOENT;
    const uint64_t ino = ino_of(p);
    paths_[ino] = p;
    return attr_of_path(p, ino, a);
}

int Fs::store_to_source_file(int plain_fd, const std::string& dir, const std::string& name, uint64_t* ino_out)
{
    std::string tmpl = dir + "/.tmpXXXXXX";
    std::vector<char> tbuf(tmpl.begin(), tmpl.end()); tbuf.push_back(0);
    const int out = mkstemp(tbuf.data());                          // NamedTempFile::new_in(dir)
    if (out < 0) return errno;
    const std::string tmp_path(tbuf.data()), path = dir + "/" + name;
    int err = 0;
    struct stat st;
    if (fsync(plain_fd) != 0 || fstat(plain_fd, &st) != 0) err = errno;
    const uint64_t real = (uint64_t)st.st_size;
    uint64_t ino = 0;
    if (!err) {
        const int src = dup(plain_fd);                             // try_clone: shares the offset, so rewind first (src/main.rs:776-779)
        if (src < 0) err = errno;
        else {
            if (lseek(src, 0, SEEK_SET) < 0) err = errno;
            if (!err) { uint8_t b[8]; ino = getxattr(path.c_str(), "user.ino", b, 8) == 8 ? be64(b) : 0; }
            if (!err && fzfs_encode(src, out, level_, real, ino, nullptr) != 0) err = EIO;      // src/errors.rs:4-10
            close(src);
        }
    }
    if (!err) {                                                    // user.ino travels with the name (the rename changes the backing inode)
        if (!ino) { auto it = mem_ino_.find(path); ino = it != mem_ino_.end() ? it->second : next_ino(); }
        uint8_t b[8]; put_be64(b, ino);
        if (!xattr_ok_ || fsetxattr(out, "user.ino", b, 8, 0) != 0) { if (errno == ENOTSUP) xattr_ok_ = false; mem_ino_[path] = ino; }
        if (fsync(out) != 0) err = errno;
    }
    if (!err && rename(tmp_path.c_str(), path.c_str()) != 0) err = errno;      // persist: atomic
    if (!err) { set_real_size(path, out, real); if (fsync(out) != 0) err = errno; fzfs_invalidate(ino); }
    if (err) unlink(tmp_path.c_str());
    close(out);
    if (ino_out) *ino_out = ino;
    return err;
}

int Fs::sync_to_fs(uint64_t fh, bool close_it, bool force)
{
    auto it = handles_.find(fh);
    if (it == handles_.end()) return close_it ? EBADF : ENOENT;
    Handle h = it->second;
    if (close_it) {
        handles_.erase(it);
        if (h.has_refs) { auto m = by_ino_.find(h.ino); if (m != by_ino_.end()) { m->second.erase(fh); if (m->second.empty()) by_ino_.erase(m); } }
    }
    int err = 0;
    if ((h.needs_sync || force) && h.has_refs && h.view) {          // fsync of an in-place handle: it needs a file to encode from
        if (!close_it) { err = materialize(h.ino); if (err) return err; h = handles_[fh]; }
        else {                                                      // (closing: this copy of the handle is the only one left)
            char tmpl[] = "/tmp/fzfs-XXXXXX"; const int tmp = mkstemp(tmpl);
            if (tmp < 0) return errno;
            unlink(tmpl);
            for (uint64_t o = 0; o < h.view_size;) { const ssize_t w = pwrite(tmp, h.view + o, h.view_size - o, (off_t)o); if (w < 0) { if (errno == EINTR) continue; break; } o += (uint64_t)w; }
            quiesce(h);
            fzfs_unview(h.pin); h.view = nullptr; h.pin = nullptr; h.fd = tmp;
        }
    }
    if ((h.needs_sync || force) && h.has_refs) {
        const size_t k = h.path.rfind('/');
        err = store_to_source_file(h.fd, h.path.substr(0, k), h.path.substr(k + 1), nullptr);
        if (!err && !close_it) handles_[fh].needs_sync = false;
    }
    if (close_it) drop_handle_backing(h);
    return err;
}

// "t3.0.9" < "t3.0.10": digit runs compare by value, so that the window follows the order in which a job walks its files
static bool natural_less(const std::string& a, const std::string& b)
{
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (isdigit((unsigned char)a[i]) && isdigit((unsigned char)b[j])) {
            size_t i2 = i, j2 = j;
            while (i2 < a.size() && a[i2] == '0') i2++;
            while (j2 < b.size() && b[j2] == '0') j2++;
            size_t i3 = i2, j3 = j2;
            while (i3 < a.size() && isdigit((unsigned char)a[i3])) i3++;
            while (j3 < b.size() && isdigit((unsigned char)b[j3])) j3++;
            if (i3 - i2 != j3 - j2) return i3 - i2 < j3 - j2;
            const int c = a.compare(i2, i3 - i2, b, j2, j3 - j2);
            if (c) return c < 0;
            i = i3; j = j3;
        } else {
            if (a[i] != b[j]) return a[i] < b[j];
            i++; j++;
        }
    }
    return a.size() - i < b.size() - j;
}

// Batch formation: the file being opened and the next kRaWindow - 1 entries of its directory go to the codec as
// one batch, and the window is topped up whenever fewer than half of the entries ahead have been asked for.  A window, not the
// whole directory: parallel-files.fio has 16 jobs walking 1250 files each -- 20 GB of plain bytes, more than any cache --
// and a directory decoded at once was evicted by the other jobs' directories before its reader arrived (measured: 365 MB/s
// against 610 for the CPU path; see profiles/r02_notes.md).
constexpr size_t kRaWindow = 64;
constexpr int kWaitForBatch = -1;               // do_open: try again once fzfs_wait(ino) has returned
void Fs::readahead_dir(const std::string& dir, const std::string& path)
{
    if (!readahead_) return;
    struct stat ds;
    if (stat(dir.c_str(), &ds) != 0) return;
    const int64_t mt = (int64_t)ds.st_mtim.tv_sec * 1000000000ll + ds.st_mtim.tv_nsec;
    DirRa& ra = ra_[dir];
    if (ra.mtime_ns != mt) {                                         // first visit, or entries came / went: list again
        ra = DirRa(); ra.mtime_ns = mt;
        if (DIR* d = opendir(dir.c_str())) {
            while (struct dirent* e = readdir(d)) { const std::string n = e->d_name; if (ends_with(n, ".zst")) ra.names.push_back(n); }
            closedir(d);
        }
        std::sort(ra.names.begin(), ra.names.end(), natural_less);
        ra.asked.assign(ra.names.size(), 0);
        for (size_t i = 0; i < ra.names.size(); i++) ra.index[ra.names[i]] = i;
    }
    size_t i0 = 0;
    if (!path.empty()) {
        auto it = ra.index.find(path.substr(path.rfind('/') + 1));
        if (it == ra.index.end()) return;
        i0 = it->second;
    }
    const size_t i1 = std::min(ra.names.size(), i0 + kRaWindow);
    if (i0 >= i1) return;
    // entries well behind the reader have been served (and their slabs may be reused): a later pass over the directory asks again
    for (size_t i = i0 > 2 * kRaWindow ? i0 - 2 * kRaWindow : 0, e = i0 > kRaWindow ? i0 - kRaWindow : 0; i < e; i++) ra.asked[i] = 0;
    size_t ahead = 0;
    for (size_t i = i0; i < i1; i++) ahead += ra.asked[i];
    if (ahead * 2 > i1 - i0) return;                                  // more than half of the window is on its way or here
    std::vector<std::string> paths; std::vector<uint64_t> keys;
    for (size_t i = i0; i < i1; i++) {
        if (ra.asked[i]) continue;
        ra.asked[i] = 1;
        const std::string p = dir + "/" + ra.names[i];
        struct stat st;
        if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
        const uint64_t ino = ino_of(p);
        paths_[ino] = p; paths.push_back(p); keys.push_back(ino);
    }
    std::vector<const char*> c(paths.size());
    for (size_t i = 0; i < paths.size(); i++) c[i] = paths[i].c_str();
    if (!c.empty()) fzfs_prefetch(c.data(), keys.data(), c.size());
}

int Fs::do_open(uint64_t ino, int flags, uint64_t* fh_out)
{
    auto m = by_ino_.find(ino);
    if (m != by_ino_.end() && !m->second.empty()) {                 // OpenedFiles::duplicate: dup the tmpfile, no decode
        if ((flags & O_ACCMODE) != O_RDONLY) { if (int e = materialize(ino)) return e; }    // a writer joins: every handle moves to one tmpfile
        const Handle& h0 = handles_[*m->second.begin()];
        if (h0.view) {                                               // another reader of bytes that are served in place: share them
            const int src = open(h0.path.c_str(), O_RDONLY | O_CLOEXEC);
            const void* data = nullptr; uint64_t vs = 0; void* pin = nullptr;
            const bool ok = src >= 0 && fzfs_view(src, ino, &data, &vs, &pin) == 0 && data == (const void*)h0.view;
            if (src >= 0) close(src);
            if (!ok) { if (pin) fzfs_unview(pin); if (int e = materialize(ino)) return e; }
            else { const std::string hp = h0.path; *fh_out = insert_handle(ino, flags, -1, hp); Handle& nh = handles_[*fh_out]; nh.view = (const uint8_t*)data; nh.view_size = vs; nh.pin = pin; return 0; }
        }
        const Handle& h1 = handles_[*m->second.begin()];
        const int fd = dup(h1.fd);
        if (fd < 0) return errno;
        *fh_out = insert_handle(ino, flags, fd, h1.path);
        return 0;
    }
    std::string path;
    if (int e = path_of(ino, path)) return e;
    readahead_dir(dir_of(path), path);
    if (readahead_ && fzfs_pending(ino)) return kWaitForBatch;      // its batch is on the way (maybe asked for just now): the caller waits WITHOUT the lock
    const int src = open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (src < 0) return errno;
    if (readahead_ && (flags & O_ACCMODE) == O_RDONLY) {            // a reader of a file the readahead has decoded: its bytes in place, no tmpfile
        const void* data = nullptr; uint64_t vs = 0; void* pin = nullptr;
        if (fzfs_view(src, ino, &data, &vs, &pin) == 0) {
            uint8_t b[8]; put_be64(b, vs);
            uint8_t cur[8];
            int verr = 0;
            const bool same = xattr_ok_ && fgetxattr(src, "user.real_size", cur, 8) == 8 && memcmp(cur, b, 8) == 0;
            if (!same) { set_real_size(path, src, vs); if (xattr_ok_ && fsync(src) != 0) verr = errno; }   // as below (src/main.rs:473-484)
            close(src);
            if (verr) { fzfs_unview(pin); return verr; }
            *fh_out = insert_handle(ino, flags, -1, path);
            Handle& nh = handles_[*fh_out]; nh.view = (const uint8_t*)data; nh.view_size = vs; nh.pin = pin;
            return 0;
        }
    }
    char tmpl[] = "/tmp/fzfs-XXXXXX";
    const int tmp = mkstemp(tmpl);                                  // tempfile::tempfile(): unlinked at once
    if (tmp < 0) { const int e = errno; close(src); return e; }
    unlink(tmpl);
    uint64_t size = 0;
    if (fzfs_decode(src, tmp, ino, &size) != 0) { close(src); close(tmp); return EFAULT; }     // src/main.rs:467
    lseek(tmp, 0, SEEK_SET);
    struct stat st;
    int err = fstat(tmp, &st) != 0 ? errno : 0;
    if (!err) {
        uint8_t b[8]; put_be64(b, (uint64_t)st.st_size);
        uint8_t cur[8];
        const bool same = xattr_ok_ && fgetxattr(src, "user.real_size", cur, 8) == 8 && memcmp(cur, b, 8) == 0;
        if (!same) {                                                // the reference rewrites and fsyncs on EVERY first open (src/main.rs:473-484);
            set_real_size(path, src, (uint64_t)st.st_size);         // an unchanged value needs neither
            if (xattr_ok_ && fsync(src) != 0) err = errno;
        }
    }
    close(src);
    if (err) { close(tmp); return err; }
    *fh_out = insert_handle(ino, flags, tmp, path);
    return 0;
}

void Fs::dispatch(const struct fuse_in_header* in, const uint8_t* arg, size_t arglen)
{
    const uint64_t u = in->unique, node = in->nodeid;
    const uint64_t w0 = g_stats ? now_ns() : 0;
    if (in->opcode == FUSE_READ) {
        // the lock only for the lookup; the copy (128 KiB into the kernel: most of a READ's time) runs beside everything else
        const struct fuse_read_in* r = (const struct fuse_read_in*)arg;
        std::shared_ptr<std::atomic<int>> busy; const uint8_t* view = nullptr; uint64_t vs = 0; int fd = -1;
        {
            std::shared_lock<std::shared_mutex> lk(mu_);
            if (g_stats) t_wait_ns = now_ns() - w0;
            auto h = handles_.find(r->fh);
            if (h == handles_.end()) { lk.unlock(); return reply(u, ENOENT); }
            busy = h->second.busy; busy->fetch_add(1, std::memory_order_acquire);
            view = h->second.view; vs = h->second.view_size; fd = h->second.fd;
        }
        struct Idle { std::atomic<int>& b; ~Idle() { b.fetch_sub(1, std::memory_order_release); } } idle{ *busy };
        if (view) {                                                  // served in place: straight from the cache's pinned memory
            const uint64_t off = r->offset < vs ? r->offset : vs;
            return reply(u, 0, view + off, (size_t)std::min<uint64_t>(r->size, vs - off));
        }
        static thread_local std::vector<uint8_t> rbuf;             // one reply buffer per serving thread: the reference's
        if (rbuf.size() < r->size) rbuf.resize(r->size);           // vec![0; size] per request (src/main.rs:503) zeroes 128 KiB each time
        const ssize_t n = pread(fd, rbuf.data(), r->size, (off_t)r->offset);
        if (n < 0) return reply(u, errno);
        return reply(u, 0, rbuf.data(), (size_t)n);
    }
    if (in->opcode == FUSE_WRITE) {
        // like READ: the lock for the lookup, the pwrite beside everything else.  Appends stay exclusive (end-of-file + write must not
        // interleave between handles of one file, src/main.rs:576-588), and so does a handle without a file of its own.
        const struct fuse_write_in* w = (const struct fuse_write_in*)arg;
        std::shared_ptr<std::atomic<int>> busy; int fd = -1;
        {
            std::shared_lock<std::shared_mutex> lk(mu_);
            if (g_stats) t_wait_ns = now_ns() - w0;
            auto h = handles_.find(w->fh);
            if (h == handles_.end()) { lk.unlock(); return reply(u, EBADF); }
            if (!(h->second.flags & O_APPEND) && h->second.fd >= 0) {
                busy = h->second.busy; busy->fetch_add(1, std::memory_order_acquire);
                fd = h->second.fd;
                __atomic_store_n(&h->second.needs_sync, true, __ATOMIC_RELAXED);      // several writers may say so at once
            }
        }
        if (busy) {
            struct Idle { std::atomic<int>& b; ~Idle() { b.fetch_sub(1, std::memory_order_release); } } idle{ *busy };
            const ssize_t n = pwrite(fd, arg + sizeof *w, w->size, (off_t)w->offset);
            if (n < 0) return reply(u, errno);
            struct fuse_write_out o; memset(&o, 0, sizeof o); o.size = (uint32_t)n;
            return reply(u, 0, &o, sizeof o);
        }
    }
    std::unique_lock<std::shared_mutex> lk(mu_);
    if (g_stats) t_wait_ns = now_ns() - w0;
    switch (in->opcode) {
    case FUSE_LOOKUP: {
        struct fuse_attr a; struct fuse_entry_out e;
        const int err = lookup(node, (const char*)arg, a);
        if (err) return reply(u, err);
        entry_out(e, a); return reply(u, 0, &e, sizeof e);
    }
    case FUSE_FORGET: case FUSE_BATCH_FORGET: case FUSE_INTERRUPT: return;      // no reply
    case FUSE_GETATTR: {
        std::string p; struct fuse_attr_out o; memset(&o, 0, sizeof o);
        int err = path_of(node, p);
        if (!err) err = attr_of_path(p, node, o.attr);
        if (err) return reply(u, err);
        o.attr_valid = kTtlSec; return reply(u, 0, &o, sizeof o);
    }
    case FUSE_SETATTR: {                                            // only truncation (src/main.rs:408-449)
        const struct fuse_setattr_in* s = (const struct fuse_setattr_in*)arg;
        int err = 0;
        if (s->valid & FATTR_SIZE) {
            err = materialize(node);
            if (!err && (s->valid & FATTR_FH)) { auto h = handles_.find(s->fh); if (h != handles_.end() && ftruncate(h->second.fd, (off_t)s->size) != 0) err = errno; }
            auto m = by_ino_.find(node);
            if (!err && m != by_ino_.end()) for (uint64_t fh : m->second) if (ftruncate(handles_[fh].fd, (off_t)s->size) != 0) err = errno;
        }
        std::string p; struct fuse_attr_out o; memset(&o, 0, sizeof o);
        if (!err) err = path_of(node, p);
        if (!err) err = attr_of_path(p, node, o.attr);
        if (err) return reply(u, err);
        o.attr_valid = kTtlSec; return reply(u, 0, &o, sizeof o);
    }
    case FUSE_OPEN: {
        const struct fuse_open_in* oi = (const struct fuse_open_in*)arg;
        uint64_t fh = 0;
        const int flags = (int)oi->flags;
        int err = do_open(node, flags, &fh);
        if (err == kWaitForBatch) {
            // A reader that has caught up with the readahead finds its file in a batch in flight.  The request is parked on a thread
            // of its own, which answers it when the batch has landed (a reply may come from any thread, in any order): the serving
            // threads go on with everybody else's requests meanwhile.
            lk.unlock();
            parked_.fetch_add(1);
            try {
                std::thread([this, u, node, flags] {
                    uint64_t fh2 = 0; int e;
                    do { fzfs_wait(node); std::unique_lock<std::shared_mutex> lk2(mu_); e = do_open(node, flags, &fh2); } while (e == kWaitForBatch);
                    struct fuse_open_out o2; memset(&o2, 0, sizeof o2); o2.fh = fh2;
                    if (e) reply(u, e);
                    else if (!send(u, 0, &o2, sizeof o2)) { std::unique_lock<std::shared_mutex> lk3(mu_); sync_to_fs(fh2, true, false); }   // the opener is gone: no RELEASE will come
                    parked_.fetch_sub(1);
                }).detach();
                return;
            } catch (...) { parked_.fetch_sub(1); }                 // no thread to be had: wait here after all
            lk.lock();
            while ((err = do_open(node, flags, &fh)) == kWaitForBatch) { lk.unlock(); fzfs_wait(node); lk.lock(); }
        }
        if (err) return reply(u, err);
        struct fuse_open_out o; memset(&o, 0, sizeof o); o.fh = fh;
        if (!send(u, 0, &o, sizeof o)) sync_to_fs(fh, true, false);   // the opener is gone: no RELEASE will come
        return;
    }
    case FUSE_WRITE: {
        const struct fuse_write_in* w = (const struct fuse_write_in*)arg;
        auto h = handles_.find(w->fh);
        if (h == handles_.end()) return reply(u, EBADF);
        h->second.needs_sync = true;
        off_t off = (off_t)w->offset;
        if (h->second.flags & O_APPEND) off = lseek(h->second.fd, 0, SEEK_END);     // src/main.rs:576-588
        const ssize_t n = pwrite(h->second.fd, arg + sizeof *w, w->size, off);
        if (n < 0) return reply(u, errno);
        struct fuse_write_out o; memset(&o, 0, sizeof o); o.size = (uint32_t)n;
        return reply(u, 0, &o, sizeof o);
    }
    case FUSE_FLUSH: return reply(u, sync_to_fs(((const struct fuse_flush_in*)arg)->fh, false, false));
    case FUSE_FSYNC: return reply(u, sync_to_fs(((const struct fuse_fsync_in*)arg)->fh, false, true));
    case FUSE_RELEASE: {
        const int err = sync_to_fs(((const struct fuse_release_in*)arg)->fh, true, false);
        return reply(u, err == EBADF ? 0 : err);                    // src/main.rs:1010-1013
    }
    case FUSE_OPENDIR: {
        std::string p;
        if (int e = path_of(node, p)) return reply(u, e);
        readahead_dir(p, std::string());
        struct fuse_open_out o; memset(&o, 0, sizeof o);
        return reply(u, 0, &o, sizeof o);
    }
    case FUSE_RELEASEDIR: case FUSE_FSYNCDIR: return reply(u, 0);
    case FUSE_READDIR: {                                            // readdir_wrapper, src/main.rs:307-387
        const struct fuse_read_in* r = (const struct fuse_read_in*)arg;
        std::string dir;
        if (int e = path_of(node, dir)) return reply(u, e);
        DIR* d = opendir(dir.c_str());
        if (!d) return reply(u, errno);
        std::vector<uint8_t> out; uint64_t idx = 0;
        while (struct dirent* e = readdir(d)) {
            std::string n = e->d_name;
            if (n == "." || n == "..") continue;
            if (node == FUSE_ROOT_ID && n == ".fuse-zstd-inode_cache") continue;
            const std::string p = dir + "/" + n;
            struct stat st;
            if (stat(p.c_str(), &st) != 0) continue;
            uint32_t type;
            if (S_ISDIR(st.st_mode)) type = DT_DIR;
            else if (S_ISREG(st.st_mode) && ends_with(n, ".zst")) { type = DT_REG; n.resize(n.size() - 4); }
            else continue;                                          // other regular files are hidden, other types skipped
            idx++;
            if (idx <= r->offset) continue;
            const uint64_t ino = ino_of(p);
            paths_[ino] = p;
            const size_t ent = FUSE_NAME_OFFSET + n.size(), padded = FUSE_DIRENT_ALIGN(ent);
            if (out.size() + padded > r->size) break;
            const size_t at = out.size(); out.resize(at + padded, 0);
            struct fuse_dirent* de = (struct fuse_dirent*)(out.data() + at);
            de->ino = ino; de->off = idx; de->namelen = (uint32_t)n.size(); de->type = type;
            memcpy(de->name, n.data(), n.size());
        }
        closedir(d);
        return reply(u, 0, out.data(), out.size());
    }
    case FUSE_CREATE: {                                             // create_wrapper, src/main.rs:515-555: an EMPTY file goes through the encoder
        const struct fuse_create_in* c = (const struct fuse_create_in*)arg;
        const std::string name = std::string((const char*)arg + sizeof *c) + ".zst";
        std::string dir;
        if (int e = path_of(node, dir)) return reply(u, e);
        char tmpl[] = "/tmp/fzfs-XXXXXX";
        const int tmp = mkstemp(tmpl);
        if (tmp < 0) return reply(u, errno);
        unlink(tmpl);
        uint64_t ino = 0;
        if (int e = store_to_source_file(tmp, dir, name, &ino)) { close(tmp); return reply(u, e); }
        const std::string p = dir + "/" + name;
        paths_[ino] = p;
        struct { struct fuse_entry_out e; struct fuse_open_out o; } out; memset(&out, 0, sizeof out);
        struct fuse_attr a; struct stat st;
        if (stat(p.c_str(), &st) != 0) { const int e = errno; close(tmp); return reply(u, e); }
        fill_attr(a, st, ino, 0);
        entry_out(out.e, a);
        out.o.fh = insert_handle(ino, (int)c->flags, tmp, p);
        return reply(u, 0, &out, sizeof out);
    }
    case FUSE_MKDIR: {
        const struct fuse_mkdir_in* mk = (const struct fuse_mkdir_in*)arg;
        std::string dir;
        if (int e = path_of(node, dir)) return reply(u, e);
        const std::string p = dir + "/" + (const char*)(arg + sizeof *mk);
        if (mkdir(p.c_str(), 0777) != 0) return reply(u, errno);
        struct stat st;
        if (stat(p.c_str(), &st) != 0) return reply(u, errno);
        const uint64_t ino = ino_of(p);
        paths_[ino] = p;
        struct fuse_attr a; struct fuse_entry_out e;
        fill_attr(a, st, ino, (uint64_t)st.st_size);
        entry_out(e, a); return reply(u, 0, &e, sizeof e);
    }
    case FUSE_UNLINK: case FUSE_RMDIR: {
        std::string dir;
        if (int e = path_of(node, dir)) return reply(u, e);
        const bool file = in->opcode == FUSE_UNLINK;
        const std::string p = dir + "/" + (const char*)arg + (file ? ".zst" : "");
        if (!file && p == data_ + "/.fuse-zstd-inode_cache") return reply(u, ENOENT);
        struct stat st;
        const bool had = stat(p.c_str(), &st) == 0;
        const uint64_t ino = had ? ino_of(p) : 0;
        if ((file ? unlink(p.c_str()) : rmdir(p.c_str())) != 0) return reply(u, errno);
        // forget the entry only once it is gone (the reference drops its inode mapping first, src/main.rs:601-648, so a failed
        // rmdir of a non-empty directory leaves that directory unreachable until the kernel looks it up again)
        if (had) { paths_.erase(ino); forget_ino(ino); mem_ino_.erase(p); mem_size_.erase(p); fzfs_invalidate(ino); }
        return reply(u, 0);
    }
    case FUSE_RENAME: case FUSE_RENAME2: {                          // rename_wrapper, src/main.rs:650-707
        uint64_t newdir; const char* names;
        if (in->opcode == FUSE_RENAME) { newdir = ((const struct fuse_rename_in*)arg)->newdir; names = (const char*)arg + sizeof(struct fuse_rename_in); }
        else { newdir = ((const struct fuse_rename2_in*)arg)->newdir; names = (const char*)arg + sizeof(struct fuse_rename2_in); }
        const char* oldname = names; const char* newname = names + strlen(names) + 1;
        struct fuse_attr a;
        if (int e = lookup(node, oldname, a)) return reply(u, e);
        const bool file = (a.mode & S_IFMT) == S_IFREG;
        std::string from, to;
        if (int e = path_of(node, from)) return reply(u, e);
        if (int e = path_of(newdir, to)) return reply(u, e);
        from += std::string("/") + oldname + (file ? ".zst" : ""); to += std::string("/") + newname + (file ? ".zst" : "");
        struct stat st;
        if (stat(to.c_str(), &st) == 0) { const uint64_t old = ino_of(to); paths_.erase(old); forget_ino(old); fzfs_invalidate(old); }
        if (rename(from.c_str(), to.c_str()) != 0) return reply(u, errno);
        paths_[a.ino] = to;
        // every path that starts with the old name moves with it: the entry itself, and for a directory everything below it
        // (the kernel keeps using the inode numbers it has cached; the reference leaves those, and open handles, pointing at
        // the old location -- its TODO at src/main.rs:703)
        auto moved = [&](const std::string& p, std::string& out) { if (p == from) { out = to; return true; } if (p.size() > from.size() && p.compare(0, from.size(), from) == 0 && p[from.size()] == '/') { out = to + p.substr(from.size()); return true; } return false; };
        std::string np;
        for (auto& kv : paths_) if (moved(kv.second, np)) kv.second = np;
        for (auto& kv : handles_) if (kv.second.has_refs && moved(kv.second.path, np)) kv.second.path = np;
        auto rekey = [&](std::unordered_map<std::string, uint64_t>& mp) { std::vector<std::pair<std::string, uint64_t>> add; for (auto it = mp.begin(); it != mp.end();) { if (moved(it->first, np)) { add.emplace_back(np, it->second); it = mp.erase(it); } else ++it; } for (auto& kv : add) mp[kv.first] = kv.second; };
        rekey(mem_ino_); rekey(mem_size_);
        return reply(u, 0);
    }
    case FUSE_STATFS: {
        struct statvfs sv; struct fuse_statfs_out o; memset(&o, 0, sizeof o);
        if (statvfs(data_.c_str(), &sv) != 0) return reply(u, errno);
        o.st.blocks = sv.f_blocks; o.st.bfree = sv.f_bfree; o.st.bavail = sv.f_bavail; o.st.files = sv.f_files; o.st.ffree = sv.f_ffree;
        o.st.bsize = (uint32_t)sv.f_bsize; o.st.namelen = 251; o.st.frsize = (uint32_t)sv.f_frsize;
        return reply(u, 0, &o, sizeof o);
    }
    case FUSE_ACCESS: return reply(u, 0);
    default: (void)arglen; return reply(u, ENOSYS);
    }
}

volatile sig_atomic_t g_stop = 0;
std::string g_mountpoint;

void Fs::serve()
{
    std::vector<uint8_t> buf((1u << 20) + 65536);
    for (;;) {
        const ssize_t n = read(dev, buf.data(), buf.size());
        if (n < 0) { if (errno == EINTR && !g_stop) continue; if (errno == ENOENT) continue; break; }      // ENODEV: unmounted
        if ((size_t)n < sizeof(struct fuse_in_header)) break;
        const struct fuse_in_header* in = (const struct fuse_in_header*)buf.data();
        const uint8_t* arg = buf.data() + sizeof *in; const size_t arglen = (size_t)n - sizeof *in;
        logf("fzfs: op %u node %llu", in->opcode, (unsigned long long)in->nodeid);
        if (in->opcode == FUSE_INIT) {
            const struct fuse_init_in* ii = (const struct fuse_init_in*)arg;
            struct fuse_init_out o; memset(&o, 0, sizeof o);
            o.major = FUSE_KERNEL_VERSION; o.minor = ii->minor < FUSE_KERNEL_MINOR_VERSION ? ii->minor : FUSE_KERNEL_MINOR_VERSION;
            o.max_readahead = ii->max_readahead;
            o.flags = (FUSE_BIG_WRITES | FUSE_MAX_PAGES) & ii->flags;     // no atomic O_TRUNC: truncation arrives as SETATTR, as with fuser
            o.max_background = 16; o.congestion_threshold = 12; o.max_write = 1u << 20; o.max_pages = 256; o.time_gran = 1;
            if (n_threads > 1) { o.max_background = (uint16_t)std::min(1024, 16 * n_threads); o.congestion_threshold = (uint16_t)(o.max_background * 3 / 4); }   // the kernel's own readahead is background traffic
            reply(in->unique, 0, &o, sizeof o);
            return;                                                   // the caller starts the serving threads now
        }
        if (in->opcode == FUSE_DESTROY) { reply(in->unique, 0); g_stop = 1; break; }
        if (!g_stats) { dispatch(in, arg, arglen); continue; }
        const uint32_t op = in->opcode < 64 ? in->opcode : 0;
        const uint64_t t0 = now_ns();
        t_wait_ns = 0;
        dispatch(in, arg, arglen);
        const uint64_t dt = now_ns() - t0;
        g_op[op].n++; g_op[op].ns += dt; g_op[op].wait_ns += t_wait_ns;
        uint64_t mx = g_op[op].max_ns.load(); while (dt > mx && !g_op[op].max_ns.compare_exchange_weak(mx, dt)) { }
    }
}

// The kernel hands each request to whichever thread is waiting in read() on the device (what libfuse's multi-threaded loop does).
// The first call to serve() answers FUSE_INIT on this thread and returns; then `threads` of them serve until the unmount.
void Fs::loop(int threads)
{
    n_threads = threads;
    serve();
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back([this] { serve(); });
    serve();
    if (!pool.empty()) umount2(g_mountpoint.c_str(), MNT_DETACH);    // the others are waiting in read(): the unmount sends them home
    for (auto& t : pool) t.join();
    while (parked_.load() > 0) usleep(1000);                         // batches always land (or fail): nobody waits for long
    if (g_stats) {
        static const char* names[64] = {};
        names[FUSE_LOOKUP] = "lookup"; names[FUSE_GETATTR] = "getattr"; names[FUSE_SETATTR] = "setattr"; names[FUSE_OPEN] = "open"; names[FUSE_READ] = "read";
        names[FUSE_WRITE] = "write"; names[FUSE_FLUSH] = "flush"; names[FUSE_RELEASE] = "release"; names[FUSE_FSYNC] = "fsync"; names[FUSE_OPENDIR] = "opendir";
        names[FUSE_READDIR] = "readdir"; names[FUSE_CREATE] = "create"; names[FUSE_UNLINK] = "unlink"; names[FUSE_RENAME] = "rename"; names[FUSE_MKDIR] = "mkdir";
        for (int op = 0; op < 64; op++) if (g_op[op].n) fprintf(stderr, "fzfs: stats %-8s n %8llu  in the host %9.1f ms (%.1f us each, max %.1f us), of which waiting %9.1f ms\n", names[op] ? names[op] : "other",
            (unsigned long long)g_op[op].n.load(), g_op[op].ns / 1e6, g_op[op].ns / 1e3 / g_op[op].n, g_op[op].max_ns / 1e3, g_op[op].wait_ns / 1e6);
    }
    for (auto& kv : handles_) drop_handle_backing(kv.second);
}

void on_signal(int) { g_stop = 1; umount2(g_mountpoint.c_str(), MNT_DETACH); }

}  // namespace

int main(int argc, char** argv)
{
    std::string data_dir, mountpoint; int level = 0; bool readahead = true; size_t cache_mb = 1024; int threads = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&](const char* name) -> const char* { if (i + 1 >= argc) { fprintf(stderr, "fzfs: %s needs a value\n", name); exit(2); } return argv[++i]; };
        if (a == "--data-dir") data_dir = val("--data-dir");
        else if (a == "--mount-point") mountpoint = val("--mount-point");
        else if (a == "--compression-level") level = atoi(val("--compression-level"));
        else if (a == "--no-readahead") readahead = false;
        else if (a == "--cache-mb") cache_mb = (size_t)atoll(val("--cache-mb"));
        else if (a == "--threads") threads = atoi(val("--threads"));
        else if (a == "--verbose") g_verbose = true;
        else if (a == "--stats") g_stats = true;
        else { fprintf(stderr, "usage: %s --data-dir DIR --mount-point DIR [--compression-level 0..19] [--no-readahead] [--cache-mb N] [--threads N] [--stats] [--verbose]\n", argv[0]); return 2; }
    }
    if (data_dir.empty() || mountpoint.empty()) { fprintf(stderr, "fzfs: --data-dir and --mount-point are required\n"); return 2; }
    if (level < 0 || level > 19) level = 0;                         // src/main.rs:1283-1296
    if (int rc = fzfs_codec_init(readahead ? cache_mb << 20 : 0)) { fprintf(stderr, "fzfs: codec unavailable (%s): %s\n", fzfs_codec_name(), strerror(rc < 0 ? -rc : rc)); return 1; }
    if (threads <= 0) threads = fzfs_codec_threads();                  // the codec's default: 1 for the reference's (fuser's loop), more for the GPU's
    if (threads > 64) threads = 64;
    Fs fs(data_dir, level, readahead);
    fs.dev = open("/dev/fuse", O_RDWR | O_CLOEXEC);
    if (fs.dev < 0) { perror("fzfs: /dev/fuse"); return 1; }
    char opts[160];
    snprintf(opts, sizeof opts, "fd=%d,rootmode=40000,user_id=%u,group_id=%u,allow_other,default_permissions", fs.dev, getuid(), getgid());
    if (mount("fuse-zstd", mountpoint.c_str(), "fuse.fzfs", MS_NOSUID | MS_NODEV, opts) != 0) { perror("fzfs: mount"); return 1; }
    g_mountpoint = mountpoint;
    struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_handler = on_signal;
    sigaction(SIGINT, &sa, nullptr); sigaction(SIGTERM, &sa, nullptr);
    fprintf(stderr, "fzfs: %s mounted on %s (codec: %s, level %d, readahead %s, %d serving thread%s)\n", data_dir.c_str(), mountpoint.c_str(), fzfs_codec_name(), level, readahead ? "on" : "off", threads, threads == 1 ? "" : "s");
    fs.loop(threads);
    umount2(mountpoint.c_str(), MNT_DETACH);
    fzfs_codec_shutdown();                             // no readahead thread may still be decoding when the process exits
    return 0;
}
