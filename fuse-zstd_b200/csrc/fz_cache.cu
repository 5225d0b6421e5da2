/*
 * fz_cache.cu -- batch formation and the decoded-file cache, host code only.
 *
 * fuse-zstd decodes one file per open(), on one thread (/root/reference/DESIGN.md:5-7,
 * src/main.rs:451-493), and drops the decoded tmpfile with the last handle (src/file.rs:104-117).  The GPU codec is a
 * batch machine: one 1 MiB frame takes ~5 ms whether it travels alone or with ten thousand others.  This component turns
 * the mount's access pattern into batches: when a directory is listed or a file looked up (readdir_wrapper
 * src/main.rs:307-387, lookup_wrapper :215-305) the host hands the sibling .zst paths to fzg_cache_prefetch, which decodes
 * all of them in ONE fzg_decode_batch call and keeps the plain bytes; open_wrapper then calls fzg_cache_open, which is a
 * memcpy + write when the file is cached (and still the file that was decoded: size + mtime are compared) and an
 * ordinary fzg_decode_fd otherwise.  Entries leave in LRU order when the capacity is exceeded, or by
 * fzg_cache_invalidate after a write / rename / unlink (store_to_source_file, src/main.rs:755-832).
 */
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/fzgpu.h"

namespace {

struct Bytes {                                      // plain bytes, NOT zero-filled on allocation (a vector would memset them first)
    std::unique_ptr<uint8_t[]> p; size_t n = 0;
    void alloc(size_t k) { p.reset(new uint8_t[k ? k : 1]); n = k; }
    uint8_t* data() const { return p.get(); }
    size_t size() const { return n; }
};
struct Entry {
    Bytes plain;
    uint64_t src_size; int64_t mtime_ns;           // the source file the bytes were decoded from
    std::list<uint64_t>::iterator lru;
};

std::mutex g_mu;
std::unordered_map<uint64_t, Entry> g_map;         // key = fuse-zstd inode (src/main.rs:744-753)
std::list<uint64_t> g_lru;                         // front = most recently used
size_t g_capacity = (size_t)1 << 30, g_bytes = 0;
uint64_t g_hits = 0, g_misses = 0, g_prefetched = 0;

int64_t mtime_ns(const struct stat& st) { return (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec; }

void evict_locked(size_t need)
{
    while (!g_lru.empty() && g_bytes + need > g_capacity) {
        const uint64_t k = g_lru.back(); g_lru.pop_back();
        auto it = g_map.find(k);
        if (it != g_map.end()) { g_bytes -= it->second.plain.size(); g_map.erase(it); }
    }
}

int read_file(const char* path, std::vector<uint8_t>& buf, struct stat& st)
{
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return -errno;
    if (fstat(fd, &st) != 0) { const int e = -errno; close(fd); return e; }
    buf.resize((size_t)st.st_size);
    size_t got = 0;
    while (got < buf.size()) {
        const ssize_t r = read(fd, buf.data() + got, buf.size() - got);
        if (r < 0) { if (errno == EINTR) continue; const int e = -errno; close(fd); return e; }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    buf.resize(got);
    return 0;
}

int write_all(int fd, const uint8_t* p, size_t n)
{
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return -errno; }
        p += w; n -= (size_t)w;
    }
    return 0;
}

}  // namespace

extern "C" int fzg_cache_configure(size_t capacity_bytes)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_capacity = capacity_bytes;
    evict_locked(0);
    return 0;
}

// Decodes the listed .zst files that are not cached yet as ONE batch on `device` and keeps the results.  Files that
// cannot be read or do not decode are skipped (open() will report them the ordinary way).  Returns the number of files
// added, or -errno.
extern "C" int fzg_cache_prefetch(int device, const char* const* paths, const uint64_t* keys, size_t n)
{
    if (!paths || !keys) return -EINVAL;
    std::vector<size_t> todo;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_capacity == 0) return 0;
        for (size_t i = 0; i < n; i++) if (!g_map.count(keys[i])) todo.push_back(i);
    }
    if (todo.empty()) return 0;
    const size_t m = todo.size();
    std::vector<std::vector<uint8_t>> comp(m); std::vector<Bytes> plain(m);
    std::vector<struct stat> sts(m);
    std::vector<const void*> sp; std::vector<void*> dp; std::vector<size_t> sl, dc, which;
    size_t budget = 0;
    for (size_t j = 0; j < m; j++) {
        if (read_file(paths[todo[j]], comp[j], sts[j]) != 0 || comp[j].empty()) continue;
        uint64_t content = 0, csize = 0;
        if (fzg_frame_info(comp[j].data(), comp[j].size(), &content, &csize) != 0 || content == UINT64_MAX) continue;   // unknown size: left to open()
        if (budget + content > g_capacity) break;              // a prefetch never evicts more than the cache holds
        budget += content;
        plain[j].alloc(content);
        sp.push_back(comp[j].data()); sl.push_back(comp[j].size()); dp.push_back(plain[j].data()); dc.push_back(content); which.push_back(j);
    }
    const size_t k = which.size();
    if (k == 0) return 0;
    std::vector<size_t> dl(k); std::vector<int> st(k);
    const int rc = fzg_decode_batch(device, k, sp.data(), sl.data(), dp.data(), dc.data(), dl.data(), st.data(), 0);
    if (rc) return rc;
    int added = 0;
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t a = 0; a < k; a++) {
        if (st[a] != 0) continue;
        const size_t j = which[a]; const uint64_t key = keys[todo[j]];
        if (g_map.count(key)) continue;
        plain[j].n = dl[a];
        evict_locked(plain[j].size());
        if (g_bytes + plain[j].size() > g_capacity) continue;
        g_lru.push_front(key);
        Entry e; e.plain = std::move(plain[j]); e.src_size = (uint64_t)sts[j].st_size; e.mtime_ns = mtime_ns(sts[j]); e.lru = g_lru.begin();
        g_bytes += e.plain.size();
        g_map.emplace(key, std::move(e));
        added++; g_prefetched++;
    }
    return added;
}

// The same, on a detached thread: the FUSE loop (one thread, src/main.rs:1325) does not wait for the batch.
extern "C" int fzg_cache_prefetch_async(int device, const char* const* paths, const uint64_t* keys, size_t n)
{
    if (!paths || !keys) return -EINVAL;
    std::vector<std::string> p(n); std::vector<uint64_t> k(keys, keys + n);
    for (size_t i = 0; i < n; i++) p[i] = paths[i];
    std::thread([device, p = std::move(p), k = std::move(k)]() {
        std::vector<const char*> c(p.size());
        for (size_t i = 0; i < p.size(); i++) c[i] = p[i].c_str();
        fzg_cache_prefetch(device, c.data(), k.data(), c.size());
    }).detach();
    return 0;
}

// open_wrapper's codec call (src/main.rs:463-467) with the cache in front: the plain bytes of `src_fd` (a .zst file whose
// inode is `key`) are written to dst_fd at its current offset.  Served from the cache when the cached entry was decoded from
// a file of the same size and mtime; otherwise decoded by fzg_decode_fd.  *hit (optional) reports which.
extern "C" int fzg_cache_open(int src_fd, int dst_fd, uint64_t key, uint64_t* out_size, int* hit)
{
    struct stat st;
    if (fstat(src_fd, &st) != 0) return -errno;
    {
        std::unique_lock<std::mutex> lk(g_mu);
        auto it = g_map.find(key);
        if (it != g_map.end() && it->second.src_size == (uint64_t)st.st_size && it->second.mtime_ns == mtime_ns(st)) {
            g_lru.erase(it->second.lru); g_lru.push_front(key); it->second.lru = g_lru.begin();
            g_hits++;
            const size_t sz = it->second.plain.size();
            const int rc = write_all(dst_fd, it->second.plain.data(), sz);      // under the lock: the entry cannot be evicted meanwhile
            lk.unlock();
            if (rc) return rc;
            if (lseek(src_fd, 0, SEEK_END) < 0) return -errno;  // copy_decode leaves the source at its end
            if (out_size) *out_size = sz;
            if (hit) *hit = 1;
            return 0;
        }
        if (it != g_map.end()) { g_bytes -= it->second.plain.size(); g_lru.erase(it->second.lru); g_map.erase(it); }   // stale
        g_misses++;
    }
    if (hit) *hit = 0;
    return fzg_decode_fd(src_fd, dst_fd, key, out_size);
}

extern "C" int fzg_cache_invalidate(uint64_t key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_map.find(key);
    if (it == g_map.end()) return 0;
    g_bytes -= it->second.plain.size(); g_lru.erase(it->second.lru); g_map.erase(it);
    return 1;
}

extern "C" void fzg_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* files)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (hits) *hits = g_hits;
    if (misses) *misses = g_misses;
    if (bytes) *bytes = g_bytes;
    if (files) *files = g_map.size();
}
