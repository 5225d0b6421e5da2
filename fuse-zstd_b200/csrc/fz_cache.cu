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
 * ordinary fzg_decode_fd otherwise.  The plain bytes live in pinned slabs that are filled batch after batch and reused
 * oldest-first once the capacity is reached (everything in the reused slab leaves together); single entries leave by
 * fzg_cache_invalidate after a write / rename / unlink (store_to_source_file, src/main.rs:755-832).
 */
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/fzgpu.h"

namespace {

// Plain bytes live in PINNED slabs (cudaMallocHost is slow -- ~0.4 ms per MiB -- so slabs are allocated once, up to the
// capacity, and then reused oldest-first): a batch's results are contiguous in one slab, so its device -> host copy is one DMA
// at PCIe speed instead of a staged copy per file into freshly faulted pageable memory, which made the host side of a batch
// ten times longer than its kernels.
constexpr size_t kSlabBytes = (size_t)256 << 20;
struct Slab {
    uint8_t* p = nullptr; size_t cap = 0, used = 0;
    int pins = 0;                                    // views handed out by fzg_cache_view: a pinned slab is never reused
    std::vector<uint64_t> keys;                      // entries that live here (stale keys are harmless: see evict_slab)
};
struct Entry {
    int slab = 0; size_t off = 0, n = 0; uint64_t id = 0;
    uint64_t src_size; int64_t mtime_ns;            // the source file the bytes were decoded from
};

std::mutex g_mu;
std::condition_variable g_cv;
std::unordered_set<uint64_t> g_pending;            // keys of the batches in flight: an open of one of them waits for its batch
std::unordered_map<uint64_t, Entry> g_map;         // key = fuse-zstd inode (src/main.rs:744-753)
std::vector<Slab> g_slabs;
int g_cur = -1;                                    // slab being filled
size_t g_capacity = (size_t)1 << 30, g_bytes = 0;
uint64_t g_hits = 0, g_misses = 0, g_prefetched = 0, g_next_id = 1;

std::mutex g_batch_mu;                             // one prefetch batch at a time (they share the compressed-input arena)
uint8_t* g_arena = nullptr; size_t g_arena_cap = 0;   // pinned, grow-only: the compressed files of the batch being formed

void evict_slab_locked(int si)
{
    Slab& sl = g_slabs[si];
    for (uint64_t k : sl.keys) { auto it = g_map.find(k); if (it != g_map.end() && it->second.slab == si) { g_bytes -= it->second.n; g_map.erase(it); } }
    sl.keys.clear(); sl.used = 0;
}
void drop_all_locked()
{
    for (Slab& sl : g_slabs) if (sl.p) cudaFreeHost(sl.p);
    g_slabs.clear(); g_map.clear(); g_cur = -1; g_bytes = 0;
}
// `need` contiguous pinned bytes for one batch: the current slab, a new slab while the capacity allows, else the oldest slab
// (whose entries leave).  Returns the slab index or -1.
int alloc_locked(size_t need, size_t* off)
{
    if (g_cur >= 0 && g_slabs[g_cur].used + need <= g_slabs[g_cur].cap) { *off = g_slabs[g_cur].used; g_slabs[g_cur].used += need; return g_cur; }
    size_t total = 0;
    for (const Slab& sl : g_slabs) total += sl.cap;
    const size_t want = need > kSlabBytes ? need : (g_capacity < kSlabBytes ? (need > g_capacity ? need : g_capacity) : kSlabBytes);
    for (size_t t = 1; t < g_slabs.size(); t++) {                  // a reserved slab that has not been used yet
        const int c = (int)((g_cur + t) % g_slabs.size());
        if (g_slabs[c].used == 0 && g_slabs[c].keys.empty() && g_slabs[c].cap >= need) { g_cur = c; *off = 0; g_slabs[c].used = need; return c; }
    }
    if (g_slabs.empty() || total + want <= g_capacity) {
        Slab sl;
        if (cudaMallocHost((void**)&sl.p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        sl.cap = want; g_slabs.push_back(std::move(sl)); g_cur = (int)g_slabs.size() - 1;
    } else {
        int next = -1;
        for (size_t t = 1; t <= g_slabs.size(); t++) { const int c = (int)((g_cur + t) % g_slabs.size()); if (g_slabs[c].cap >= need && g_slabs[c].pins == 0) { next = c; break; } }
        if (next < 0) return -1;
        g_cur = next; evict_slab_locked(g_cur);
    }
    *off = 0; g_slabs[g_cur].used = need;
    return g_cur;
}

int64_t mtime_ns(const struct stat& st) { return (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec; }

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
    std::lock_guard<std::mutex> batch(g_batch_mu);
    std::lock_guard<std::mutex> lk(g_mu);
    if (capacity_bytes != g_capacity) {
        for (const Slab& sl : g_slabs) if (sl.pins) return -EBUSY;  // a view still points into a slab
        drop_all_locked();                                        // the slabs are sized for a capacity: a new one starts empty
    }
    g_capacity = capacity_bytes;
    return 0;
}

// Allocates the pinned slabs up to the configured capacity now (cudaMallocHost costs ~1 ms per MiB on this box: a mount does
// it once at start instead of inside its first batches).  Returns the number of slabs, or -errno.
extern "C" int fzg_cache_reserve(void)
{
    std::lock_guard<std::mutex> batch(g_batch_mu);
    std::lock_guard<std::mutex> lk(g_mu);
    size_t total = 0;
    for (const Slab& sl : g_slabs) total += sl.cap;
    while (total + kSlabBytes <= g_capacity) {
        Slab sl;
        if (cudaMallocHost((void**)&sl.p, kSlabBytes) != cudaSuccess) { cudaGetLastError(); return -ENOMEM; }
        sl.cap = kSlabBytes; total += kSlabBytes; g_slabs.push_back(std::move(sl));
    }
    if (g_cur < 0 && !g_slabs.empty()) g_cur = 0;
    return (int)g_slabs.size();
}

// Decodes the listed .zst files that are not cached yet as ONE batch on `device` and keeps the results.  Files that
// cannot be read or do not decode are skipped (open() will report them the ordinary way).  Returns the number of files
// added, or -errno.
extern "C" int fzg_cache_prefetch(int device, const char* const* paths, const uint64_t* keys, size_t n)
{
    if (!paths || !keys) return -EINVAL;
    std::lock_guard<std::mutex> batch(g_batch_mu);
    std::vector<size_t> todo;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_capacity == 0) return 0;
        for (size_t i = 0; i < n; i++) if (!g_map.count(keys[i]) && g_pending.insert(keys[i]).second) todo.push_back(i);
    }
    if (todo.empty()) return 0;
    struct Done {                                      // whatever happens, the keys stop being pending and waiters wake up
        const std::vector<size_t>& todo; const uint64_t* keys;
        ~Done() { { std::lock_guard<std::mutex> lk(g_mu); for (size_t i : todo) g_pending.erase(keys[i]); } g_cv.notify_all(); }
    } done{ todo, keys };
    static const bool trace = getenv("FZG_TRACE") != nullptr;
    struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    auto ms = [&]() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6; };
    const size_t m = todo.size();
    // ---- the compressed files, one after the other in the pinned arena (the batch is then ONE host -> device copy)
    std::vector<struct stat> sts(m); std::vector<int> fds(m, -1); std::vector<size_t> coff(m, 0), clen(m, 0);
    size_t ctotal = 0;
    for (size_t j = 0; j < m; j++) {
        fds[j] = open(paths[todo[j]], O_RDONLY | O_CLOEXEC);
        if (fds[j] < 0 || fstat(fds[j], &sts[j]) != 0 || sts[j].st_size == 0) { if (fds[j] >= 0) close(fds[j]); fds[j] = -1; continue; }
        coff[j] = ctotal; clen[j] = (size_t)sts[j].st_size; ctotal += (clen[j] + 15 & ~(size_t)15) + 16;
    }
    if (ctotal + 64 > g_arena_cap) {
        if (g_arena) cudaFreeHost(g_arena);
        g_arena = nullptr; g_arena_cap = 0;
        const size_t want = ctotal + ctotal / 4 + (1 << 20);
        if (cudaMallocHost((void**)&g_arena, want) != cudaSuccess) { cudaGetLastError(); g_arena = nullptr; for (int fd : fds) if (fd >= 0) close(fd); return -ENOMEM; }
        g_arena_cap = want;
    }
    // the reads are page-cache copies (~6 GB/s on one thread: 14 ms for a directory of 256 x 1 MiB): a few threads share them
    auto read_range = [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; j++) {
            if (fds[j] < 0) continue;
            size_t got = 0;
            while (got < clen[j]) { const ssize_t r = read(fds[j], g_arena + coff[j] + got, clen[j] - got); if (r < 0 && errno == EINTR) continue; if (r <= 0) break; got += (size_t)r; }
            close(fds[j]);
            if (got != clen[j]) fds[j] = -1;
        }
    };
    {
        const size_t nt = ctotal >= (8u << 20) ? std::min<size_t>(4, m) : 1;
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++) th.emplace_back(read_range, m * t / nt, m * (t + 1) / nt);
        read_range(0, m / nt);
        for (auto& x : th) x.join();
    }
    // ---- sizes from the frame headers.  The results of a GROUP of files are contiguous in one pinned slab (one device -> host
    // copy per group); a group never asks for more than a slab holds, so a directory larger than a slab is decoded as
    // several batches instead of not at all (the slabs are reserved at mount time with kSlabBytes each).
    size_t limit, budget;                                  // per group / for the whole call: a prefetch never takes more than the cache holds
    {
        std::lock_guard<std::mutex> lk(g_mu);
        budget = g_capacity;
        limit = g_capacity < kSlabBytes ? g_capacity : kSlabBytes;
        for (const Slab& sl : g_slabs) if (sl.cap > limit && sl.cap <= g_capacity) limit = sl.cap;
    }
    std::vector<size_t> which, dcs;                        // files with a declared size that the cache can hold
    for (size_t j = 0; j < m; j++) {
        if (fds[j] < 0) continue;
        uint64_t content = 0, csize = 0;
        if (fzg_frame_info(g_arena + coff[j], clen[j], &content, &csize) != 0 || content == UINT64_MAX) continue;   // unknown size: left to open()
        if (content > limit) continue;                     // larger than a slab (or a lying header): left to open(), the others go on
        if (content > budget) break;                       // the cache is full of this very call's files
        budget -= (size_t)content;
        which.push_back(j); dcs.push_back((size_t)content);
    }
    int added = 0;
    for (size_t g0 = 0; g0 < which.size();) {
        size_t g1 = g0, dtotal = 0;
        std::vector<size_t> doff;
        while (g1 < which.size() && dcs[g1] <= limit - dtotal) { doff.push_back(dtotal); dtotal += dcs[g1]; g1++; }
        const size_t k = g1 - g0;                          // >= 1: every file fits a slab by itself
        int si; size_t base = 0; uint8_t* slab_p;
        { std::lock_guard<std::mutex> lk(g_mu); si = alloc_locked(dtotal ? dtotal : 1, &base); if (si < 0) return added ? added : -ENOMEM; slab_p = g_slabs[si].p; }
        std::vector<const void*> sp(k); std::vector<void*> dp(k); std::vector<size_t> sl(k), dc(k), dl(k); std::vector<int> st(k);
        for (size_t a = 0; a < k; a++) { const size_t j = which[g0 + a]; sp[a] = g_arena + coff[j]; sl[a] = clen[j]; dp[a] = slab_p + base + doff[a]; dc[a] = dcs[g0 + a]; }
        const double t_read = ms();
        const int rc = fzg_decode_batch(device, k, sp.data(), sl.data(), dp.data(), dc.data(), dl.data(), st.data(), 0);
        if (rc) return added ? added : rc;
        if (trace) fprintf(stderr, "fzgpu: prefetch of %zu files (%zu bytes): up to here %.1f ms, decode batch %.1f ms\n", k, dtotal, t_read, ms() - t_read);
        std::unique_lock<std::mutex> lk(g_mu);
        for (size_t a = 0; a < k; a++) {
            if (st[a] != 0) continue;
            const size_t j = which[g0 + a]; const uint64_t key = keys[todo[j]];
            if (g_map.count(key) || si >= (int)g_slabs.size()) continue;
            Entry e; e.slab = si; e.off = base + doff[a]; e.n = dl[a]; e.id = g_next_id++; e.src_size = (uint64_t)sts[j].st_size; e.mtime_ns = mtime_ns(sts[j]);
            g_bytes += e.n;
            g_slabs[si].keys.push_back(key);
            g_map.emplace(key, e);
            added++; g_prefetched++;
        }
        g0 = g1;
    }
    return added;
}

// Prefetch threads in flight: fzg_cache_drain (fzg_shutdown, the fzfs daemon before it exits) waits for them, so that none
// is inside a CUDA call when the process tears the runtime down.
static std::mutex g_async_mu;
static std::condition_variable g_async_cv;
static int g_async_live = 0;
extern "C" void fzg_cache_drain(void)
{
    std::unique_lock<std::mutex> lk(g_async_mu);
    g_async_cv.wait(lk, [] { return g_async_live == 0; });
}

// The same, on a detached thread: the FUSE loop (one thread, src/main.rs:1325) does not wait for the batch.
extern "C" int fzg_cache_prefetch_async(int device, const char* const* paths, const uint64_t* keys, size_t n)
{
    if (!paths || !keys) return -EINVAL;
    std::vector<std::string> p(n); std::vector<uint64_t> k(keys, keys + n);
    for (size_t i = 0; i < n; i++) p[i] = paths[i];
    { std::lock_guard<std::mutex> lk(g_async_mu); g_async_live++; }
    std::thread([device, p = std::move(p), k = std::move(k)]() {
        std::vector<const char*> c(p.size());
        for (size_t i = 0; i < p.size(); i++) c[i] = p[i].c_str();
        try { fzg_cache_prefetch(device, c.data(), k.data(), c.size()); } catch (...) { }
        { std::lock_guard<std::mutex> lk(g_async_mu); g_async_live--; }
        g_async_cv.notify_all();
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
        g_cv.wait(lk, [&] { return !g_pending.count(key); });      // its batch is in flight: waiting costs less than a decode of its own
        auto it = g_map.find(key);
        if (it != g_map.end() && it->second.src_size == (uint64_t)st.st_size && it->second.mtime_ns == mtime_ns(st)) {
            g_hits++;
            const size_t sz = it->second.n;
            const int rc = write_all(dst_fd, g_slabs[it->second.slab].p + it->second.off, sz);   // under the lock: its slab cannot be reused meanwhile
            lk.unlock();
            if (rc) return rc;
            if (lseek(src_fd, 0, SEEK_END) < 0) return -errno;  // copy_decode leaves the source at its end
            if (out_size) *out_size = sz;
            if (hit) *hit = 1;
            return 0;
        }
        if (it != g_map.end()) { g_bytes -= it->second.n; g_map.erase(it); }   // stale
        g_misses++;
    }
    if (hit) *hit = 0;
    return fzg_decode_fd(src_fd, dst_fd, key, out_size);
}


// The plain bytes of a cached file IN PLACE: a read-only open needs no tmpfile at all (open_wrapper copies the decoded bytes into
// one, src/main.rs:462-466, and every read then copies them out again).  The slab that holds the entry is pinned until
// fzg_cache_unview; the entry may be invalidated meanwhile, the bytes stay (a reader keeps what it opened, as with the
// reference's tmpfile).  -ENOENT: not cached (or stale): the caller decodes the ordinary way.
extern "C" int fzg_cache_view(int src_fd, uint64_t key, const void** data, uint64_t* size, void** token)
{
    if (!data || !size || !token) return -EINVAL;
    struct stat st;
    if (fstat(src_fd, &st) != 0) return -errno;
    std::unique_lock<std::mutex> lk(g_mu);
    g_cv.wait(lk, [&] { return !g_pending.count(key); });
    auto it = g_map.find(key);
    if (it == g_map.end() || it->second.src_size != (uint64_t)st.st_size || it->second.mtime_ns != mtime_ns(st)) return -ENOENT;
    Slab& sl = g_slabs[it->second.slab];
    sl.pins++; g_hits++;
    *data = sl.p + it->second.off; *size = it->second.n; *token = (void*)(uintptr_t)(it->second.slab + 1);
    return 0;
}
extern "C" void fzg_cache_unview(void* token)
{
    const size_t si = (size_t)(uintptr_t)token;
    std::lock_guard<std::mutex> lk(g_mu);
    if (si >= 1 && si <= g_slabs.size() && g_slabs[si - 1].pins > 0) g_slabs[si - 1].pins--;
}

extern "C" int fzg_cache_invalidate(uint64_t key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_map.find(key);
    if (it == g_map.end()) return 0;
    g_bytes -= it->second.n; g_map.erase(it);
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
