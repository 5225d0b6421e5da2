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

#include <algorithm>
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

// The compressed files of a batch being formed live in a pinned arena of its own (grow-only, pooled: pinned memory is slow to
// allocate), so the file reads of one batch overlap the decode of another.
struct Arena { uint8_t* p = nullptr; size_t cap = 0; bool busy = false; };
constexpr int kArenas = 8;
constexpr size_t kArenaBytes = (size_t)48 << 20;      // reserved per arena at mount time (a window of 64 x 1 MiB files at ratio 2 needs 32 MiB)
std::mutex g_arena_mu;
std::condition_variable g_arena_cv;
Arena g_arenas[kArenas];
Arena* arena_acquire(size_t need)
{
    std::unique_lock<std::mutex> lk(g_arena_mu);
    Arena* a = nullptr;
    g_arena_cv.wait(lk, [&] {
        for (Arena& x : g_arenas) if (!x.busy && x.cap >= need) { a = &x; return true; }      // one that is large enough already
        for (Arena& x : g_arenas) if (!x.busy) { a = &x; return true; }
        return false;
    });
    a->busy = true;
    lk.unlock();
    if (a->cap < need) {
        if (a->p) cudaFreeHost(a->p);
        a->p = nullptr; a->cap = 0;
        const size_t want = need + need / 4 + (1 << 20);
        if (cudaMallocHost((void**)&a->p, want) != cudaSuccess) { cudaGetLastError(); a->p = nullptr; lk.lock(); a->busy = false; lk.unlock(); g_arena_cv.notify_one(); return nullptr; }
        a->cap = want;
    }
    return a;
}
void arena_release(Arena* a)
{
    { std::lock_guard<std::mutex> lk(g_arena_mu); a->busy = false; }
    g_arena_cv.notify_one();
}

// Group commit of the decode calls.  One 1 MiB frame takes ~2.5 ms on the GPU whether it travels with 30 others or with 500, and
// the windows of sixteen directories are topped up at about the same time: every batch queues its files here, and whoever
// gets the decode lock next takes EVERYTHING that is queued -- its own batch and the ones that arrived while the previous call
// ran -- to the device as one fzg_decode_batch call.  The batch size follows the load.
struct DecodeJob {
    int device;
    std::vector<const void*> sp; std::vector<void*> dp; std::vector<size_t> sl, dc, dl; std::vector<int> st;
    int rc = 0; bool done = false;
};
std::mutex g_decode_mu, g_queue_mu;
std::vector<DecodeJob*> g_queue;
void decode_commit(DecodeJob* mine)
{
    { std::lock_guard<std::mutex> q(g_queue_mu); g_queue.push_back(mine); }
    std::lock_guard<std::mutex> leader(g_decode_mu);
    if (mine->done) return;                                        // a leader before this one took it along
    std::vector<DecodeJob*> jobs;
    { std::lock_guard<std::mutex> q(g_queue_mu); jobs.swap(g_queue); }
    for (size_t a = 0; a < jobs.size();) {                          // one call per device among the queued jobs
        const int device = jobs[a]->device;
        std::vector<DecodeJob*> grp;
        for (size_t b = a; b < jobs.size(); b++) if (jobs[b] && jobs[b]->device == device) { grp.push_back(jobs[b]); jobs[b] = nullptr; }
        if (grp.size() == 1) {
            DecodeJob* j = grp[0];
            j->rc = fzg_decode_batch(device, j->sp.size(), j->sp.data(), j->sl.data(), j->dp.data(), j->dc.data(), j->dl.data(), j->st.data(), 0);
        } else {
            std::vector<const void*> sp; std::vector<void*> dp; std::vector<size_t> sl, dc;
            for (DecodeJob* j : grp) { sp.insert(sp.end(), j->sp.begin(), j->sp.end()); dp.insert(dp.end(), j->dp.begin(), j->dp.end()); sl.insert(sl.end(), j->sl.begin(), j->sl.end()); dc.insert(dc.end(), j->dc.begin(), j->dc.end()); }
            std::vector<size_t> dl(sp.size()); std::vector<int> st(sp.size());
            const int rc = fzg_decode_batch(device, sp.size(), sp.data(), sl.data(), dp.data(), dc.data(), dl.data(), st.data(), 0);
            size_t o = 0;
            for (DecodeJob* j : grp) { j->rc = rc; for (size_t i = 0; i < j->sp.size(); i++) { j->dl[i] = dl[o + i]; j->st[i] = st[o + i]; } o += j->sp.size(); }
        }
        for (DecodeJob* j : grp) j->done = true;
        while (a < jobs.size() && !jobs[a]) a++;
    }
}

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
    std::lock_guard<std::mutex> lk(g_mu);
    if (capacity_bytes != g_capacity) {
        if (!g_pending.empty()) return -EBUSY;                        // a batch in flight is about to write into a slab
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
    std::lock_guard<std::mutex> lk(g_mu);
    size_t total = 0;
    for (const Slab& sl : g_slabs) total += sl.cap;
    while (total + kSlabBytes <= g_capacity) {
        Slab sl;
        if (cudaMallocHost((void**)&sl.p, kSlabBytes) != cudaSuccess) { cudaGetLastError(); return -ENOMEM; }
        sl.cap = kSlabBytes; total += kSlabBytes; g_slabs.push_back(std::move(sl));
    }
    if (g_cur < 0 && !g_slabs.empty()) g_cur = 0;
    // the arenas of the batches and the HBM staging of their decode calls (up to kArenas batches travel together): allocated now,
    // because pinning memory or growing HBM buffers later stalls the batches that are in flight (measured: 0.7 s once, early in a run)
    if (g_capacity) {
        std::lock_guard<std::mutex> al(g_arena_mu);
        for (Arena& a : g_arenas) if (!a.busy && a.cap < kArenaBytes) {
            if (a.p) cudaFreeHost(a.p);
            a.p = nullptr; a.cap = 0;
            if (cudaMallocHost((void**)&a.p, kArenaBytes) != cudaSuccess) { cudaGetLastError(); a.p = nullptr; return -ENOMEM; }
            a.cap = kArenaBytes;
        }
        const size_t per_call = std::min<size_t>(g_capacity, kArenas * kSlabBytes / 4);
        for (int d = 0; d < fzg_device_count(); d++) if (int rc = fzg_reserve_staging(d, 4096, kArenas * kArenaBytes, per_call + (4096u << 5))) return rc;
    }
    return (int)g_slabs.size();
}

// Decodes the listed .zst files that are not cached yet as ONE batch on `device` and keeps the results.  Files that
// cannot be read or do not decode are skipped (open() will report them the ordinary way).  Returns the number of files
// added, or -errno.
// which of the listed files are neither cached nor on their way: they become pending (an open of one of them waits for its batch)
static std::vector<size_t> prefetch_select(const uint64_t* keys, size_t n)
{
    std::vector<size_t> todo;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_capacity == 0) return todo;
    for (size_t i = 0; i < n; i++) if (!g_map.count(keys[i]) && g_pending.insert(keys[i]).second) todo.push_back(i);
    return todo;
}
static int prefetch_run(int device, const char* const* paths, const uint64_t* keys, const std::vector<size_t>& todo);

extern "C" int fzg_cache_prefetch(int device, const char* const* paths, const uint64_t* keys, size_t n)
{
    if (!paths || !keys) return -EINVAL;
    const std::vector<size_t> todo = prefetch_select(keys, n);
    return prefetch_run(device, paths, keys, todo);
}

static int prefetch_run(int device, const char* const* paths, const uint64_t* keys, const std::vector<size_t>& todo)
{
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
    Arena* const arena = arena_acquire(ctotal + 64);
    if (!arena) { for (int fd : fds) if (fd >= 0) close(fd); return -ENOMEM; }
    struct ArenaGuard { Arena* a; ~ArenaGuard() { arena_release(a); } } arena_guard{ arena };
    uint8_t* const abuf = arena->p;
    // the reads are page-cache copies (~6 GB/s on one thread: 14 ms for a directory of 256 x 1 MiB): a few threads share them
    auto read_range = [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; j++) {
            if (fds[j] < 0) continue;
            size_t got = 0;
            while (got < clen[j]) { const ssize_t r = read(fds[j], abuf + coff[j] + got, clen[j] - got); if (r < 0 && errno == EINTR) continue; if (r <= 0) break; got += (size_t)r; }
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
        if (fzg_frame_info(abuf + coff[j], clen[j], &content, &csize) != 0 || content == UINT64_MAX) continue;   // unknown size: left to open()
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
        // (the slab stays pinned while the batch is in flight: another batch must not reuse it before this one has landed)
        { std::lock_guard<std::mutex> lk(g_mu); si = alloc_locked(dtotal ? dtotal : 1, &base); if (si < 0) return added ? added : -ENOMEM; slab_p = g_slabs[si].p; g_slabs[si].pins++; }
        DecodeJob job; job.device = device;
        job.sp.resize(k); job.dp.resize(k); job.sl.resize(k); job.dc.resize(k); job.dl.assign(k, 0); job.st.assign(k, 0);
        for (size_t a = 0; a < k; a++) { const size_t j = which[g0 + a]; job.sp[a] = abuf + coff[j]; job.sl[a] = clen[j]; job.dp[a] = slab_p + base + doff[a]; job.dc[a] = dcs[g0 + a]; }
        const double t_read = ms();
        decode_commit(&job);
        const std::vector<size_t>& dl = job.dl; const std::vector<int>& st = job.st;
        if (trace) fprintf(stderr, "fzgpu: prefetch of %zu files (%zu bytes): up to here %.1f ms, decode (own batch or a combined one) %.1f ms\n", k, dtotal, t_read, ms() - t_read);
        std::unique_lock<std::mutex> lk(g_mu);
        if (si < (int)g_slabs.size() && g_slabs[si].pins > 0) g_slabs[si].pins--;
        if (job.rc) return added ? added : job.rc;
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
    // the files become pending HERE, before this returns: the caller can ask fzg_cache_pending / fzg_cache_wait about them at once
    std::vector<size_t> todo = prefetch_select(keys, n);
    if (todo.empty()) return 0;
    std::vector<std::string> p(n); std::vector<uint64_t> k(keys, keys + n);
    for (size_t i : todo) p[i] = paths[i];
    { std::lock_guard<std::mutex> lk(g_async_mu); g_async_live++; }
    const std::vector<size_t> undo = todo;
    try {
    std::thread([device, p = std::move(p), k = std::move(k), todo = std::move(todo)]() {
        std::vector<const char*> c(p.size());
        for (size_t i = 0; i < p.size(); i++) c[i] = p[i].c_str();
        try { prefetch_run(device, c.data(), k.data(), todo); }
        catch (...) { { std::lock_guard<std::mutex> lk(g_mu); for (size_t i : todo) g_pending.erase(k[i]); } g_cv.notify_all(); }
        { std::lock_guard<std::mutex> lk(g_async_mu); g_async_live--; }
        g_async_cv.notify_all();
    }).detach();
    } catch (...) {                                                  // no thread to be had: nobody must wait for these files
        { std::lock_guard<std::mutex> lk(g_mu); for (size_t i : undo) g_pending.erase(keys[i]); }
        g_cv.notify_all();
        { std::lock_guard<std::mutex> lk(g_async_mu); g_async_live--; }
        g_async_cv.notify_all();
        return -EAGAIN;
    }
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
// Blocks while `key` belongs to a prefetch batch in flight (a caller that serialises its opens waits here, outside its own lock).
extern "C" int fzg_cache_pending(uint64_t key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return g_pending.count(key) ? 1 : 0;
}
extern "C" void fzg_cache_wait(uint64_t key)
{
    std::unique_lock<std::mutex> lk(g_mu);
    g_cv.wait(lk, [&] { return !g_pending.count(key); });
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
