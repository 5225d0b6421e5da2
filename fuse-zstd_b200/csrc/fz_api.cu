/*
 * fz_api.cu -- the C ABI of libfzgpu.so (include/fzgpu.h): contexts, host<->device staging and
 * the fd entry points that replace fuse-zstd's two codec call sites
 * (/root/reference/src/main.rs:463-467 and :781-791).  No CPU codec path exists in this
 * library: every compute entry point needs a CUDA device and fails with -ENODEV without one.
 */
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "fz_core.cuh"
#include "fz_kernels.cuh"
#include "fz_host.h"

using namespace fz;

int fzh_encode_setup(void);
int fzh_encode_run(FzCtx* c, uint32_t first, uint32_t n, int level, size_t chunk, int flags);
size_t fzh_encode_bound(size_t src_len, size_t chunk);
const char* fzh_encode_stage_name(int s);

static std::mutex g_mu;
static std::vector<FzCtx*> g_ctx;       // index = position in the init list
static std::vector<int> g_dev;

#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "fzgpu: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return -EIO; } } while (0)

extern "C" int fzg_init(const int* devices, int n_devices)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx.empty()) return 0;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) { cudaGetLastError(); return -ENODEV; }
    std::vector<int> devs;
    if (n_devices <= 0) for (int i = 0; i < visible; i++) devs.push_back(i);
    else for (int i = 0; i < n_devices; i++) devs.push_back(devices ? devices[i] : i);
    for (int d : devs) if (d < 0 || d >= visible) return -EINVAL;
    for (int d : devs) {
        CKR(cudaSetDevice(d));
        if (const char* g = getenv("FZG_L2_GRAN")) {                                                                       // experiment knob
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)); size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            fprintf(stderr, "fzgpu: L2 fetch granularity: asked %s, set -> %s, now %zu\n", g, cudaGetErrorString(e), got);
        }
        if (const char* g = getenv("FZG_L2_PERSIST_MB")) {                                                                 // experiment knob (FZ_EXEC_L2OUT)
            int mx = 0; cudaDeviceGetAttribute(&mx, cudaDevAttrMaxPersistingL2CacheSize, d);
            const size_t want = std::min<size_t>((size_t)atoi(g) << 20, (size_t)mx);
            cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want); size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
            fprintf(stderr, "fzgpu: persisting L2: max %d, asked %zu, set -> %s, now %zu\n", mx, want, cudaGetErrorString(e), got);
        }
        FzCtx* c = new FzCtx();
        c->dev = d;
        const char* nl = getenv("FZG_LANES");
        c->n_lanes = nl ? std::max(1, std::min(kMaxLanes, atoi(nl))) : 1;   // measured: no gain from 2-4 lanes on B200 (profiles/r01_notes.md)
        for (int l = 0; l < kMaxLanes; l++) {
            CKR(cudaStreamCreateWithFlags(&c->lane[l].stream, cudaStreamNonBlocking));
            for (auto& e : c->lane[l].ev) CKR(cudaEventCreate(&e));
            CKR(cudaEventCreateWithFlags(&c->lane[l].ev_entropy, cudaEventDisableTiming));
            CKR(cudaEventCreateWithFlags(&c->lane[l].ev_prog, cudaEventDisableTiming));
            { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);      // lowest priority: its CTAs take what the main stream leaves
              CKR(cudaStreamCreateWithPriority(&c->lane[l].side, cudaStreamNonBlocking, lo)); }
            CKR(cudaEventCreateWithFlags(&c->lane[l].ev_fork, cudaEventDisableTiming));
            CKR(cudaEventCreateWithFlags(&c->lane[l].ev_join, cudaEventDisableTiming));
        }
        c->stream = c->lane[0].stream;
        CKR(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CKR(cudaStreamCreateWithFlags(&c->copy_stream2, cudaStreamNonBlocking));
        CKR(cudaStreamCreateWithFlags(&c->poll_stream, cudaStreamNonBlocking));
        for (auto& e : c->ev) CKR(cudaEventCreate(&e));
        int rc = fzh_decode_setup(); if (rc) return rc;
        rc = fzh_encode_setup(); if (rc) return rc;
        g_ctx.push_back(c); g_dev.push_back(d);
    }
    return 0;
}

extern "C" void fzg_shutdown(void)
{
    fzg_cache_drain();                                 // no prefetch thread may be inside a CUDA call while the contexts go
    std::lock_guard<std::mutex> lk(g_mu);
    for (FzCtx* c : g_ctx) {
        cudaSetDevice(c->dev);
        cudaStreamSynchronize(c->stream);
        FzDevBuf* db[] = { &c->d_stage_src, &c->d_stage_dst, &c->e_items, &c->e_outs, &c->e_work, &c->e_tab, &c->d_outs, &c->d_totals };
        for (auto* b : db) b->release();
        FzPinBuf* pb[] = { &c->h_items, &c->h_outs, &c->h_totals, &c->h_stage_src, &c->h_stage_dst, &c->e_chunks_h, &c->e_first_h };
        for (auto* b : pb) b->release();
        for (int l = 0; l < kMaxLanes; l++) {
            FzLane& L = c->lane[l];
            cudaStreamSynchronize(L.stream);
            FzDevBuf* lb[] = { &L.d_infos, &L.d_bases, &L.d_outs, &L.d_totals, &L.d_frames, &L.d_blocks, &L.d_seq_jobs, &L.d_huf_jobs,
                               &L.d_lit, &L.d_seq, &L.d_seq_tabs, &L.d_seq_hdrs, &L.d_prog };
            for (auto* b : lb) b->release();
            L.h_totals.release();
            for (auto& e : L.ev) cudaEventDestroy(e);
            cudaEventDestroy(L.ev_entropy); cudaEventDestroy(L.ev_prog); L.h_prog.release();
            cudaEventDestroy(L.ev_fork); cudaEventDestroy(L.ev_join);
            cudaStreamDestroy(L.side);
            cudaStreamDestroy(L.stream);
        }
        for (auto& e : c->ev) cudaEventDestroy(e);
        cudaStreamDestroy(c->copy_stream); cudaStreamDestroy(c->copy_stream2); cudaStreamDestroy(c->poll_stream);
        delete c;
    }
    g_ctx.clear(); g_dev.clear();
}

extern "C" int fzg_device_count(void) { std::lock_guard<std::mutex> lk(g_mu); return (int)g_ctx.size(); }

static FzCtx* ctx_for(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_ctx.size(); i++) if (g_dev[i] == device) return g_ctx[i];
    return nullptr;
}
static FzCtx* ctx_for_key(uint64_t key)
{
    std::lock_guard<std::mutex> lk(g_mu);
    return g_ctx.empty() ? nullptr : g_ctx[key % g_ctx.size()];
}
static int ensure_init(void)
{
    { std::lock_guard<std::mutex> lk(g_mu); if (!g_ctx.empty()) return 0; }
    return fzg_init(nullptr, 0);
}

static bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static inline size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }

// ---------------------------------------------------------------------------------- batched decode / encode
// Device-resident batches run as one launch sequence.  Host-resident batches are cut into chunks of
// ~512 MiB of (input + output) and pipelined on three streams: while chunk c is decoded, chunk c+1 crosses PCIe
// host->device and chunk c-1 device->host, so the call costs about max(PCIe, kernels) instead of their sum.
static size_t chunk_bytes()                             // (input + output) bytes per chunk of a host-resident batch
{
    static const size_t v = [] { const char* e = getenv("FZG_CHUNK_MB"); return e && atoi(e) > 0 ? (size_t)atoi(e) << 20 : (size_t)512 << 20; }();
    return v;
}

static size_t stream_out_min_bytes()                     // output bytes of a chunk from which its device -> host copy follows the execute stage
{
    const char* e = getenv("FZG_STREAM_OUT_MB");           // read per call (tests); "0" streams every eligible chunk, a huge value none
    return e ? (size_t)atoll(e) << 20 : (size_t)256 << 20;
}

static int run_batch(bool encode, int device, size_t n, const void* const* src, const size_t* src_len, void* const* dst,
                     const size_t* dst_cap, size_t* dst_len, int* status, int flags, int level, size_t chunk)
{
    if (n == 0) return 0;
    if (!src || !src_len || !dst || !dst_cap || !dst_len || !status) return -EINVAL;
    if (n >= (1ull << 31)) return -EINVAL;
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for(device);
    if (!c) return -ENODEV;
    std::lock_guard<std::mutex> lk(c->mu);
    CKR(cudaSetDevice(c->dev));
    const bool src_dev = flags & FZG_SRC_DEVICE, dst_dev = flags & FZG_DST_DEVICE;
    if ((rc = c->h_items.reserve(n * sizeof(Item)))) return rc;
    if ((rc = c->h_outs.reserve(n * sizeof(ItemOut)))) return rc;
    if ((rc = c->h_totals.reserve(64))) return rc;
    Item* items = (Item*)c->h_items.p;
    const ItemOut* outs = (const ItemOut*)c->h_outs.p;
    for (size_t i = 0; i < n; i++) { items[i].src_len = src_len[i]; items[i].dst_cap = dst_cap[i]; }
    fzg_timing_t acc = {};

    // ---- device layout of the sources
    const uint8_t* h2d_base = nullptr;            // pinned host region mirrored at d_stage_src
    bool src_scatter = false;                     // or: every item in pinned memory of its own
    if (!src_dev) {
        // a caller's pinned buffer whose items lie in increasing order is copied as it is (gaps included)
        bool span_ok = src_len[0] && is_pinned(src[0]) && is_pinned((const uint8_t*)src[n - 1] + (src_len[n - 1] ? src_len[n - 1] - 1 : 0));
        size_t sum = 0;
        for (size_t i = 0; i < n && span_ok; i++) {
            sum += src_len[i];
            if (i + 1 < n) {                                  // increasing, and close together: a large gap is where another allocation may begin
                const uint8_t* end = (const uint8_t*)src[i] + src_len[i];
                if (end > (const uint8_t*)src[i + 1] || (size_t)((const uint8_t*)src[i + 1] - end) > ((size_t)256 << 10)) span_ok = false;
            }
        }
        const size_t span = span_ok ? (size_t)((const uint8_t*)src[n - 1] + src_len[n - 1] - (const uint8_t*)src[0]) : 0;
        if (span_ok && span <= 2 * sum + (64u << 10) * n) {
            if ((rc = c->d_stage_src.reserve(span + 64))) return rc;
            h2d_base = (const uint8_t*)src[0];
            for (size_t i = 0; i < n; i++) items[i].src = (uint8_t*)c->d_stage_src.p + ((const uint8_t*)src[i] - h2d_base);
        } else {
            size_t total = 0;
            for (size_t i = 0; i < n; i++) total += al16(src_len[i]) + 16;
            if ((rc = c->d_stage_src.reserve(total + 64))) return rc;
            // pinned buffers that do not form one span (the cache's batches come from several arenas): a copy per item, no packing
            src_scatter = is_pinned(src[0]);
            for (size_t i = 1; i < n && src_scatter; i++) src_scatter = is_pinned(src[i]);
            uint8_t* hbase = nullptr;
            if (!src_scatter) {                   // pageable: pack into pinned staging
                if ((rc = c->h_stage_src.reserve(total + 64))) return rc;
                hbase = (uint8_t*)c->h_stage_src.p;
            }
            size_t off = 0;
            for (size_t i = 0; i < n; i++) {
                if (hbase && src_len[i]) memcpy(hbase + off, src[i], src_len[i]);
                items[i].src = (uint8_t*)c->d_stage_src.p + off; off += al16(src_len[i]) + 16;
            }
            h2d_base = hbase;
        }
    } else for (size_t i = 0; i < n; i++) items[i].src = (const uint8_t*)src[i];

    // ---- device layout of the destinations
    bool dst_contig = false;
    if (!dst_dev) {
        dst_contig = is_pinned(dst[0]);
        for (size_t i = 0; i + 1 < n && dst_contig; i++) dst_contig = (uint8_t*)dst[i] + dst_cap[i] == (uint8_t*)dst[i + 1];
        size_t total = 0;
        for (size_t i = 0; i < n; i++) {
            const size_t add = dst_contig ? dst_cap[i] : al16(dst_cap[i]) + 16;
            if (add < dst_cap[i] || total > SIZE_MAX / 2 - add) return -EINVAL;      // capacities that do not add up to a size
            total += add;
        }
        if ((rc = c->d_stage_dst.reserve(total + 64))) return rc;
        uint8_t* dbase = (uint8_t*)c->d_stage_dst.p; size_t off = 0;
        for (size_t i = 0; i < n; i++) { items[i].dst = dbase + off; off += dst_contig ? dst_cap[i] : al16(dst_cap[i]) + 16; }
    } else for (size_t i = 0; i < n; i++) items[i].dst = (uint8_t*)dst[i];

    auto add_timing = [&](const fzg_timing_t& t, bool concurrent) {
        acc.total_ms = concurrent ? std::max(acc.total_ms, t.total_ms) : acc.total_ms + t.total_ms;
        acc.launches += t.launches;
        for (int k = 0; k < 16; k++) acc.kernel_ms[k] += t.kernel_ms[k];
    };
    // items [first, first + cnt): the encoder runs on the context's stream; the decoder splits the range over the
    // context's lanes (by compressed bytes), one host thread per lane, so that different stages of different slices overlap
    auto run = [&](size_t first, size_t cnt, cudaEvent_t after) -> int {     // `after`: the chunk's H2D copy (or null)
        if (after) for (int l = 0; l < (encode ? 1 : kMaxLanes); l++) CKR(cudaStreamWaitEvent(c->lane[l].stream, after, 0));
        if (encode) {
            int r = fzh_encode_run(c, (uint32_t)first, (uint32_t)cnt, level, chunk, flags);
            if (r) return r;
            add_timing(c->timing, false);
            return 0;
        }
        const int lanes = cnt >= 64 ? c->n_lanes : 1;
        if (lanes == 1) {
            int r = fzh_decode_run(c, 0, (uint32_t)first, (uint32_t)cnt, flags, false);
            if (r) return r;
            add_timing(c->lane[0].timing, false);
            return 0;
        }
        uint64_t total = 0;
        for (size_t i = first; i < first + cnt; i++) total += src_len[i] + 1;
        size_t cut[kMaxLanes + 1]; cut[0] = first; cut[lanes] = first + cnt;
        { uint64_t run_b = 0; int l = 1;
          for (size_t i = first; i < first + cnt && l < lanes; i++) { run_b += src_len[i] + 1; if (run_b * lanes >= total * l) cut[l++] = i + 1; }
          for (; l < lanes; l++) cut[l] = first + cnt; }
        c->epoch++;
        int rcs[kMaxLanes] = { 0, 0, 0, 0 };
        std::vector<std::thread> th;
        for (int l = 1; l < lanes; l++)
            th.emplace_back([&, l]() { cudaSetDevice(c->dev); rcs[l] = fzh_decode_run(c, l, (uint32_t)cut[l], (uint32_t)(cut[l + 1] - cut[l]), flags, true); });
        rcs[0] = fzh_decode_run(c, 0, (uint32_t)cut[0], (uint32_t)(cut[1] - cut[0]), flags, true);
        for (auto& t : th) t.join();
        for (int l = 0; l < lanes; l++) { if (rcs[l]) return rcs[l]; add_timing(c->lane[l].timing, l > 0); }
        return 0;
    };

    if (src_dev && dst_dev) {
        if ((rc = run(0, n, nullptr))) return rc;
    } else {
        // chunk boundaries: a small first chunk starts the device -> host stream early.  A chunk also wants enough ITEMS: a frame
        // is a serial chain (~1 GB/s at 32 warps per frame), so a chunk of one or two large files would run at that speed while
        // 64 of them side by side cost the same time -- large files make large chunks (up to 16 x the byte target).
        std::vector<size_t> cuts{ 0 };
        size_t target = chunk_bytes() / 4;
        constexpr size_t kMinItems = 64;
        for (size_t i = 0, bytes = 0, items = 0; i < n; i++) {
            bytes += dst_cap[i] + src_len[i]; items++;
            // (few large files: 16 x 1 GiB host-resident files decode in 0.94 s as ONE chunk -- a frame is a ~1.2 GB/s chain, sixteen run side by
            //  side -- against 0.86 s for each chunk of eight; so the byte cap is what HBM staging allows, 64 chunk targets = 32 GiB, not less)
            if (((bytes >= target && (items >= kMinItems || encode)) || bytes >= 64 * chunk_bytes()) && i + 1 < n) { cuts.push_back(i + 1); bytes = 0; items = 0; target = chunk_bytes(); }
        }
        cuts.push_back(n);
        const size_t nchunks = cuts.size() - 1;
        cudaStream_t s_in = c->copy_stream, s_out = c->copy_stream2;
        auto h2d = [&](size_t k) -> int {         // chunk k: host -> device on the copy-in stream
            if (src_dev) return 0;
            const size_t lo = cuts[k], hi = cuts[k + 1];
            if (src_scatter) {
                for (size_t i = lo; i < hi; i++) if (src_len[i]) CKR(cudaMemcpyAsync((void*)items[i].src, src[i], src_len[i], cudaMemcpyHostToDevice, s_in));
                CKR(cudaEventRecord(c->ev[12 + (k & 1)], s_in));
                return 0;
            }
            const uint8_t* d_lo = items[lo].src; const uint8_t* d_hi = items[hi - 1].src + src_len[hi - 1];
            const size_t off = (size_t)(d_lo - (const uint8_t*)c->d_stage_src.p);
            if (d_hi > d_lo) {
                cudaError_t e = cudaMemcpyAsync((void*)d_lo, h2d_base + off, (size_t)(d_hi - d_lo), cudaMemcpyHostToDevice, s_in);
                if (e == cudaErrorInvalidValue && h2d_base == (const uint8_t*)src[0]) {
                    // the caller's items looked like one pinned span but are not one allocation: a copy per item (same device layout)
                    cudaGetLastError();
                    e = cudaSuccess;
                    for (size_t i = lo; i < hi && e == cudaSuccess; i++) if (src_len[i]) e = cudaMemcpyAsync((void*)items[i].src, src[i], src_len[i], cudaMemcpyHostToDevice, s_in);
                }
                CKR(e);
            }
            CKR(cudaEventRecord(c->ev[12 + (k & 1)], s_in));
            return 0;
        };
        static const bool trace = getenv("FZG_TRACE") != nullptr;
        struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
        auto now_ms = [&]() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6; };
        if ((rc = h2d(0))) return rc;
        for (size_t k = 0; k < nchunks; k++) {
            const size_t lo = cuts[k], hi = cuts[k + 1];
            if (k + 1 < nchunks && (rc = h2d(k + 1))) return rc;
            const double ta = now_ms();
            // few large frames: their bytes go home while the chains still run (FzStreamOut); otherwise after the chunk, below
            size_t out_bytes = 0;
            for (size_t i = lo; i < hi; i++) out_bytes += dst_cap[i];
            const bool arm = !dst_dev && !encode && (hi - lo < 64 || c->n_lanes == 1) && out_bytes >= stream_out_min_bytes();
            c->so.dst = arm ? dst : nullptr; c->so.copy = s_out; c->so.poll = c->poll_stream; c->so.done = false;
            rc = run(lo, hi - lo, src_dev ? nullptr : c->ev[12 + (k & 1)]);                    // returns once chunk k is decoded
            const bool streamed = c->so.done;
            c->so.dst = nullptr; c->so.done = false;
            if (rc) return rc;
            if (trace) fprintf(stderr, "fzgpu: chunk %zu/%zu items %zu: run %.2f -> %.2f ms (gpu %.2f ms)%s\n", k, nchunks, hi - lo, ta, now_ms(), c->timing.total_ms, streamed ? " [output streamed behind the execute stage]" : "");
            if (!dst_dev && !streamed) {
                bool all_full = dst_contig;
                for (size_t i = lo; i < hi && all_full; i++) all_full = !outs[i].status && outs[i].dst_len == dst_cap[i];
                if (all_full) {
                    CKR(cudaMemcpyAsync(dst[lo], items[lo].dst, (size_t)((uint8_t*)dst[hi - 1] + dst_cap[hi - 1] - (uint8_t*)dst[lo]), cudaMemcpyDeviceToHost, s_out));
                } else {
                    for (size_t i = lo; i < hi; i++)
                        if (!outs[i].status && outs[i].dst_len) CKR(cudaMemcpyAsync(dst[i], items[i].dst, outs[i].dst_len, cudaMemcpyDeviceToHost, s_out));
                }
            }
        }
        if (!dst_dev) CKR(cudaStreamSynchronize(s_out));
        if (trace) fprintf(stderr, "fzgpu: all chunks back on the host at %.2f ms\n", now_ms());
    }
    uint64_t bytes_in = 0, bytes_out = 0;
    for (size_t i = 0; i < n; i++) {
        dst_len[i] = (size_t)outs[i].dst_len; status[i] = outs[i].status;
        bytes_in += src_len[i]; bytes_out += outs[i].dst_len;
    }
    c->timing = acc;
    c->timing.bytes_in = bytes_in; c->timing.bytes_out = bytes_out;
    return 0;
}

// Grows the staging of `device` for host-resident batches now (HBM for the compressed and the plain bytes, pinned control blocks for
// `items` files).  Allocating and freeing while batches are in flight stalls them: a daemon calls this once, before it mounts.
extern "C" int fzg_reserve_staging(int device, size_t items, size_t src_bytes, size_t dst_bytes)
{
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for(device);
    if (!c) return -ENODEV;
    std::lock_guard<std::mutex> lk(c->mu);
    CKR(cudaSetDevice(c->dev));
    if ((rc = c->h_items.reserve(items * sizeof(Item)))) return rc;
    if ((rc = c->h_outs.reserve(items * sizeof(ItemOut)))) return rc;
    if ((rc = c->d_stage_src.reserve(src_bytes + 64))) return rc;
    if ((rc = c->d_stage_dst.reserve(dst_bytes + 64))) return rc;
    return 0;
}

extern "C" int fzg_decode_batch(int device, size_t n, const void* const* src, const size_t* src_len, void* const* dst,
                                const size_t* dst_cap, size_t* dst_len, int* status, int flags)
{
    try { return run_batch(false, device, n, src, src_len, dst, dst_cap, dst_len, status, flags, 0, 0); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

extern "C" int fzg_encode_batch(int device, size_t n, const void* const* src, const size_t* src_len, void* const* dst,
                                const size_t* dst_cap, size_t* dst_len, int* status, int level, size_t chunk_size, int flags)
{
    try { return run_batch(true, device, n, src, src_len, dst, dst_cap, dst_len, status, flags, level, chunk_size); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

extern "C" size_t fzg_encode_bound(size_t src_len, size_t chunk_size) { return fzh_encode_bound(src_len, chunk_size); }

// ---------------------------------------------------------------------------------- host header walk
// Walks frame, skippable-frame and block headers on the host.  content = sum of Frame_Content_Size (UINT64_MAX when a
// frame does not declare it), bound = the most the blocks can regenerate: Block_Size for Raw / RLE blocks, Block_Maximum_Size
// for a Compressed one (RFC 8878 3.1.1.2.4) -- what a caller may allocate whatever the (untrusted) FCS field says.
static int walk_headers(const uint8_t* src, size_t len, uint64_t* content, uint64_t* compressed, uint64_t* bound_out)
{
    uint64_t ip = 0, total = 0, bound = 0; bool unknown = false, over = false;
    auto add = [&](uint64_t& acc, uint64_t v) { if (acc > UINT64_MAX - v) over = true; else acc += v; };
    while (ip < len) {
        if (len - ip < 4) return FZG_E_TRUNCATED;
        uint32_t magic = rd32u(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (len - ip < 8) return FZG_E_TRUNCATED;
            uint64_t sz = rd32u(src + ip + 4);
            if (len - ip - 8 < sz) return FZG_E_TRUNCATED;
            ip += 8 + sz; continue;
        }
        if (magic != kMagic) return FZG_E_MAGIC;
        FrameHdr h; int st = parse_frame_header(src + ip, len - ip, h);
        if (st) return st;
        ip += h.hsize;
        if (h.has_fcs) add(total, h.fcs); else unknown = true;
        const uint64_t block_max = h.window < kBlockMax ? h.window : kBlockMax;
        for (;;) {
            if (len - ip < 3) return FZG_E_TRUNCATED;
            uint32_t bh = rd24(src + ip); ip += 3;
            uint32_t type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) return FZG_E_CORRUPT;
            uint64_t adv = type == BT_RLE ? 1 : bsize;
            if (len - ip < adv) return FZG_E_TRUNCATED;
            add(bound, type == BT_COMPRESSED ? block_max : (uint64_t)bsize);
            ip += adv;
            if (bh & 1) break;
        }
        if (h.checksum) { if (len - ip < 4) return FZG_E_TRUNCATED; ip += 4; }
    }
    if (over) return FZG_E_UNSUPPORTED;                    // sizes that do not fit 64 bits: nothing real
    if (content) *content = unknown ? UINT64_MAX : total;
    if (compressed) *compressed = ip;
    if (bound_out) *bound_out = bound;
    return FZG_OK;
}

extern "C" int fzg_frame_info(const void* src_, size_t len, uint64_t* content_size, uint64_t* compressed_size)
{
    return walk_headers((const uint8_t*)src_, len, content_size, compressed_size, nullptr);
}

// ---------------------------------------------------------------------------------- fd entry points
static int read_all(int fd, std::vector<uint8_t>& buf, uint64_t want /*0 = to EOF*/)
{
    size_t got = 0;
    if (want) buf.resize(want); else buf.resize(1 << 20);
    for (;;) {
        if (!want && got == buf.size()) buf.resize(buf.size() * 2);
        size_t room = buf.size() - got;
        if (want && room == 0) break;
        ssize_t r = read(fd, buf.data() + got, room);
        if (r < 0) { if (errno == EINTR) continue; return -errno; }
        if (r == 0) break;
        got += (size_t)r;
    }
    buf.resize(got);
    return 0;
}
static int write_all(int fd, const uint8_t* p, size_t n)
{
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return -errno; }
        p += w; n -= (size_t)w;
    }
    return 0;
}

// The reference streams through constant memory (copy_decode, src/main.rs:463-467) and fails ONE open with EFAULT on a bad
// file (:467); this entry point buffers whole files, so nothing in the (untrusted) input may size an allocation beyond what
// its block headers can regenerate, and no exception may cross the C boundary (the fzfs daemon is one thread).
static int decode_fd_impl(int src_fd, int dst_fd, uint64_t shard_key, uint64_t* out_size)
{
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for_key(shard_key);
    if (!c) return -ENODEV;
    std::vector<uint8_t> in;
    if ((rc = read_all(src_fd, in, 0))) return rc;
    if (out_size) *out_size = 0;
    if (in.empty()) return 0;                      // empty input: success, empty output
    uint64_t content = 0, csize = 0, bound = 0;
    int st = walk_headers(in.data(), in.size(), &content, &csize, &bound);
    if (st) return st;
    if (content != UINT64_MAX && content > bound) return FZG_E_FCS;        // an FCS its blocks cannot reach: libzstd fails the frame too
    const uint64_t cap = content == UINT64_MAX ? bound : content;         // unknown size: the block walk bounds it, one attempt
    static const uint64_t cap_max = [] { const char* e = getenv("FZG_MAX_DECODE_BYTES"); return e ? strtoull(e, nullptr, 0) : (uint64_t)1 << 40; }();
    if (cap > cap_max || cap > (uint64_t)SIZE_MAX / 2) return -ENOMEM;
    uint8_t* out = (uint8_t*)malloc(cap ? cap : 1);
    if (!out) return -ENOMEM;
    const void* sp = in.data(); size_t sl = in.size(); void* dp = out; size_t dc = cap, dl = 0; int ist = 0;
    rc = fzg_decode_batch(c->dev, 1, &sp, &sl, &dp, &dc, &dl, &ist, 0);
    if (!rc && ist) rc = ist;
    if (!rc) rc = write_all(dst_fd, out, dl);
    if (!rc && out_size) *out_size = dl;
    free(out);
    return rc;
}
extern "C" int fzg_decode_fd(int src_fd, int dst_fd, uint64_t shard_key, uint64_t* out_size)
{
    try { return decode_fd_impl(src_fd, dst_fd, shard_key, out_size); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

static int encode_fd_impl(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t shard_key, uint64_t* out_size)
{
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for_key(shard_key);
    if (!c) return -ENODEV;
    std::vector<uint8_t> in;
    if ((rc = read_all(src_fd, in, src_size))) return rc;
    if (src_size && in.size() != src_size) return -EIO;
    size_t cap = fzg_encode_bound(in.size(), 0);
    std::vector<uint8_t> out(cap);
    const void* sp = in.data(); size_t sl = in.size(); void* dp = out.data(); size_t dl = 0; int ist = 0;
    static const bool seek = getenv("FZG_SEEK_TABLE") != nullptr;          // files written through the mount carry a seek table
    rc = fzg_encode_batch(c->dev, 1, &sp, &sl, &dp, &cap, &dl, &ist, level, 0, seek ? FZG_SEEK_TABLE : 0);
    if (rc) return rc;
    if (ist) return -EIO;
    if ((rc = write_all(dst_fd, out.data(), dl))) return rc;
    if (out_size) *out_size = dl;
    return 0;
}

extern "C" int fzg_encode_fd(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t shard_key, uint64_t* out_size)
{
    try { return encode_fd_impl(src_fd, dst_fd, level, src_size, shard_key, out_size); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

// ---------------------------------------------------------------------------------- seek table, partial reads
static inline uint32_t rd32le(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Footer of the zstd seekable format at the end of a file: Number_Of_Frames | descriptor | 0x8F92EAB1.  `tail` = the last
// `tail_len` (>= 9) bytes of the file.  On success: the number of frames and the size of the whole skippable frame.
extern "C" int fzg_seek_footer(const void* tail, size_t tail_len, uint64_t file_size, uint32_t* n_frames, uint64_t* table_bytes)
{
    if (!tail || !n_frames || !table_bytes) return -EINVAL;
    if (tail_len < 9 || file_size < 17) return -ENOENT;
    const uint8_t* f = (const uint8_t*)tail + tail_len - 9;
    if (rd32le(f + 5) != 0x8F92EAB1u) return -ENOENT;
    if (f[4] & 0x7C) return -ENOENT;                         // reserved descriptor bits
    const uint32_t nf = rd32le(f); const uint32_t entry = (f[4] & 0x80) ? 12 : 8;
    const uint64_t bytes = 8 + (uint64_t)entry * nf + 9;
    if (bytes > file_size) return -ENOENT;
    *n_frames = nf; *table_bytes = bytes;
    return 0;
}

// Decodes plain bytes [offset, offset + size) of a .zst file that carries a seek table, reading and decoding only the
// frames the range touches.  rd(buf, len, file_offset) reads the compressed file.  -ENOENT: no (valid) seek table.
template <class READ>
static int decode_range(FzCtx* c, READ rd, uint64_t file_size, uint64_t offset, size_t size, void* dst, size_t* got)
{
    *got = 0;
    uint8_t foot[9];
    if (file_size < 17) return -ENOENT;
    int rc = rd(foot, 9, file_size - 9); if (rc) return rc;
    uint32_t nf = 0; uint64_t tb = 0;
    if ((rc = fzg_seek_footer(foot, 9, file_size, &nf, &tb))) return rc;
    std::vector<uint8_t> tab(tb);
    if ((rc = rd(tab.data(), tb, file_size - tb))) return rc;
    if ((rd32le(tab.data()) & 0xFFFFFFF0u) != 0x184D2A50u || rd32le(tab.data() + 4) != tb - 8) return -ENOENT;
    const uint32_t entry = (foot[4] & 0x80) ? 12 : 8;
    // frames overlapping the range
    uint64_t coff = 0, doff = 0, c_lo = 0, d_lo = 0; uint32_t f0 = nf, f1 = nf;
    std::vector<uint64_t> cs, ds;
    for (uint32_t f = 0; f < nf; f++) {
        const uint64_t csz = rd32le(tab.data() + 8 + (size_t)entry * f), dsz = rd32le(tab.data() + 12 + (size_t)entry * f);
        if (f0 == nf && doff + dsz > offset && size) { f0 = f; c_lo = coff; d_lo = doff; }
        if (f0 != nf && f1 == nf) { cs.push_back(csz); ds.push_back(dsz); if (doff + dsz >= offset + size) f1 = f + 1; }
        coff += csz; doff += dsz;
    }
    if (coff + tb != file_size) return -ENOENT;              // the table does not describe this file
    if (f0 == nf) return 0;                                  // range at or past the end: nothing to read
    if (f1 == nf) f1 = nf;
    const size_t k = cs.size();
    uint64_t c_bytes = 0, d_bytes = 0;
    for (size_t i = 0; i < k; i++) { c_bytes += cs[i]; d_bytes += ds[i]; }
    std::vector<uint8_t> comp(c_bytes ? c_bytes : 1), plain(d_bytes ? d_bytes : 1);
    if ((rc = rd(comp.data(), c_bytes, c_lo))) return rc;
    std::vector<const void*> sp(k); std::vector<void*> dp(k); std::vector<size_t> sl(k), dc(k), dl(k); std::vector<int> st(k);
    { uint64_t a = 0, b = 0; for (size_t i = 0; i < k; i++) { sp[i] = comp.data() + a; sl[i] = cs[i]; dp[i] = plain.data() + b; dc[i] = ds[i]; a += cs[i]; b += ds[i]; } }
    rc = fzg_decode_batch(c->dev, k, sp.data(), sl.data(), dp.data(), dc.data(), dl.data(), st.data(), 0);
    if (rc) return rc;
    for (size_t i = 0; i < k; i++) { if (st[i]) return st[i]; if (dl[i] != ds[i]) return FZG_E_CORRUPT; }
    const uint64_t skip = offset - d_lo, avail = d_bytes - skip;
    const size_t n = (size_t)std::min<uint64_t>(size, avail);
    memcpy(dst, plain.data() + skip, n);
    *got = n;
    return 0;
}

extern "C" int fzg_decode_range(int device, const void* src, size_t len, uint64_t offset, size_t size, void* dst, size_t* got)
{
    if (!src || !dst || !got) return -EINVAL;
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for(device);
    if (!c) return -ENODEV;
    auto rd = [&](uint8_t* buf, size_t n, uint64_t at) -> int { if (at + n > len) return -EIO; memcpy(buf, (const uint8_t*)src + at, n); return 0; };
    try { return decode_range(c, rd, len, offset, size, dst, got); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

extern "C" int fzg_decode_range_fd(int src_fd, uint64_t shard_key, uint64_t offset, size_t size, void* dst, size_t* got)
{
    if (!dst || !got) return -EINVAL;
    int rc = ensure_init(); if (rc) return rc;
    FzCtx* c = ctx_for_key(shard_key);
    if (!c) return -ENODEV;
    const off_t end = lseek(src_fd, 0, SEEK_END);
    if (end < 0) return -errno;
    auto rd = [&](uint8_t* buf, size_t n, uint64_t at) -> int {
        size_t done = 0;
        while (done < n) { const ssize_t r = pread(src_fd, buf + done, n - done, (off_t)(at + done)); if (r < 0) { if (errno == EINTR) continue; return -errno; } if (r == 0) return -EIO; done += (size_t)r; }
        return 0;
    };
    try { return decode_range(c, rd, (uint64_t)end, offset, size, dst, got); }
    catch (const std::bad_alloc&) { return -ENOMEM; }
    catch (...) { return -EIO; }
}

// ---------------------------------------------------------------------------------- misc
extern "C" const char* fzg_strerror(int code)
{
    switch (code) {
    case FZG_OK: return "ok";
    case FZG_E_MAGIC: return "unknown frame descriptor";
    case FZG_E_TRUNCATED: return "input ends inside a frame";
    case FZG_E_UNSUPPORTED: return "unsupported frame parameter (reserved bit, dictionary, window > 2^27)";
    case FZG_E_CORRUPT: return "corrupted block";
    case FZG_E_DSTSIZE: return "destination buffer too small";
    case FZG_E_CHECKSUM: return "content checksum mismatch";
    case FZG_E_FCS: return "decoded size differs from Frame_Content_Size";
    default: return code < 0 ? strerror(-code) : "unknown status";
    }
}

extern "C" int fzg_last_timing(int device, fzg_timing_t* out)
{
    FzCtx* c = ctx_for(device);
    if (!c || !out) return -EINVAL;
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->timing;
    return 0;
}

extern "C" uint64_t fzg_streamed_copies(int device)
{
    FzCtx* c = ctx_for(device);
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->so.pieces;
}

extern "C" const char* fzg_stage_name(int stage) { return stage < 16 ? fzh_decode_stage_name(stage) : fzh_encode_stage_name(stage - 16); }

extern "C" void* fzg_stream(int device)
{
    FzCtx* c = ctx_for(device);
    return c ? (void*)c->stream : nullptr;
}
