/*
 * fz_host.h -- per-device context of libfzgpu.so: stream, events, grow-only HBM scratch and
 * pinned host staging.  One FzCtx per GPU; a call holds the context mutex for its duration, so
 * several host threads (one per GPU, or several per GPU) may call the C ABI concurrently.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <mutex>

#include "../../include/fzgpu.h"

struct FzDevBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n)   // grow-only; contents are NOT preserved
    {
        if (n <= cap) return 0;
        size_t want = n + (n / 4 < ((size_t)1 << 30) ? n / 4 : ((size_t)1 << 30)) + 4096;     // grow-only with some slack, at most 1 GiB of it
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError(); p = nullptr; want = n + 4096;                               // a batch that nearly fills HBM: no slack
            if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -12; /* -ENOMEM */ }
        }
        cap = want; return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct FzPinBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        size_t want = n + n / 4 + 4096;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -12; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One decode pipeline in flight: its own stream, events and scratch.  A context runs several lanes side by side
// (each on a slice of the batch, driven by its own host thread) so that the shared-memory-bound entropy stages of
// one slice overlap the issue-bound LZ77 stage of another.
struct FzLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[16] = {};
    FzDevBuf d_infos, d_bases, d_outs, d_totals, d_frames, d_blocks, d_seq_jobs, d_huf_jobs, d_lit, d_seq, d_seq_tabs, d_seq_hdrs, d_prog;
    FzPinBuf h_totals, h_prog;
    fzg_timing_t timing = {};
    cudaEvent_t ev_prog = nullptr;         // recorded once the frames' progress words are zeroed (see FzStreamOut)
    cudaEvent_t ev_entropy = nullptr;      // recorded after this lane's entropy stages (literals, sequences, records)
    cudaStream_t side = nullptr;           // small batches: the literal stage runs here, beside the sequence stages
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::atomic<uint64_t> entropy_epoch{ 0 };   // call number whose ev_entropy has been enqueued
};
constexpr int kMaxLanes = 4;

// Few large frames into host buffers: a frame is a serial chain (~1.2 GB/s) that publishes how far it is flushed, so the
// device -> host copy of a chunk need not wait for the chunk: fzh_decode_run polls the progress words while the execute stage
// runs and sends the finished part of every frame down the copy-out stream, piece by piece.  run_batch arms it per chunk.
struct FzStreamOut {
    void* const* dst = nullptr;            // host destinations of the call's items (indexed like h_items); null: off
    cudaStream_t copy = nullptr, poll = nullptr;
    bool done = false;                     // set by fzh_decode_run when it has queued every device -> host copy of the chunk itself
    uint64_t pieces = 0;                   // copies queued while the kernels were still running (statistics)
};

struct FzCtx {
    int dev = -1;
    cudaStream_t stream = nullptr;         // = lane[0].stream
    cudaStream_t copy_stream = nullptr;    // host -> device
    cudaStream_t copy_stream2 = nullptr;   // device -> host
    cudaEvent_t ev[16] = {};               // [0..11] encoder stages, [12..13] chunk pipeline
    std::mutex mu;
    FzLane lane[kMaxLanes];
    int n_lanes = 1;
    uint64_t epoch = 0;                    // call counter (staggering of the lanes, see fzh_decode_run)
    cudaStream_t poll_stream = nullptr;    // small reads of progress words beside the running kernels
    FzStreamOut so;
    // staging for host-resident batches
    FzDevBuf d_stage_src, d_stage_dst;
    FzPinBuf h_items, h_outs, h_totals, h_stage_src, h_stage_dst, e_chunks_h, e_first_h;
    // encoder scratch
    FzDevBuf e_items, e_outs, e_work, e_tab, d_outs, d_totals;
    fzg_timing_t timing = {};
};

// fz_decode.cu
int fzh_decode_setup(void);
int fzh_decode_run(FzCtx* c, int lane, uint32_t first, uint32_t n, int flags, bool staggered);   // items [first, first + n) of c->h_items on lane `lane`
const char* fzh_decode_stage_name(int s);
