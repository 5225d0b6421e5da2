/*
 * fz_decode.cu -- sm_100a kernels and launch sequence of the batched zstd decoder.
 *
 * Replaces zstd::stream::copy_decode (/root/reference/src/main.rs:463-467) for a whole batch
 * of .zst files per call.  Stages (one kernel each, all on the context's stream):
 *
 *   count      thread/item   frame + block + section headers walked on device, sizes counted
 *   scan       1 CTA         exclusive prefix sums -> descriptor / scratch bases, totals
 *   fill       thread/item   Frame / Block descriptors, Repeat/Treeless provenance, job lists
 *   literals   4 thr/block   Huffman tree -> shared-memory table; 1 or 4 interleaved streams
 *   sequences  thread/block  LL/OF/ML FSE tables in shared memory; backward bitstream -> records
 *   offsets    thread/item   output placement, FCS / capacity checks, first-error folding
 *   execute    warp/frame    LZ77: literal + match copies in stream order, lanes split each copy
 *   checksum   4 thr/frame   XXH64 (four independent accumulators) vs the stored trailer
 *   finish     thread/item   per-item status + size
 */
#include <cuda_runtime.h>
#include <stdio.h>

#include "fz_host.h"
#include "fz_kernels.cuh"

namespace fz {

__constant__ SeqConsts c_seq_consts = { FZ_LL_BASE, FZ_ML_BASE, FZ_LL_BITS, FZ_ML_BITS, FZ_LL_DEF, FZ_OF_DEF, FZ_ML_DEF };

// ------------------------------------------------------------------ count / scan / fill
__global__ void k_count(const Item* items, ItemInfo* infos, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemInfo info;
    walk_item<false>(i, items[i], info, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    infos[i] = info;
}

// single CTA; exclusive scan of six counters over the items.  totals[0..5] = frames, blocks,
// seq jobs, huf jobs, literal bytes, sequences.
__global__ void k_scan(const ItemInfo* infos, ItemBase* bases, uint64_t* totals, uint32_t n)
{
    __shared__ uint64_t part[1024][6];
    const uint32_t t = threadIdx.x, T = blockDim.x;
    const uint32_t per = (n + T - 1) / T;
    const uint32_t lo = t * per < n ? t * per : n, hi = lo + per < n ? lo + per : n;
    uint64_t acc[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t i = lo; i < hi; i++) {
        const ItemInfo& f = infos[i];
        acc[0] += f.n_frames; acc[1] += f.n_blocks; acc[2] += f.n_seq_jobs; acc[3] += f.n_huf_jobs;
        acc[4] += f.lit_bytes; acc[5] += f.n_seq;
    }
    for (int c = 0; c < 6; c++) part[t][c] = acc[c];
    __syncthreads();
    if (t == 0) {
        uint64_t run[6] = { 0, 0, 0, 0, 0, 0 };
        for (uint32_t k = 0; k < T; k++)
            for (int c = 0; c < 6; c++) { uint64_t v = part[k][c]; part[k][c] = run[c]; run[c] += v; }
        for (int c = 0; c < 6; c++) totals[c] = run[c];
    }
    __syncthreads();
    for (int c = 0; c < 6; c++) acc[c] = part[t][c];
    for (uint32_t i = lo; i < hi; i++) {
        const ItemInfo& f = infos[i];
        ItemBase b;
        b.frame = (uint32_t)acc[0]; b.block = (uint32_t)acc[1]; b.seq_job = (uint32_t)acc[2]; b.huf_job = (uint32_t)acc[3];
        b.lit = acc[4]; b.seq = acc[5];
        bases[i] = b;
        acc[0] += f.n_frames; acc[1] += f.n_blocks; acc[2] += f.n_seq_jobs; acc[3] += f.n_huf_jobs;
        acc[4] += f.lit_bytes; acc[5] += f.n_seq;
    }
}

__global__ void k_fill(const Item* items, ItemInfo* infos, const ItemBase* bases, Frame* frames, Block* blocks,
                       uint32_t* seq_jobs, uint32_t* huf_jobs, uint8_t* lit_scratch, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemInfo info;
    walk_item<true>(i, items[i], info, &bases[i], frames, blocks, seq_jobs, huf_jobs, lit_scratch);
}

// ------------------------------------------------------------------ literals
// Four threads per block: thread 0 of the group parses the tree description and fills the
// group's table in shared memory, then each thread decodes one of the (up to) four streams.
constexpr int kLitGroups = 16;                  // blocks per CTA
constexpr int kLitThreads = kLitGroups * 4;
constexpr int kHufTableCells = 1 << kHufLogMax; // 8 KB per group

__global__ void __launch_bounds__(kLitThreads) k_literals(Block* blocks, const uint32_t* jobs, uint32_t n_jobs)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* table = (uint16_t*)smem + (threadIdx.x >> 2) * kHufTableCells;
    __shared__ int s_log[kLitGroups];
    __shared__ uint32_t s_used[kLitGroups];
    const uint32_t grp = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const uint32_t job = blockIdx.x * kLitGroups + grp;
    const bool active = job < n_jobs;
    Block* b = active ? &blocks[jobs[job]] : nullptr;
    const bool huf = active && b->lit_type >= LT_HUF;
    if (huf && sub == 0) { int log; uint32_t used; lit_build(blocks, *b, table, log, used); s_log[grp] = log; s_used[grp] = used; }
    __syncwarp();
    if (active && lit_decode_sub(*b, sub, table, huf ? s_log[grp] : 0, huf ? s_used[grp] : 0)) b->status = FZG_E_CORRUPT;
}

// ------------------------------------------------------------------ sequences
constexpr int kSeqLanes = 22;                    // blocks per CTA (one thread each); 2 CTAs per SM
constexpr int kSeqTableCells = 512 + 256 + 512;  // LL, OF, ML cells (4 B each) per thread

__global__ void __launch_bounds__(32) k_sequences(Block* blocks, const Frame* frames, const uint32_t* jobs, uint32_t n_jobs,
                                                  uint64_t* seqs)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ SeqConsts K;
    for (uint32_t i = threadIdx.x; i < sizeof(SeqConsts) / 4; i += blockDim.x) ((uint32_t*)&K)[i] = ((const uint32_t*)&c_seq_consts)[i];
    __syncthreads();
    const uint32_t job = blockIdx.x * kSeqLanes + threadIdx.x;
    if (threadIdx.x >= kSeqLanes || job >= n_jobs) return;
    uint16_t cnt[64];
    seq_thread(blocks, frames, blocks[jobs[job]], K, (uint32_t*)smem + threadIdx.x * kSeqTableCells, cnt, seqs);
}

// blocks without sequences regenerate exactly their literals
__global__ void k_rsize_nseq0(Block* blocks, uint32_t n_blocks)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_blocks && blocks[i].type == BT_COMPRESSED && blocks[i].nseq == 0) blocks[i].rsize = blocks[i].lit_regen;
}

// ------------------------------------------------------------------ offsets / execute / checksum / finish
__global__ void k_offsets(const Item* items, const ItemInfo* infos, const ItemBase* bases, Frame* frames, Block* blocks,
                          ItemOut* outs, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemOut o;
    offsets_item(items[i], infos[i], bases[i], frames, blocks, o);
    outs[i] = o;
}

struct GpuWarp {
    static constexpr int kLanes = 32;
    __device__ __forceinline__ uint32_t lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ uint64_t shfl64(uint64_t v, uint32_t src) const { return __shfl_sync(0xFFFFFFFFu, v, src); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

constexpr int kExecWarps = 4;
__global__ void __launch_bounds__(kExecWarps * 32) k_execute(Frame* frames, const Block* blocks, const Item* items,
                                                              const ItemOut* outs, const uint64_t* seqs, uint32_t n_frames)
{
    const uint32_t f = blockIdx.x * kExecWarps + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    Frame& fr = frames[f];
    if (outs[fr.item].fail) return;
    exec_frame(GpuWarp(), fr, blocks, items[fr.item], seqs);
}

// Four threads per frame, one XXH64 accumulator each (stripe = 32 bytes, lane j owns bytes 8j..8j+7).
__global__ void k_checksum(Frame* frames, const Item* items, const ItemOut* outs, uint32_t n_frames)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t f = t >> 2, j = t & 3;
    const bool active = f < n_frames && frames[f].has_checksum && !frames[f].status && !outs[frames[f].item].fail;
    uint64_t acc = 0; const uint8_t* p = nullptr; uint64_t len = 0;
    if (active) { p = items[frames[f].item].dst + frames[f].out_off; len = frames[f].out_size; acc = xx_lane(p, len, j); }
    const uint32_t base_lane = (threadIdx.x & 31) & ~3u;   // gather the group's four accumulators on its first thread
    const uint64_t v1 = __shfl_sync(0xFFFFFFFFu, acc, base_lane), v2 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 1),
                   v3 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 2), v4 = __shfl_sync(0xFFFFFFFFu, acc, base_lane + 3);
    if (active && j == 0 && (uint32_t)xx_combine(v1, v2, v3, v4, p, len) != frames[f].checksum) frames[f].status = FZG_E_CHECKSUM;
}

__global__ void k_finish(const ItemInfo* infos, const ItemBase* bases, const Frame* frames, ItemOut* outs, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemOut o = outs[i];
    finish_item(infos[i], bases[i], frames, o);
    outs[i] = o;
}

}  // namespace fz

// ====================================================================== host side
using namespace fz;

static const char* kStageNames[] = { "count", "scan", "fill", "literals", "sequences", "offsets", "execute", "checksum",
                                     "finish" };
const char* fzh_decode_stage_name(int s) { return s >= 0 && s < 9 ? kStageNames[s] : ""; }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "fzgpu: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return -5 /*-EIO*/; } } while (0)

int fzh_decode_setup(void)
{
    CK(cudaFuncSetAttribute(k_literals, cudaFuncAttributeMaxDynamicSharedMemorySize, kLitGroups * kHufTableCells * 2));
    CK(cudaFuncSetAttribute(k_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqLanes * kSeqTableCells * 4));
    return 0;
}

// Runs the whole pipeline for n items whose Item records (device pointers) are in c->h_items.
// Results land in c->h_outs (pinned).  Blocking.
int fzh_decode_run(FzCtx* c, uint32_t n, int flags)
{
    cudaStream_t s = c->stream;
    const bool prof = flags & FZG_PROFILE;
    c->timing = fzg_timing_t{};
    if (n == 0) return 0;
    int rc;
    if ((rc = c->d_items.reserve(n * sizeof(Item)))) return rc;
    if ((rc = c->d_infos.reserve(n * sizeof(ItemInfo)))) return rc;
    if ((rc = c->d_bases.reserve(n * sizeof(ItemBase)))) return rc;
    if ((rc = c->d_outs.reserve(n * sizeof(ItemOut)))) return rc;
    if ((rc = c->d_totals.reserve(64))) return rc;
    Item* d_items = (Item*)c->d_items.p; ItemInfo* d_infos = (ItemInfo*)c->d_infos.p; ItemBase* d_bases = (ItemBase*)c->d_bases.p;
    ItemOut* d_outs = (ItemOut*)c->d_outs.p; uint64_t* d_totals = (uint64_t*)c->d_totals.p;

    CK(cudaMemcpyAsync(d_items, c->h_items.p, n * sizeof(Item), cudaMemcpyHostToDevice, s));
    int ev = 0;
    auto mark = [&]() { if (prof || ev == 0) cudaEventRecord(c->ev[ev], s); ev++; };
    mark();                                                         // ev0: start
    const uint32_t tb = 128, gi = (n + tb - 1) / tb;
    k_count<<<gi, tb, 0, s>>>(d_items, d_infos, n); mark();
    k_scan<<<1, 1024, 0, s>>>(d_infos, d_bases, d_totals, n); mark();
    CK(cudaMemcpyAsync(c->h_totals.p, d_totals, 48, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t* tot = (const uint64_t*)c->h_totals.p;
    const uint64_t n_frames = tot[0], n_blocks = tot[1], n_sj = tot[2], n_hj = tot[3], lit_bytes = tot[4], n_seq = tot[5];
    if (n_blocks >= (1ull << 31) || n_frames >= (1ull << 31)) return -22;
    if ((rc = c->d_frames.reserve((n_frames + 1) * sizeof(Frame)))) return rc;
    if ((rc = c->d_blocks.reserve((n_blocks + 1) * sizeof(Block)))) return rc;
    if ((rc = c->d_seq_jobs.reserve((n_sj + 1) * 4))) return rc;
    if ((rc = c->d_huf_jobs.reserve((n_hj + 1) * 4))) return rc;
    if ((rc = c->d_lit.reserve(lit_bytes + 64))) return rc;
    if ((rc = c->d_seq.reserve((n_seq + 8) * 8))) return rc;
    Frame* d_frames = (Frame*)c->d_frames.p; Block* d_blocks = (Block*)c->d_blocks.p;
    uint32_t* d_sj = (uint32_t*)c->d_seq_jobs.p; uint32_t* d_hj = (uint32_t*)c->d_huf_jobs.p;

    k_fill<<<gi, tb, 0, s>>>(d_items, d_infos, d_bases, d_frames, d_blocks, d_sj, d_hj, (uint8_t*)c->d_lit.p, n); mark();
    int launches = 3;
    if (n_hj) { k_literals<<<(uint32_t)((n_hj + kLitGroups - 1) / kLitGroups), kLitThreads, kLitGroups * kHufTableCells * 2, s>>>(d_blocks, d_hj, (uint32_t)n_hj); launches++; }
    mark();
    if (n_blocks) { k_rsize_nseq0<<<(uint32_t)((n_blocks + 255) / 256), 256, 0, s>>>(d_blocks, (uint32_t)n_blocks); launches++; }
    if (n_sj) { k_sequences<<<(uint32_t)((n_sj + kSeqLanes - 1) / kSeqLanes), 32, kSeqLanes * kSeqTableCells * 4, s>>>(d_blocks, d_frames, d_sj, (uint32_t)n_sj, (uint64_t*)c->d_seq.p); launches++; }
    mark();
    k_offsets<<<gi, tb, 0, s>>>(d_items, d_infos, d_bases, d_frames, d_blocks, d_outs, n); mark(); launches++;
    if (n_frames) { k_execute<<<(uint32_t)((n_frames + kExecWarps - 1) / kExecWarps), kExecWarps * 32, 0, s>>>(d_frames, d_blocks, d_items, d_outs, (const uint64_t*)c->d_seq.p, (uint32_t)n_frames); launches++; }
    mark();
    if (n_frames && !(flags & FZG_NO_VERIFY_CHECKSUM)) { k_checksum<<<(uint32_t)((n_frames * 4 + 127) / 128), 128, 0, s>>>(d_frames, d_items, d_outs, (uint32_t)n_frames); launches++; }
    mark();
    k_finish<<<gi, tb, 0, s>>>(d_infos, d_bases, d_frames, d_outs, n); launches++;
    if (!prof) ev = 9;
    cudaEventRecord(c->ev[ev], s);                                   // last event
    CK(cudaMemcpyAsync(c->h_outs.p, d_outs, n * sizeof(ItemOut), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    c->timing.launches = launches;
    cudaEventElapsedTime(&c->timing.total_ms, c->ev[0], c->ev[ev]);
    if (prof) for (int k = 0; k < 9; k++) cudaEventElapsedTime(&c->timing.kernel_ms[k], c->ev[k], c->ev[k + 1]);
    return 0;
}
