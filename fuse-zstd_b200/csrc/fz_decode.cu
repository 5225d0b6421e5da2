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

#include <time.h>
#include <algorithm>
#include <thread>

#include "fz_host.h"
#include "fz_kernels.cuh"
#include "fz_exec_tile.cuh"

namespace fz {


// ------------------------------------------------------------------ count / scan / fill
__global__ void k_count(const Item* items, ItemInfo* infos, uint32_t n, uint32_t* tickets)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 8) tickets[i] = 0;
    if (i >= n) return;
    ItemInfo info;
    walk_item<false>(i, items[i], info, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    infos[i] = info;
}

// single CTA; exclusive scan of six counters over the items.  totals[0..5] = frames, blocks,
// seq jobs, huf jobs, literal bytes, sequence records.
__global__ void k_scan(const ItemInfo* infos, ItemBase* bases, uint64_t* totals, uint32_t n)   // totals: pinned host memory (zero-copy)
{
    __shared__ uint64_t part[512][6];
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
// Four threads per block: thread 0 of the group parses the tree description (direct or FSE-compressed
// weights, or the tree of an earlier block for Treeless literals) and fills the group's table in shared
// memory, then each thread decodes one of the (up to) four Huffman streams, pulling its bitstream
// through a cp.async ring.  24 groups (9 KB each) per CTA, one CTA per SM, blocks drawn from a ticket.
#ifndef FZ_LIT_WARP_GROUPS
#define FZ_LIT_WARP_GROUPS 2
#endif
// The Huffman chain is latency-bound (index -> LDS -> length -> shift per symbol), so the stage is as fast as the number
// of tables an SM holds, all of them busy:
// * tables of 2^11 cells (the deepest tree libzstd's encoder builds): 45 blocks in flight per SM; a deeper tree (the
//   format allows 2^12) is put on a list and decoded by a second launch with 24 tables of 2^12 cells per SM;
// * a warp carries just kLitWarpGroups blocks (4 lanes each, the other lanes idle), because the blocks of a warp run in
//   lockstep and finish together -- with eight blocks per warp the streams sat idle 60 % of the time waiting for the
//   longest one (ncu: 13 of 32 lanes active).  The idle lanes cost nothing: issue slots are plentiful.
constexpr int kLitWarpGroups = FZ_LIT_WARP_GROUPS;        // blocks per warp
template <int LOG> struct LitCfg {
    static constexpr int fit = (227 * 1024 - 1024) / (int)lit_group_bytes(LOG);
    static constexpr int groups = fit < 32 * kLitWarpGroups ? fit : 32 * kLitWarpGroups;   // blocks in flight per CTA (one CTA per SM, <= 1024 threads)
    static constexpr int warps = (groups + kLitWarpGroups - 1) / kLitWarpGroups;
    static constexpr int threads = warps * 32;
    static constexpr int smem = groups * (int)lit_group_bytes(LOG);
};

// n_jobs_dev != nullptr: the job count is read from device memory (the list of deferred blocks).  deferred != nullptr:
// blocks whose tree is deeper than LOG are appended there instead of being decoded.
template <int LOG>
__global__ void __launch_bounds__(LitCfg<LOG>::threads, 1) k_literals(Block* blocks, const uint32_t* jobs, uint32_t n_jobs, const uint32_t* n_jobs_dev,
                                                                       uint32_t* ticket, uint32_t* deferred, uint32_t* n_deferred)
{
    constexpr int kGroups = LitCfg<LOG>::groups;
    constexpr uint32_t kGroupBytes = lit_group_bytes(LOG);
    extern __shared__ __align__(256) uint8_t smem_lit[];     // the group size is a multiple of 256: the rings are 256-byte aligned
    uint8_t* const smem = smem_lit;
    __shared__ int s_log[kGroups];
    __shared__ uint32_t s_used[kGroups];
    if (n_jobs_dev) n_jobs = *n_jobs_dev;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = warp * kLitWarpGroups, mine = min((uint32_t)kLitWarpGroups, (uint32_t)kGroups - first);   // this warp's groups
    const uint32_t sub = lane & 3, lg = lane >> 2;
    const bool lane_on = lg < mine;
    const uint32_t grp = first + (lane_on ? lg : 0);
    uint8_t* gmem = smem + grp * kGroupBytes;
    uint16_t* table = (uint16_t*)gmem;
    uint8_t* rings = gmem + (1u << LOG) * 2;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, mine);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_jobs) return;
        const uint32_t job = base + lg;
        bool active = lane_on && job < n_jobs;
        Block* b = active ? &blocks[jobs[job]] : nullptr;
        const bool huf = active && b->lit_type >= LT_HUF;
        if (huf && sub == 0) {
            int log; uint32_t used; lit_build(blocks, *b, table, *(LitScratch*)rings, LOG, log, used); s_log[grp] = log; s_used[grp] = used;
            if (log == -2) { if (deferred) deferred[atomicAdd(n_deferred, 1u)] = jobs[job]; else s_log[grp] = -1; }
        }
        __syncwarp();
        if (huf && s_log[grp] == -2) active = false;           // decoded by the second launch
        LitWork wk{ nullptr, nullptr, 0, 0, 0 };
        if (active) wk = lit_plan(*b, sub, huf ? s_log[grp] : 0, huf ? s_used[grp] : 0);
        const uint32_t bound = __reduce_max_sync(0xFFFFFFFFu, wk.kind == 2 ? wk.n_out / 4 : 0u);
        const int bad = lit_run(wk, sub, table, huf && active ? s_log[grp] : 0, rings + sub * 256, bound, 0xFFFFFFFFu);
        if (active && bad) b->status = FZG_E_CORRUPT;
        __syncwarp();
    }
}

// ------------------------------------------------------------------ sequences, table stage
// One thread per block builds its three FSE tables straight into HBM (seq_tables_thread): every block of the batch is in
// flight at once, so the serial, data-dependent build code costs its own latency once instead of stalling the chain warps.
constexpr int kTabThreads = 32;
__global__ void __launch_bounds__(kTabThreads) k_seq_tables(Block* blocks, const uint32_t* jobs, uint32_t n_jobs, uint8_t* tabs, SeqJobHdr* hdrs)
{
    __shared__ SeqConsts K;
    __shared__ uint8_t s_sym[512 * kTabThreads];              // lane-wise work arrays (Lanewise): element i of thread t at [i * kTabThreads + t]
    __shared__ int16_t s_norm[64 * kTabThreads];
    __shared__ uint16_t s_cnt[64 * kTabThreads];
    for (uint32_t i = threadIdx.x; i < sizeof(SeqConsts) / 4; i += blockDim.x) ((uint32_t*)&K)[i] = ((const uint32_t*)&c_seq_consts)[i];
    __syncthreads();
    const uint32_t job = blockIdx.x * kTabThreads + threadIdx.x;
    if (job >= n_jobs) return;
    Block& b = blocks[jobs[job]];
    SeqJobHdr h;
    const TabWork w{ { s_sym + threadIdx.x, kTabThreads }, { s_norm + threadIdx.x, kTabThreads }, { s_cnt + threadIdx.x, kTabThreads } };
    seq_tables_thread(blocks, b, K, tabs + (size_t)job * kJobTableBytes, h, w);
    hdrs[job] = h;
    if (h.bad && !b.status) b.status = FZG_E_CORRUPT;
}

// ------------------------------------------------------------------ sequences, stage A: the FSE chain
// One thread per block (the FSE state chain is serial), one CTA per SM.  What bounds this stage is the
// latency of that chain times the number of chains an SM can hold, and the latter is set by shared
// memory: a stream needs 2816 bytes (16-bit chain cells + bitstream ring), so 82 streams fit in the
// 227 KB of an SM.  A warp carries only a few streams (data-dependent branches cost little
// that way, and the SM has issue slots to spare); each warp draws its next batch of blocks from a ticket.
#ifndef FZ_SEQ_STREAMS
#define FZ_SEQ_STREAMS 82
#endif
constexpr int kSeqStreams = FZ_SEQ_STREAMS;
#ifndef FZ_SEQ_LANES
#define FZ_SEQ_LANES 21
#endif
constexpr int kSeqLanes = FZ_SEQ_LANES;
constexpr int kSeqWarps = (kSeqStreams + kSeqLanes - 1) / kSeqLanes;
constexpr int kSeqThreads = kSeqWarps * 32;
constexpr int kSeqSmem = kSeqStreams * kChainBytes;

__global__ void __launch_bounds__(kSeqThreads, 1) k_sequences(Block* blocks, const uint32_t* jobs, uint32_t n_jobs, uint64_t* seqs,
                                                                uint32_t* ticket, const uint8_t* tabs, const SeqJobHdr* hdrs)
{
    extern __shared__ __align__(256) uint8_t smem_seq[];   // kChainBytes is a multiple of 256: every stream's ring is 256-byte aligned
    uint8_t* const smem = smem_seq;
    if (((uint32_t)__cvta_generic_to_shared(smem) & 255u) != 0) __trap();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = warp * kSeqLanes, lanes = min((uint32_t)kSeqLanes, kSeqStreams - first);
    uint8_t* mine = smem + (first + (lane < lanes ? lane : 0)) * kChainBytes;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(ticket, lanes);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n_jobs) return;
        const uint32_t job = base + lane;
        const bool on = lane < lanes && job < n_jobs;
        Block* b = on ? &blocks[jobs[job]] : nullptr;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, on);
        const uint32_t bound = __reduce_max_sync(0xFFFFFFFFu, on ? b->nseq - 1 : 0u);     // nseq >= 1 for a sequence job
        if (on) seq_chain_thread(*b, tabs + (size_t)job * kJobTableBytes, hdrs[job], mine, seqs, bound, mask);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ sequences, stage B: records
// One warp per block, 32 sequences per step, in place over the RAW records of stage A: extra bits ->
// (literal length, match length, offset value); warp scans -> cumulative positions E / LE; repeat
// offsets; the positional record (rec_pack).  Repeat offsets are the only serial
// part: a step without repeat codes takes its history from the last three lanes, otherwise the history
// hops from one repeat code to the next (a few per step), never lane by lane.
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= (uint32_t)d) v += t; }
    return v;
}

#ifndef FZ_REC_WARPS
#define FZ_REC_WARPS 8
#endif
constexpr int kRecWarps = FZ_REC_WARPS;
__global__ void __launch_bounds__(kRecWarps * 32) k_records(Block* blocks, const Frame* frames, const uint32_t* jobs, uint32_t n_jobs,
                                                            uint64_t* seqs, const uint8_t* tabs, const SeqJobHdr* hdrs)
{
    __shared__ SeqConsts K;
    __shared__ __align__(16) uint8_t s_y[kRecWarps][2][512]; // state -> symbol maps of the block (LL, ML), from the table stage
    constexpr uint32_t kFullMask = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < sizeof(SeqConsts) / 4; i += blockDim.x) ((uint32_t*)&K)[i] = ((const uint32_t*)&c_seq_consts)[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t job = blockIdx.x * kRecWarps + warp;
    if (job >= n_jobs) return;
    Block& b = blocks[jobs[job]];
    if (b.status) return;
    const uint8_t* yLL = s_y[warp][0]; const uint8_t* yML = s_y[warp][1];
    {
        const uint4* src = (const uint4*)(tabs + (size_t)job * kJobTableBytes + kChainCellBytes);
        ((uint4*)s_y[warp][0])[lane] = src[lane]; ((uint4*)s_y[warp][0])[lane + 32] = src[lane + 32];
    }
    __syncwarp();
    const uint8_t* const bits = b.src + hdrs[job].bits_off;
    const uint32_t nseq = b.nseq, lit_regen = b.lit_regen, block_max = frames[b.frame].block_max;
    uint64_t* sq = seqs + b.seq_base;                                      // 32-byte aligned (walk_item pads the record counts)
    uint32_t rep0 = off_sym(0), rep1 = off_sym(1), rep2 = off_sym(2);      // history, warp-uniform
    uint32_t Ebase = 0, LEbase = 0; bool bad = false;
    // 64 sequences per step, two per lane (positions 2 * lane and 2 * lane + 1): the per-step work of the warp -- scans, repeat-offset
    // history, carries -- is paid once per 64 sequences, and a lane moves its two records as one 16-byte access.
    uint4 qn = make_uint4(0, 0, 0, 0);                                     // the records of a step are loaded a step early
    if (2 * lane + 1 < nseq) qn = *(const uint4*)(sq + 2 * lane); else if (2 * lane < nseq) { const uint64_t t = sq[2 * lane]; qn.x = (uint32_t)t; qn.y = (uint32_t)(t >> 32); }
    for (uint32_t g = 0; g < nseq && !bad; g += 64) {
        const uint32_t i0 = g + 2 * lane;
        const bool v0 = i0 < nseq, v1 = i0 + 1 < nseq;
        const uint32_t nv = min(64u, nseq - g);
        const uint64_t r0 = (uint64_t)qn.x | ((uint64_t)qn.y << 32), r1 = (uint64_t)qn.z | ((uint64_t)qn.w << 32);
        qn = make_uint4(0, 0, 0, 0);
        if (i0 + 65 < nseq) qn = *(const uint4*)(sq + i0 + 64); else if (i0 + 64 < nseq) { const uint64_t t = sq[i0 + 64]; qn.x = (uint32_t)t; qn.y = (uint32_t)(t >> 32); }
        uint32_t ll0 = 0, ml0 = 0, ofv0 = 4, ll1 = 0, ml1 = 0, ofv1 = 4; bool ok = true;
        if (v0) ok = raw_unpack(r0, K, yLL, yML, bits, ll0, ml0, ofv0);
        if (v1) ok = raw_unpack(r1, K, yLL, yML, bits, ll1, ml1, ofv1) && ok;
        // ---- cumulative positions: scan of the lane pairs (one packed scan when the lengths are small: 15 + 17 bits)
        const uint32_t pLL = ll0 + ll1, pE = pLL + ml0 + ml1;
        uint32_t xLL, xE;                                                  // exclusive prefix of this lane's pair
        if (__all_sync(kFullMask, pLL < 1024u && pE < 4096u)) {
            const uint32_t v = warp_scan_incl(pLL | (pE << 15), lane);
            xLL = (v & 0x7FFFu) - pLL; xE = (v >> 15) - pE;
        } else { xLL = warp_scan_incl(pLL, lane) - pLL; xE = warp_scan_incl(pE, lane) - pE; }
        const uint32_t LE0 = LEbase + xLL + ll0, LE1 = LE0 + ll1;
        const uint32_t E0 = Ebase + xE + ll0 + ml0, E1 = E0 + ll1 + ml1;
        // ---- repeat offsets: the history hops from one repeat code to the next, in position order
        uint32_t off0 = ofv0 - 3, off1 = ofv1 - 3;
        const uint32_t b0 = __ballot_sync(kFullMask, v0 && ofv0 <= 3), b1 = __ballot_sync(kFullMask, v1 && ofv1 <= 3);
        uint32_t pos = 0;                                                  // history is valid as of position `pos`
        auto off_at = [&](uint32_t q) { return __shfl_sync(kFullMask, (q & 1) ? off1 : off0, (q >> 1) & 31); };     // q is warp-uniform
        auto advance = [&](uint32_t upto) {                                // positions [pos, upto) are plain offsets: push the last three
            const uint32_t cnt = upto - pos;
            const uint32_t o1 = off_at(upto - 1), o2 = off_at(upto - 2), o3 = off_at(upto - 3);
            const uint32_t n0 = cnt >= 1 ? o1 : rep0, n1 = cnt >= 2 ? o2 : (cnt == 1 ? rep0 : rep1),
                           n2 = cnt >= 3 ? o3 : (cnt == 2 ? rep0 : (cnt == 1 ? rep1 : rep2));
            rep0 = n0; rep1 = n1; rep2 = n2; pos = upto;
        };
        uint32_t lanes = b0 | b1;
        while (lanes) {
            const uint32_t L = __ffs(lanes) - 1; lanes &= lanes - 1;
            for (uint32_t h = 0; h < 2; h++) {
                if (!(((h ? b1 : b0) >> L) & 1u)) continue;
                const uint32_t q = 2 * L + h;
                advance(q);
                const uint32_t ofr = __shfl_sync(kFullMask, h ? ofv1 : ofv0, L); const bool llz = __shfl_sync(kFullMask, h ? ll1 : ll0, L) == 0;
                const uint32_t o = rep_update(ofr, llz, rep0, rep1, rep2);
                if (lane == L) { if (h) off1 = o; else off0 = o; }
                pos = q + 1;
            }
        }
        advance(nv);
        ok = ok && !(ofv0 > 3 && off0 > kOffMax) && !(ofv1 > 3 && off1 > kOffMax) && LE1 <= lit_regen && E1 <= block_max;
        bad = __any_sync(kFullMask, v0 && !ok);
        if (bad) break;
        if (v1) {
            const uint64_t a0 = rec_pack(E0, LE0, off0), a1 = rec_pack(E1, LE1, off1);
            *(uint4*)(sq + i0) = make_uint4((uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32));
        } else if (v0) sq[i0] = rec_pack(E0, LE0, off0);
        Ebase = __shfl_sync(kFullMask, E1, 31); LEbase = __shfl_sync(kFullMask, LE1, 31);
    }
    const uint32_t rsize = Ebase + (lit_regen - LEbase);
    if (bad || rsize > block_max) { if (lane == 0) b.status = FZG_E_CORRUPT; return; }
    if (lane == 0) {
        b.rsize = rsize; b.rep_out[0] = rep0; b.rep_out[1] = rep1; b.rep_out[2] = rep2;
        sq[nseq] = rec_pack(rsize, lit_regen, 1);        // the TAIL record: the literals after the last sequence as one more literal run (k_execute_tile)
    }
}

// ------------------------------------------------------------------ offsets
__global__ void k_offsets(const Item* items, const ItemInfo* infos, const ItemBase* bases, Frame* frames, Block* blocks,
                          ItemOut* outs, uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemOut o;
    offsets_item(items[i], infos[i], bases[i], frames, blocks, o);
    outs[i] = o;
}

// ------------------------------------------------------------------ execute (LZ77)
// What bounds this stage is the dependency chain inside a frame (on text-like data almost every
// stretch of output copies something from just before it), so a frame is not spread over warps: ONE
// WARP owns a frame at a time (frames are drawn from a ticket) and executes its blocks and sequences
// in order, and the parallelism comes from the ~48 frames an SM has in flight.  There is no
// cross-warp synchronisation at all.
//
// A warp takes 32 sequences per round, one per lane.  The positional records (rec_e / rec_le) give
// every lane its literal run [S, M) and match [M, E) without a scan.  The round's output (a few
// hundred bytes) is assembled in a per-warp shared-memory stage and then written to HBM with
// 16-byte stores:
//   1. literal runs: every lane copies its own run from the literal buffer, 8 bytes per step;
//   2. matches: every lane copies its own match, 8 bytes per step (two aligned loads + funnel shift),
//      from HBM/L2 when the source lies before the round (earlier rounds, earlier blocks = the window)
//      or from the stage when it lies inside the round.  A lane whose source is not written yet
//      sits out; the frontier is the position reached by the first unfinished lane, which can always
//      advance (its sources lie below its own position), so the passes terminate;
//   3. flush.
// Offsets < 8 (the match overlaps its own 8-byte step) are expanded from the period; a sequence too
// large for the stage (long literal run or long match) is copied by the whole warp straight to HBM.
#ifndef FZ_EXEC_WARPS
#define FZ_EXEC_WARPS 4
#endif
#ifndef FZ_EXEC_CTAS
#define FZ_EXEC_CTAS 8
#endif
constexpr int kExecWarps = FZ_EXEC_WARPS;                    // warps (= frames in flight) per CTA
constexpr int kExecCtasPerSm = FZ_EXEC_CTAS;
#ifndef FZ_EXEC_STAGE
#define FZ_EXEC_STAGE 512
#endif
constexpr uint32_t kStage = FZ_EXEC_STAGE;                            // bytes of round output a warp assembles in shared memory
constexpr uint32_t kStageBytes = kStage + 48;                // + alignment slack (the stage mirrors the low 4 address bits) + load slack
constexpr uint32_t kFull = 0xFFFFFFFFu;

// nb (1..8) bytes starting at generic address g (HBM or shared memory); never touches an 8-byte word that holds no wanted byte
__device__ __forceinline__ uint64_t ld8_any(const uint8_t* g, uint32_t nb)
{
    const uintptr_t a = (uintptr_t)g & ~(uintptr_t)7;
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7);
    const uint2 w0 = *(const uint2*)a;
    uint2 w1 = make_uint2(0, 0);
    if (sh + nb > 8) w1 = *(const uint2*)(a + 8);
    return funnel8(w0.x, w0.y, w1.x, w1.y, sh);
}
// L2 residency (FZ_EXEC_L2HINT).  The window reads of ~4 700 frames in flight miss L2 almost always and every miss ALLOCATES a line:
// 55 GB of fills per 12 ms turn the 126 MB L2 over every ~26 us, in which a frame produces 2 KB -- so only matches closer than
// that find their source in L2 (15 % of them), although 47 % lie within the 26 KB that are a frame's fair share.  With the
// hint the streaming data (window misses, records, literals) is marked evict-first and leaves the plain output stores alone.
#ifndef FZ_EXEC_L2OUT
#define FZ_EXEC_L2OUT 0
#endif
#ifndef FZ_EXEC_L2HINT
#define FZ_EXEC_L2HINT 0
#endif
__device__ __forceinline__ uint64_t l2_stream_policy()
{
    uint64_t p = 0;
#if FZ_EXEC_L2HINT
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
// nb (1..8) bytes at global address g (never shared memory), streamed through L2
__device__ __forceinline__ uint64_t ld8_stream(const uint8_t* g, uint32_t nb, uint64_t pol)
{
#if FZ_EXEC_L2HINT
    const uintptr_t a = (uintptr_t)g & ~(uintptr_t)7;
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7);
    uint32_t x0, x1, x2 = 0, x3 = 0;
    asm volatile("ld.global.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(x0), "=r"(x1) : "l"(a), "l"(pol) : "memory");
    if (sh + nb > 8) asm volatile("ld.global.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(x2), "=r"(x3) : "l"(a + 8), "l"(pol) : "memory");
    return funnel8(x0, x1, x2, x3, sh);
#else
    (void)pol; return ld8_any(g, nb);
#endif
}
__device__ __forceinline__ uint64_t ldrec_stream(const uint64_t* p, uint64_t pol)
{
#if FZ_EXEC_L2HINT
    uint64_t v; asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
#else
    (void)pol; return __ldg(p);
#endif
}

// nb (1..8) low bytes of v -> the stage at any alignment: predicated byte stores through the shared window (no generic
// addressing, no branches).  (Word-sized red.shared.or on a zeroed stage was measured slower: 27.8 -> 31.3 ms.)
#define FZ_ST_BYTE(I, W, SH) asm volatile("{ .reg .pred q; .reg .b32 t; setp.gt.u32 q, %2, " #I "; shr.b32 t, %1, " #SH "; @q st.shared.u8 [%0+" #I "], t; }" ::"r"(a), "r"(W), "r"(nb) : "memory")
__device__ __forceinline__ void st_stage(uint8_t* p, uint64_t v, uint32_t nb)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    FZ_ST_BYTE(0, lo, 0); FZ_ST_BYTE(1, lo, 8); FZ_ST_BYTE(2, lo, 16); FZ_ST_BYTE(3, lo, 24);
    FZ_ST_BYTE(4, hi, 0); FZ_ST_BYTE(5, hi, 8); FZ_ST_BYTE(6, hi, 16); FZ_ST_BYTE(7, hi, 24);
}
#undef FZ_ST_BYTE

// the same through 32-bit shared-memory addresses (k_execute_cta: the stage and the bitmap are addressed from one conversion
// per block instead of one per access, and stage reads are LDS instead of generic loads)
#define FZ_ST_BYTE_S(I, W, SH) asm volatile("{ .reg .pred q; .reg .b32 t; setp.gt.u32 q, %2, " #I "; shr.b32 t, %1, " #SH "; @q st.shared.u8 [%0+" #I "], t; }" ::"r"(a), "r"(W), "r"(nb) : "memory")
__device__ __forceinline__ void st_stage_s(uint32_t a, uint64_t v, uint32_t nb)
{
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    FZ_ST_BYTE_S(0, lo, 0); FZ_ST_BYTE_S(1, lo, 8); FZ_ST_BYTE_S(2, lo, 16); FZ_ST_BYTE_S(3, lo, 24);
    FZ_ST_BYTE_S(4, hi, 0); FZ_ST_BYTE_S(5, hi, 8); FZ_ST_BYTE_S(6, hi, 16); FZ_ST_BYTE_S(7, hi, 24);
}
#undef FZ_ST_BYTE_S
#ifndef FZ_EXEC_ADDRSPACE
#define FZ_EXEC_ADDRSPACE 1          // k_execute: explicit ld.global / ld.shared / st.shared addresses instead of generic ones (measured, see profiles/r02_notes.md)
#endif
__device__ __forceinline__ uint64_t ld8_global(const uint8_t* g, uint32_t nb)       // nb (1..8) bytes at a GLOBAL address
{
    const uintptr_t a = (uintptr_t)g & ~(uintptr_t)7;
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7);
    uint32_t x0, x1, x2 = 0, x3 = 0;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(x0), "=r"(x1) : "l"(a) : "memory");
    if (sh + nb > 8) asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(x2), "=r"(x3) : "l"(a + 8) : "memory");
    return funnel8(x0, x1, x2, x3, sh);
}
__device__ __forceinline__ uint64_t ld8_shared(uint32_t g, uint32_t nb)
{
    const uint32_t a = g & ~7u, sh = g & 7u;
    uint32_t x0, x1, x2 = 0, x3 = 0;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x0), "=r"(x1) : "r"(a) : "memory");
    if (sh + nb > 8) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x2), "=r"(x3) : "r"(a + 8) : "memory");
    return funnel8(x0, x1, x2, x3, sh);
}

// warp-cooperative copy / fill, any alignment, any size
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const uint32_t nv = n >> 4;
        for (uint32_t i = lane; i < nv; i += 32) ((uint4*)dst)[i] = ((const uint4*)src)[i];
        for (uint32_t i = (nv << 4) + lane; i < n; i += 32) dst[i] = src[i];
    } else {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
}

// One sequence that does not fit the stage: literal run + match copied by the whole warp, straight to HBM.
// dst = output position of the sequence start, all arguments warp-uniform.
__device__ __forceinline__ void warp_big_sequence(uint8_t* dst, const uint8_t* lit, uint32_t ll, uint32_t ml, uint32_t off, uint32_t lane)
{
    warp_copy(dst, lit, ll, lane);
    __syncwarp();
    uint8_t* m = dst + ll;
    if (off == 0) return;                                    // flagged corrupt by the caller
    if (off >= 512) {                                        // source and destination of a 512-byte round never overlap
        const uint8_t* s = m - off;
        for (uint32_t i = 0; i < ml; i += 512) {
            const uint32_t nb = min(16u, ml > i + 16 * lane ? ml - i - 16 * lane : 0u);
            for (uint32_t k = 0; k < nb; k++) m[i + 16 * lane + k] = s[i + 16 * lane + k];
            __syncwarp();
        }
    } else {                                                 // periodic with period `off`: byte i equals byte i % off of the period
        const uint8_t* s = m - off;
        for (uint32_t i = lane; i < ml; i += 32) m[i] = s[i % off];
    }
    __syncwarp();
}

// The next round's match sources, asked for a round ahead.  FZ_EXEC_PREFETCH: 1 = prefetch.global.L2 of the first and last byte's
// line (round 1), 2 = cp.async.bulk.prefetch.L2 of exactly the 16-byte units that hold the match source (TMA unit, no registers).
#ifndef FZ_EXEC_PREFETCH
#define FZ_EXEC_PREFETCH 1
#endif
__device__ __forceinline__ void exec_prefetch(const uint8_t* sp, uint32_t len)
{
#if FZ_EXEC_PREFETCH == 1
    asm volatile("prefetch.global.L2 [%0];" ::"l"(sp));
    if ((((uintptr_t)sp + len - 1) ^ (uintptr_t)sp) & ~(uintptr_t)31) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + len - 1));
#elif FZ_EXEC_PREFETCH == 2
    const uintptr_t a = (uintptr_t)sp & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)((((uintptr_t)sp + len + 15) & ~(uintptr_t)15) - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes < 64u ? bytes : 64u) : "memory");
#endif
}

__device__ __forceinline__ void exec_block_warp(uint8_t* stage, const Block& b, const uint64_t* __restrict__ sq, uint8_t* g0,
                                                uint64_t done, int& status, uint32_t lane)
{
    const uint32_t nseq = b.nseq, rsize = b.rsize, lit_regen = b.lit_regen;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    const uint64_t pol = l2_stream_policy();
#if FZ_EXEC_ADDRSPACE
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
#endif
#if FZ_EXEC_L2OUT
    uint64_t opol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(opol));
#endif
    uint32_t Ecarry = 0, LEcarry = 0;
    uint64_t rcur = lane < nseq ? ldrec_stream(sq + lane, pol) : 0;          // records of the current round; the next round's are loaded a round early
    for (uint32_t g = 0; g < nseq;) {
        const uint32_t nv = min(32u, nseq - g);
        const uint64_t r = lane < nv ? rcur : 0;
        uint32_t E = rec_e(r), LE = rec_le(r);
        const uint32_t Elast = __shfl_sync(kFull, E, nv - 1), LElast = __shfl_sync(kFull, LE, nv - 1);
        if (lane >= nv) { E = Elast; LE = LElast; }
        uint32_t S = __shfl_up_sync(kFull, E, 1), LEp = __shfl_up_sync(kFull, LE, 1);
        if (lane == 0) { S = Ecarry; LEp = LEcarry; }
        const uint32_t gS = Ecarry;                              // output position where this round starts
        const uint32_t M = S + (LE - LEp);
        uint32_t off = lane < nv ? off_resolve(rec_off(r), in0, in1, in2) : 1;
        if (lane < nv && (uint64_t)off > done + M) { off = 0; status = FZG_E_CORRUPT; }     // reaches before the frame start
        // sequences of this round: the leading ones whose output fits the stage
        const uint32_t fit = __ballot_sync(kFull, lane < nv && E - gS <= kStage);
        const uint32_t m = fit == kFull ? 32u : (uint32_t)__ffs((int)~fit) - 1u;
        if (m == 0) {                                            // sequence g alone is larger than the stage
            const uint32_t ll0 = __shfl_sync(kFull, LE - LEp, 0), ml0 = __shfl_sync(kFull, E - M, 0), off0 = __shfl_sync(kFull, off, 0);
            warp_big_sequence(g0 + gS, lit + LEcarry, ll0, ml0, off0, lane);
            Ecarry = __shfl_sync(kFull, E, 0); LEcarry = __shfl_sync(kFull, LE, 0);
            g += 1;
            rcur = g + lane < nseq ? ldrec_stream(sq + g + lane, pol) : 0;
            continue;
        }
        rcur = g + m + lane < nseq ? ldrec_stream(sq + g + m + lane, pol) : 0;
        const bool mine = lane < m;
        const uint32_t gE = __shfl_sync(kFull, E, m - 1);         // end of the round's output
        const uint32_t a = (uint32_t)((uintptr_t)(g0 + gS) & 15); // stage[a + i] <-> g0[gS + i]: same low address bits as HBM
        uint8_t* const st = stage + a - gS;                       // st[p] is the stage byte of output position p (gS <= p < gE)
#if FZ_EXEC_ADDRSPACE
        const uint32_t sts = stage_s + a - gS;                    // the same as a shared-memory address
#define FZ_STAGE_ST(pos, v, nb) st_stage_s(sts + (pos), v, nb)
#define FZ_STAGE_LD(s, nb) ld8_shared(sts + (uint32_t)(s), nb)
#define FZ_GLOBAL_LD(p, nb) ld8_global(p, nb)
#else
#define FZ_STAGE_ST(pos, v, nb) st_stage(st + (pos), v, nb)
#define FZ_STAGE_LD(s, nb) ld8_any((const uint8_t*)st + (s), nb)
#define FZ_GLOBAL_LD(p, nb) ld8_stream(p, nb, pol)
#endif
        // ---- 1. literal runs
        {
            uint32_t pos = S; const uint8_t* src = lit + LEp;
            bool go = mine && pos < M;
            while (__any_sync(kFull, go)) {
                if (go) {
                    const uint32_t nb = min(8u, M - pos);
                    FZ_STAGE_ST(pos, FZ_GLOBAL_LD(src, nb), nb);
                    pos += nb; src += nb; go = pos < M;
                }
            }
        }
        __syncwarp();
        // ---- 2. matches
        {
            uint32_t pos = M;
            bool pending = mine && pos < E;
            uint32_t front = gS;                                  // every output byte below `front` is written (HBM or stage)
            while (__any_sync(kFull, pending)) {
                const uint32_t pm = __ballot_sync(kFull, pending);
                const uint32_t first = (uint32_t)__ffs((int)pm) - 1u;
                front = __shfl_sync(kFull, pos, first);           // lanes below `first` are complete, `first` has written up to pos
                bool go = pending;
                while (__any_sync(kFull, go)) {
                    if (go) {
                        uint32_t nb = min(8u, E - pos);
                        uint64_t v = 0;
                        if (off == 0) { /* corrupt: zeros */ }
                        else if (off < 8 && off < nb) {           // the step overlaps itself: expand the period byte by byte
                            const int32_t s0 = (int32_t)pos - (int32_t)off;
                            const bool ok = lane == first || (uint32_t)(s0 + (int32_t)off) <= front || s0 >= (int32_t)M;   // period written?
                            if (ok) {
                                const uint32_t take = s0 < (int32_t)gS ? min(off, gS - (uint32_t)s0) : off;   // a period straddling the round start
                                uint64_t pat = s0 < (int32_t)gS ? FZ_GLOBAL_LD((const uint8_t*)g0 + s0, take) : FZ_STAGE_LD(s0, take);
                                if (take < off) pat = (pat & ((1ull << (8 * take)) - 1)) | (FZ_STAGE_LD(gS, off - take) << (8 * take));
                                for (uint32_t i = 0; i < nb; i++) v |= ((pat >> (8 * (i % off))) & 0xFF) << (8 * i);
                            } else nb = 0;
                        } else {
                            const int32_t s = (int32_t)pos - (int32_t)off;
                            // available bytes: below `front`, or this lane's own match bytes written so far
                            const uint32_t lim = lane == first ? pos : ((s >= (int32_t)M) ? pos : front);
                            if (s < (int32_t)gS) {                 // before the round: HBM / L2 (earlier rounds, earlier blocks)
                                nb = min(nb, gS - (uint32_t)s);    // a step straddling the round start is split
                                v = FZ_GLOBAL_LD((const uint8_t*)g0 + s, nb);
                            } else if ((uint32_t)s + nb <= lim) v = FZ_STAGE_LD(s, nb);
                            else if ((uint32_t)s < lim) { nb = lim - (uint32_t)s; v = FZ_STAGE_LD(s, nb); }
                            else nb = 0;
                        }
                        if (nb) { FZ_STAGE_ST(pos, v, nb); pos += nb; go = pos < E; }
                        else go = false;                          // its source is still being produced by a lower lane
                    }
                }
                __syncwarp();
                pending = mine && pos < E;
            }
        }
        __syncwarp();
        // ---- 3. flush stage[a .. a + (gE - gS)) -> g0 + gS: head bytes, aligned 16-byte body, tail bytes
        {
            const uint32_t n = gE - gS;
            uint8_t* gd = g0 + gS;
            const uint32_t head = min(n, (16 - a) & 15);
            const uint32_t nvec = (n - head) >> 4;
            const uint32_t tail0 = head + (nvec << 4);
            // (explicit ld.shared / st.global here as well was measured slower: 25.6 -> 26.1 ms)
#if FZ_EXEC_L2OUT
            // the output IS the window of the sequences to come: stored with an L2 policy that keeps it resident longer than the
            // window lines that misses bring in (experiment, see profiles/r02_notes.md)
            if (lane < head) asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(gd + lane), "r"((uint32_t)stage[a + lane]), "l"(opol) : "memory");
            for (uint32_t i = lane; i < nvec; i += 32) {
                const uint4 v = *(const uint4*)(stage + a + head + 16 * i);
                asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(gd + head + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(opol) : "memory");
            }
            if (tail0 + lane < n) asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(gd + tail0 + lane), "r"((uint32_t)stage[a + tail0 + lane]), "l"(opol) : "memory");
#else
            if (lane < head) gd[lane] = stage[a + lane];
            for (uint32_t i = lane; i < nvec; i += 32) *(uint4*)(gd + head + 16 * i) = *(const uint4*)(stage + a + head + 16 * i);
            if (tail0 + lane < n) gd[tail0 + lane] = stage[a + tail0 + lane];
#endif
        }
        __syncwarp();
#undef FZ_STAGE_ST
#undef FZ_STAGE_LD
#undef FZ_GLOBAL_LD
        (void)st;
        Ecarry = gE; LEcarry = __shfl_sync(kFull, LE, m - 1);
        g += m;
#if FZ_EXEC_PREFETCH
        // The next round's match sources lie a random distance back in the window (HBM): ask for them now, a whole round
        // (thousands of cycles: ~48 warps share the SM) before they are read, so that the round finds them in L2.
        {
            const uint32_t nn = g < nseq ? min(32u, nseq - g) : 0u;
            const uint32_t En = rec_e(rcur), LEn = rec_le(rcur);
            uint32_t Sn = __shfl_up_sync(kFull, En, 1), LEpn = __shfl_up_sync(kFull, LEn, 1);
            if (lane == 0) { Sn = Ecarry; LEpn = LEcarry; }
            const uint32_t Mn = Sn + (LEn - LEpn);
            const uint32_t offn = off_resolve(rec_off(rcur), in0, in1, in2);
            if (lane < nn && offn != 0 && (uint64_t)offn <= done + Mn) exec_prefetch(g0 + Mn - offn, En - Mn);
        }
#endif
    }
    // literals after the last sequence
    warp_copy(g0 + Ecarry, lit + LEcarry, rsize - Ecarry, lane);
    (void)lit_regen;
    __syncwarp();
}

#include "fz_exec_alt.cuh"      // k_execute2, exec_block_steps, k_execute_pass<W>: the alternatives that were measured and not kept as defaults

template <bool STEPS>
__global__ void __launch_bounds__(kExecWarps * 32, kExecCtasPerSm) k_execute(Frame* frames, const Block* blocks, const Item* items,
                                                                              const ItemOut* outs, const uint64_t* seqs,
                                                                              uint32_t n_frames, uint32_t* ticket)
{
    __shared__ __align__(16) uint8_t s_stage[kExecWarps][kStageBytes];
    const uint32_t lane = threadIdx.x & 31;
    uint8_t* stage = s_stage[threadIdx.x >> 5];
    for (;;) {
        uint32_t f = 0;
        if (lane == 0) f = atomicAdd(ticket, 1);
        f = __shfl_sync(kFull, f, 0);
        if (f >= n_frames) return;
        Frame& fr = frames[f];
        if (outs[fr.item].fail) continue;
        uint8_t* const fbase = items[fr.item].dst + fr.out_off;
        uint64_t done = 0;
        int status = 0;
        for (uint32_t kb = 0; kb < fr.n_blocks; kb++) {
            const Block& b = blocks[fr.first_block + kb];
            uint8_t* const g0 = fbase + done;
            const uint32_t rsize = b.rsize;
            if (b.type == BT_RAW) warp_copy(g0, b.src, rsize, lane);
            else if (b.type == BT_RLE) {
                const uint8_t v = b.src[0];
                for (uint32_t i = lane; i < rsize; i += 32) g0[i] = v;
            } else if (b.nseq == 0) warp_copy(g0, b.lit, rsize, lane);
            else if (STEPS) exec_block_steps(stage, b, seqs + b.seq_base, g0, done, status, lane);
            else exec_block_warp(stage, b, seqs + b.seq_base, g0, done, status, lane);
            __syncwarp();                      // later blocks read this one back (the window)
            done += rsize;
        }
        status = __reduce_max_sync(kFull, status);
        if (lane == 0 && status) fr.status = status;
    }
}

// ------------------------------------------------------------------ execute, several warps per frame
// One warp per frame needs thousands of frames to fill the GPU and takes ~5 ms for a 1 MiB frame.  A batch with FEW
// frames (one file opened through the mount, config 4's large single-frame files) is executed by W warps per frame
// instead: a CTA owns a frame, a CTA round takes 32 W sequences (one per thread) into one shared-memory stage of
// W x 512 bytes.  A thread's literal run and match are placed by the positional records as before; what changes is
// how a thread learns that the source bytes of its match are there.  Sources below the round are in HBM (the
// previous flush is separated from this round by a barrier).  Sources inside the round are tracked by a BITMAP with
// one bit per stage byte: a writer sets the bits of the bytes it has stored (release), a reader takes the longest
// ready prefix of the (up to 8) bytes it wants (acquire) and retries later for the rest.  This is exact dataflow: no
// frontier, no ordering between lanes or warps, and the lowest unfinished sequence can always advance (everything
// below its position belongs to finished sequences or to itself), so the polling terminates.  An offset below 8 is
// widened to its first multiple >= 8 as soon as the match has produced that many bytes (the output is periodic from
// M - off on), so short periods also move 8 bytes per step after the first few.
// Two barriers per round: before the flush (the stage is complete) and after it (HBM readable, bitmap cleared).
template <int W> struct ExecCta {
    static constexpr bool hash = W == 8 || W == 16;                     // a checksum warp rides along (below)
    static constexpr uint32_t exec_threads = W * 32;
    static constexpr uint32_t threads = exec_threads + (hash ? 32 : 0);
    static constexpr uint32_t stage = W * kStage;
    static constexpr uint32_t stage_bytes = stage + 48;
    static constexpr uint32_t bitmap_words = stage_bytes / 32 + 2;     // one bit per stage byte (+ the word a window may spill into)
    static constexpr int ctas_per_sm = 32 / W > 0 ? 32 / W : 1;
};
// barrier among the executing warps only (the checksum warp never takes part)
template <int W> __device__ __forceinline__ void exec_sync()
{
    if constexpr (ExecCta<W>::hash) asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");
    else __syncthreads();
}

// stage bytes [o, o + nb) are stored: publish them (release: the byte stores above are visible to whoever sees the bits)
__device__ __forceinline__ void bm_mark(uint32_t bm_s, uint32_t o, uint32_t nb)          // bm_s: shared-memory address of the bitmap
{
    const uint32_t sh = o & 31u, m = (1u << nb) - 1u;
    const uint32_t a0 = bm_s + ((o >> 5) << 2);
    asm volatile("red.release.cta.shared.or.b32 [%0], %1;" ::"r"(a0), "r"(m << sh) : "memory");
    if (sh + nb > 32) asm volatile("red.relaxed.cta.shared.or.b32 [%0], %1;" ::"r"(a0 + 4), "r"(m >> (32 - sh)) : "memory");
}
// how many of the stage bytes starting at o are stored (0 .. 32)
__device__ __forceinline__ uint32_t bm_ready(uint32_t bm_s, uint32_t o)
{
    const uint32_t a0 = bm_s + ((o >> 5) << 2);
    uint32_t w0, w1;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(w0) : "r"(a0) : "memory");
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(w1) : "r"(a0 + 4) : "memory");
    const uint32_t bits = __funnelshift_r(w0, w1, o & 31u);
    return (uint32_t)__ffs((int)~bits) - 1u;                     // all 32 ready: 0xFFFFFFFF, larger than any request
}

// After a barrier among the executing warps: the frame's bytes below `upto` are in HBM, tell whoever hashes the frame --
// the CTA's own checksum warp (shared memory) or, for W = 32, the frame's checksum CTA on another SM (global memory).
struct ExecProg { volatile unsigned long long* s; unsigned long long* g; };
template <int W> __device__ __forceinline__ void exec_publish(const ExecProg& prog, uint64_t upto, uint32_t tid)
{
    if constexpr (ExecCta<W>::hash) {
        if (tid == 0) { __threadfence_block(); *prog.s = upto; }
    } else if constexpr (W == 32) {
        if (tid == 0 && prog.g) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(prog.g), "l"((unsigned long long)upto) : "memory");
    }
}

// One sequence that does not fit the stage, copied by the whole CTA straight to HBM (arguments CTA-uniform).
// Every byte below dst is readable on entry (barrier after the previous flush).
template <int W>
__device__ __forceinline__ void cta_big_sequence(uint8_t* dst, const uint8_t* lit, uint32_t ll, uint32_t ml, uint32_t off, uint32_t tid)
{
    constexpr uint32_t T = W * 32;
    group_copy(dst, lit, ll, tid, T);
    exec_sync<W>();
    uint8_t* m = dst + ll;
    if (off != 0) {                                          // 0: flagged corrupt by the caller
        const uint8_t* s = m - off;
        if (off >= 16 * T) {                                 // source and destination of a 16 T-byte round never overlap
            for (uint32_t i = 0; i < ml; i += 16 * T) {
                const uint32_t nb = min(16u, ml > i + 16 * tid ? ml - i - 16 * tid : 0u);
                for (uint32_t k = 0; k < nb; k++) m[i + 16 * tid + k] = s[i + 16 * tid + k];
                exec_sync<W>();
            }
        } else {                                             // periodic with period `off`; the period lies below m: written
            for (uint32_t i = tid; i < ml; i += T) m[i] = s[i % off];
        }
    }
    exec_sync<W>();
}

template <int W>
__device__ __forceinline__ void exec_block_cta(uint8_t* stage, uint32_t* bm, uint32_t* cnt, const ExecProg& prog, const Block& b,
                                               const uint64_t* __restrict__ sq, uint8_t* g0, uint64_t done, int& status, uint32_t tid)
{
    constexpr uint32_t T = W * 32, kCtaStage = ExecCta<W>::stage;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t nseq = b.nseq, rsize = b.rsize;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    for (uint32_t i = tid; i < ExecCta<W>::bitmap_words; i += T) bm[i] = 0;
    exec_sync<W>();
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage), bm_s = (uint32_t)__cvta_generic_to_shared(bm);
    uint32_t Ecarry = 0, LEcarry = 0;                            // CTA-uniform
    uint64_t rcur = tid < nseq ? __ldg(sq + tid) : 0;            // the round's records; the next round's are loaded a round early
    uint64_t rprev = (lane == 0 && warp > 0 && tid <= nseq) ? __ldg(sq + tid - 1) : 0;   // record before a warp's first one
    for (uint32_t g = 0; g < nseq;) {
        const uint32_t nv = min(T, nseq - g);
        const uint64_t rl = __ldg(sq + g + nv - 1);
        const uint32_t Elast = rec_e(rl), LElast = rec_le(rl);
        const uint64_t r = tid < nv ? rcur : 0;
        uint32_t E = rec_e(r), LE = rec_le(r);
        if (tid >= nv) { E = Elast; LE = LElast; }
        uint32_t S = __shfl_up_sync(kFull, E, 1), LEp = __shfl_up_sync(kFull, LE, 1);
        if (lane == 0) {
            if (warp == 0) { S = Ecarry; LEp = LEcarry; }
            else if (tid <= nv) { S = rec_e(rprev); LEp = rec_le(rprev); }
            else { S = Elast; LEp = LElast; }
        }
        const uint32_t gS = Ecarry;                              // output position where this round starts
        uint32_t off = tid < nv ? off_resolve(rec_off(r), in0, in1, in2) : 1;
        if (tid < nv && (uint64_t)off > done + S + (LE - LEp)) { off = 0; status = FZG_E_CORRUPT; }   // reaches before the frame start
        // sequences of this round: the leading ones whose output fits the stage (E never decreases)
        uint32_t n = nv;
        if (Elast - gS > kCtaStage) {
            const uint32_t fit = __ballot_sync(kFull, tid < nv && E - gS <= kCtaStage);
            if (lane == 0) cnt[warp] = (uint32_t)__popc(fit);
            exec_sync<W>();
            n = 0;
            for (int w = 0; w < W; w++) n += cnt[w];
            exec_sync<W>();
        }
        if (n == 0) {                                            // sequence g alone is larger than the stage
            const uint64_t r0 = __ldg(sq + g);
            const uint32_t E0 = rec_e(r0), LE0 = rec_le(r0), M0 = gS + (LE0 - LEcarry);
            uint32_t off0 = off_resolve(rec_off(r0), in0, in1, in2);
            if ((uint64_t)off0 > done + M0) off0 = 0;
            cta_big_sequence<W>(g0 + gS, lit + LEcarry, LE0 - LEcarry, E0 - M0, off0, tid);
            exec_publish<W>(prog, done + E0, tid);
            Ecarry = E0; LEcarry = LE0;
            g += 1;
            rcur = g + tid < nseq ? __ldg(sq + g + tid) : 0;
            rprev = (lane == 0 && warp > 0 && g + tid <= nseq) ? __ldg(sq + g + tid - 1) : 0;
            continue;
        }
        const uint64_t re = __ldg(sq + g + n - 1);
        const uint32_t gE = rec_e(re), LEend = rec_le(re);       // end of the round's output / literals
        rcur = g + n + tid < nseq ? __ldg(sq + g + n + tid) : 0;
        rprev = (lane == 0 && warp > 0 && g + n + tid <= nseq) ? __ldg(sq + g + n + tid - 1) : 0;
        const bool mine = tid < n;
        if (!mine) { S = gE; E = gE; LE = LEend; LEp = LEend; }
        const uint32_t M = S + (LE - LEp);
        const uint32_t a = (uint32_t)((uintptr_t)(g0 + gS) & 15); // stage[a + i] <-> g0[gS + i]: same low address bits as HBM
        const uint32_t st = stage_s + a - gS;                     // st + p: shared-memory address of the stage byte of output position p (gS <= p < gE)
        const uint32_t ob = a - gS;                               // ob + p: bitmap index of output position p
        // ---- 1. literal runs
        {
            uint32_t pos = S; const uint8_t* src = lit + LEp;
            bool go = mine && pos < M;
            while (__any_sync(kFull, go)) {
                if (go) {
                    const uint32_t nb = min(8u, M - pos);
                    st_stage_s(st + pos, ld8_any(src, nb), nb);
                    bm_mark(bm_s, ob + pos, nb);
                    pos += nb; src += nb; go = pos < M;
                }
            }
        }
        // ---- 2. matches
        {
            uint32_t pos = M;
            bool pending = mine && pos < E;
            const uint32_t eoff = off < 8 ? off * ((off + 7) / max(off, 1u)) : off;   // first multiple of a short period that is >= 8
            while (__any_sync(kFull, pending)) {
                bool moved = false;
                if (pending) {
                    uint32_t nb = min(8u, E - pos);
                    uint64_t v = 0;
                    if (off != 0) {                               // 0: corrupt, zeros
                        const uint32_t d = pos - M >= eoff - off ? eoff : off;        // the source stays at or above M - off
                        nb = min(nb, d);
                        const int32_t s = (int32_t)pos - (int32_t)d;
                        if (s < (int32_t)gS) {                    // before the round: HBM / L2 (earlier rounds, earlier blocks)
                            nb = min(nb, gS - (uint32_t)s);       // a step straddling the round start is split
                            v = ld8_any((const uint8_t*)g0 + s, nb);
                        } else {
                            nb = min(nb, bm_ready(bm_s, ob + (uint32_t)s));
                            if (nb) v = ld8_shared(st + (uint32_t)s, nb);
                        }
                    }
                    if (nb) { st_stage_s(st + pos, v, nb); bm_mark(bm_s, ob + pos, nb); pos += nb; pending = pos < E; moved = true; }
                }
                if (!__any_sync(kFull, moved)) __nanosleep(32);   // every unfinished lane waits for another warp
            }
        }
        exec_sync<W>();
        // ---- 3. flush stage[a .. a + (gE - gS)) -> g0 + gS: head bytes, aligned 16-byte body, tail bytes
        {
            const uint32_t nby = gE - gS;
            uint8_t* gd = g0 + gS;
            const uint32_t head = min(nby, (16 - a) & 15);
            if (tid < head) gd[tid] = stage[a + tid];
            const uint32_t nvec = (nby - head) >> 4;
            for (uint32_t i = tid; i < nvec; i += T) *(uint4*)(gd + head + 16 * i) = *(const uint4*)(stage + a + head + 16 * i);
            const uint32_t tail0 = head + (nvec << 4);
            if (tail0 + tid < nby) gd[tail0 + tid] = stage[a + tail0 + tid];
            for (uint32_t i = tid; i < ExecCta<W>::bitmap_words; i += T) bm[i] = 0;
        }
        exec_sync<W>();
        exec_publish<W>(prog, done + gE, tid);
        Ecarry = gE; LEcarry = LEend;
        g += n;
        // the next round's match sources: ask for them now (see k_execute)
        {
            const uint32_t nn = g < nseq ? min(T, nseq - g) : 0u;
            const uint32_t En = rec_e(rcur), LEn = rec_le(rcur);
            uint32_t Sn = __shfl_up_sync(kFull, En, 1), LEpn = __shfl_up_sync(kFull, LEn, 1);
            if (lane == 0) { if (warp == 0) { Sn = Ecarry; LEpn = LEcarry; } else { Sn = rec_e(rprev); LEpn = rec_le(rprev); } }
            const uint32_t Mn = Sn + (LEn - LEpn);
            const uint32_t offn = off_resolve(rec_off(rcur), in0, in1, in2);
            if (tid < nn && offn != 0 && (uint64_t)offn <= done + Mn && offn > Mn - Ecarry) {
                const uint8_t* sp = g0 + Mn - offn;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp));
                if ((((uintptr_t)sp + (En - Mn) - 1) ^ (uintptr_t)sp) & ~(uintptr_t)31) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (En - Mn) - 1));
            }
        }
    }
    // literals after the last sequence
    group_copy(g0 + Ecarry, lit + LEcarry, rsize - Ecarry, tid, T);
}

// Checksum beside the execution.  XXH64 is one serial multiply chain per frame (~40 cycles per 32-byte stripe, four
// accumulators = four lanes), which costs as much as the LZ77 itself once a frame has many warps, and a separate pass
// over a large frame is bound by HBM latency on top.  So the hash follows the executing warps through the frame: they
// publish how far the frame is flushed (release), the hashing warp takes the stripes below that mark straight from L2
// (acquire, ld.cg) and compares with the stored trailer at the end.
//   W = 8, 16: a 9th / 17th warp of the CTA (it shares the SM's issue slots with the polling warps, so its loop is kept
//              short: no software pipelining -- measured);
//   W = 32:    when the batch has at most half as many frames as the GPU has SMs, a second CTA per frame on an SM of
//              its own (all CTAs of the grid are resident, so waiting on each other is safe), at the full chain speed.
__device__ __forceinline__ uint64_t ldcg64_any(const uint8_t* q)
{
    const uintptr_t a = (uintptr_t)q & ~(uintptr_t)7;
    const uint32_t sh = (uint32_t)((uintptr_t)q & 7) * 8;
    const uint64_t w0 = __ldcg((const unsigned long long*)a);
    if (sh == 0) return w0;
    const uint64_t w1 = __ldcg((const unsigned long long*)(a + 8));
    return (w0 >> sh) | (w1 << (64 - sh));
}
// how far the frame is flushed (acquire: the bytes below the mark are visible; no fence, so loads in flight stay in flight)
template <bool REMOTE> __device__ __forceinline__ uint64_t hash_peek(const ExecProg& prog)
{
    unsigned long long v;
    if constexpr (REMOTE) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(prog.g) : "memory");
    else asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"((uint32_t)__cvta_generic_to_shared((const void*)prog.s)) : "memory");
    return v;
}
template <bool REMOTE> __device__ __forceinline__ uint64_t hash_wait(const ExecProg& prog, uint64_t need)
{
    uint64_t avail;
    while ((avail = hash_peek<REMOTE>(prog)) < need) __nanosleep(100);
    return avail;
}
template <bool REMOTE>
__device__ __forceinline__ bool hash_frame(const ExecProg& prog, const uint8_t* p, uint64_t len, uint32_t stored, uint32_t lane)
{
    const uint32_t j = lane & 3;                                   // eight redundant groups of four accumulators
    uint64_t acc = j == 0 ? XP1 + XP2 : (j == 1 ? XP2 : (j == 2 ? 0 : 0 - XP1));
    const uint64_t stripes = len >> 5;
    const uint8_t* q = p + 8 * j;
    const bool aligned = ((uintptr_t)p & 7) == 0;
    uint64_t cs = 0;
    while (cs < stripes) {
        uint64_t upto = min(hash_wait<REMOTE>(prog, (cs + 1) << 5) >> 5, stripes);
        if (!aligned) {
            for (; cs < upto; cs++) acc = xx_round(acc, ldcg64_any(q + 32 * cs));
        } else if (REMOTE && upto - cs >= 64) {
            // The data comes from L2 (~1000 cycles away), the chain takes ~40 cycles per stripe: keep 64 + 64 stripes in
            // flight.  One load instruction of the warp fetches eight stripes (lane = stripe-in-group * 4 + accumulator,
            // 256 contiguous bytes); a shuffle hands each word to its accumulator lane when its turn comes.
            uint64_t cur[8], nxt[8];
            const uint8_t* ql = p + 32 * (lane >> 2) + 8 * j;
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = __ldcg((const unsigned long long*)(ql + 32 * (cs + 8 * k)));
            for (;;) {
                if (upto - cs < 128) upto = min(hash_peek<REMOTE>(prog) >> 5, stripes);
                const bool more = upto - cs >= 128;
                if (more) {
#pragma unroll
                    for (int k = 0; k < 8; k++) nxt[k] = __ldcg((const unsigned long long*)(ql + 32 * (cs + 64 + 8 * k)));
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
#pragma unroll
                    for (int g = 0; g < 8; g++) acc = xx_round(acc, __shfl_sync(kFull, cur[k], g * 4 + j));
                }
                cs += 64;
                if (!more) break;
#pragma unroll
                for (int k = 0; k < 8; k++) cur[k] = nxt[k];
            }
        } else {
            for (; cs + 8 <= upto; cs += 8) {
                uint64_t x[8];
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = __ldcg((const unsigned long long*)(q + 32 * (cs + k)));
#pragma unroll
                for (int k = 0; k < 8; k++) acc = xx_round(acc, x[k]);
            }
            for (; cs < upto; cs++) acc = xx_round(acc, __ldcg((const unsigned long long*)(q + 32 * cs)));
        }
    }
    hash_wait<REMOTE>(prog, len);
    const uint64_t v1 = __shfl_sync(kFull, acc, 0), v2 = __shfl_sync(kFull, acc, 1), v3 = __shfl_sync(kFull, acc, 2), v4 = __shfl_sync(kFull, acc, 3);
    return (uint32_t)xx_combine(v1, v2, v3, v4, p, len) == stored;
}

// gprog (W = 32 only): null, or one progress word per frame (zeroed): the grid is then 2 n_frames CTAs, CTA f executes
// frame f and CTA n_frames + f hashes it.
template <int W>
__global__ void __launch_bounds__(ExecCta<W>::threads, ExecCta<W>::ctas_per_sm) k_execute_cta(Frame* frames, const Block* blocks, const Item* items,
                                                                                             const ItemOut* outs, const uint64_t* seqs,
                                                                                             uint32_t n_frames, uint32_t* ticket, int verify,
                                                                                             unsigned long long* gprog)
{
    constexpr uint32_t T = ExecCta<W>::exec_threads;
    constexpr bool H = ExecCta<W>::hash;
    __shared__ __align__(16) uint8_t s_stage[ExecCta<W>::stage_bytes];
    __shared__ uint32_t s_bm[ExecCta<W>::bitmap_words], s_cnt[W], s_f, s_bad;
    __shared__ unsigned long long s_prog;
    const uint32_t tid = threadIdx.x;
    const bool remote = W == 32 && gprog != nullptr;
    if (remote && blockIdx.x >= n_frames) {                        // a checksum CTA: one warp works
        if (tid >= 32) return;
        const uint32_t f = blockIdx.x - n_frames;
        Frame& fr = frames[f];
        if (outs[fr.item].fail || !verify || !fr.has_checksum) return;
        const ExecProg prog{ nullptr, gprog + f };
        if (!hash_frame<true>(prog, items[fr.item].dst + fr.out_off, fr.out_size, fr.checksum, tid) && tid == 0)
            atomicCAS(&fr.status, 0, FZG_E_CHECKSUM);             // an error found by the executing CTA wins, in either order
        return;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_f = remote ? blockIdx.x : atomicAdd(ticket, 1); s_prog = 0; s_bad = 0; }
        __syncthreads();
        const uint32_t f = s_f;
        if (f >= n_frames) return;
        Frame& fr = frames[f];
        const bool skip = outs[fr.item].fail;
        const ExecProg prog{ &s_prog, remote ? gprog + f : nullptr };
        uint8_t* const fbase = items[fr.item].dst + fr.out_off;
        if (skip) { }
        else if (H && tid >= T) {                                  // the checksum warp
            if (verify && fr.has_checksum && !hash_frame<false>(prog, fbase, fr.out_size, fr.checksum, tid & 31) && tid == T) s_bad = 1;
        } else {
            uint64_t done = 0;
            int status = 0;
            for (uint32_t kb = 0; kb < fr.n_blocks; kb++) {
                const Block& b = blocks[fr.first_block + kb];
                uint8_t* const g0 = fbase + done;
                const uint32_t rsize = b.rsize;
                if (b.type == BT_RAW) group_copy(g0, b.src, rsize, tid, T);
                else if (b.type == BT_RLE) {
                    const uint8_t v = b.src[0];
                    for (uint32_t i = tid; i < rsize; i += T) g0[i] = v;
                } else if (b.nseq == 0) group_copy(g0, b.lit, rsize, tid, T);
                else exec_block_cta<W>(s_stage, s_bm, s_cnt, prog, b, seqs + b.seq_base, g0, done, status, tid);
                exec_sync<W>();                    // later blocks read this one back (the window)
                done += rsize;
                exec_publish<W>(prog, done, tid);
            }
            status = __reduce_max_sync(kFull, status);
            if ((tid & 31) == 0 && status) fr.status = status;      // plain store: overrides a checksum verdict
        }
        if constexpr (H) {
            __syncthreads();
            if (tid == 0 && s_bad && !fr.status) fr.status = FZG_E_CHECKSUM;
        }
        if (remote) return;
    }
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

__global__ void k_finish(const ItemInfo* infos, const ItemBase* bases, const Frame* frames, const ItemOut* outs, ItemOut* host_outs,
                         uint32_t n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ItemOut o = outs[i];
    finish_item(infos[i], bases[i], frames, o);
    host_outs[i] = o;                          // pinned host memory (zero-copy)
}

}  // namespace fz

// ====================================================================== host side
using namespace fz;

static const char* kStageNames[] = { "count", "scan", "fill", "literals", "tables", "sequences", "records", "offsets", "execute", "checksum",
                                     "finish" };
const char* fzh_decode_stage_name(int s) { return s >= 0 && s < 11 ? kStageNames[s] : ""; }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "fzgpu: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return -5 /*-EIO*/; } } while (0)

static int g_sm_count = 148;
static int exec_warps_override()          // FZG_EXEC_W: the execute kernel, read per call (tests).  t128 / t256 / t512 / t1024: k_execute_tile with that many
{                                         // threads per frame (returned as -threads); 1: k_execute (warp per frame); 2..32: k_execute_cta<W>; else by batch shape
    const char* e = getenv("FZG_EXEC_W");
    if (e && e[0] == 't') { const int t = atoi(e + 1); return (t == 128 || t == 256 || t == 512 || t == 1024) ? -t : 0; }
    if (e && e[0] == 'p') { const int t = atoi(e + 1); return (t == 2 || t == 4 || t == 8) ? 200 + t : 0; }     // k_execute_pass<W>
    if (e && e[0] == 'd') return 65;                                  // k_execute2: two sequences per lane, half the warps per SM
    if (e && e[0] == 's') return 64;                                  // k_execute<true>: warp per frame, steps dealt out over the lanes
    const int v = e ? atoi(e) : 0;
    return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
}

int fzh_decode_setup(void)
{
    CK(cudaFuncSetAttribute(k_literals<kHufLogCommon>, cudaFuncAttributeMaxDynamicSharedMemorySize, LitCfg<kHufLogCommon>::smem));
    CK(cudaFuncSetAttribute(k_literals<kHufLogMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, LitCfg<kHufLogMax>::smem));
    CK(cudaFuncSetAttribute(k_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqSmem));
    CK(cudaFuncSetAttribute(k_execute_tile<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<128>::smem));
    CK(cudaFuncSetAttribute(k_execute_tile<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<256>::smem));
    CK(cudaFuncSetAttribute(k_execute_tile<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<512>::smem));
    CK(cudaFuncSetAttribute(k_execute_tile<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<1024>::smem));
    int dev = 0; CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

// Device -> host while the execute stage runs (FzStreamOut, fz_host.h).  Every frame of the chunk has an executing CTA that
// publishes (release, GPU scope) the number of bytes it has flushed; the words are read on a stream of their own every ~0.2 ms
// and each frame's newly finished bytes are queued on the copy-out stream in pieces of at least kPiece.  The copy engine reads
// HBM after the progress word that covers the bytes was observed, and the flush precedes the release: it sees final bytes.
// A frame that fails (corrupt, checksum) may leave part of its output in the caller's buffer; its status says so, as always.
static int stream_out(FzCtx* ctx, FzLane* c, uint32_t first, uint32_t n_frames, const Frame* d_frames, const unsigned long long* d_prog, cudaEvent_t last)
{
    constexpr uint64_t kPiece = 4ull << 20;
    FzStreamOut& so = ctx->so;
    int rc = c->h_prog.reserve((size_t)n_frames * (sizeof(Frame) + 16));
    if (rc) return rc;
    Frame* hf = (Frame*)c->h_prog.p;
    unsigned long long* hp = (unsigned long long*)(hf + n_frames);
    uint64_t* sent = (uint64_t*)(hp + n_frames);
    const Item* items = (const Item*)ctx->h_items.p + first;
    const ItemOut* outs = (const ItemOut*)ctx->h_outs.p + first;
    CK(cudaStreamWaitEvent(so.poll, c->ev_prog, 0));
    CK(cudaMemcpyAsync(hf, d_frames, (size_t)n_frames * sizeof(Frame), cudaMemcpyDeviceToHost, so.poll));
    for (uint32_t f = 0; f < n_frames; f++) sent[f] = 0;
    auto send = [&](uint32_t f, uint64_t upto) -> int {
        const Frame& fr = hf[f];
        uint8_t* h = (uint8_t*)so.dst[first + fr.item] + fr.out_off + sent[f];
        const uint8_t* d = items[fr.item].dst + fr.out_off + sent[f];
        CK(cudaMemcpyAsync(h, d, (size_t)(upto - sent[f]), cudaMemcpyDeviceToHost, so.copy));
        sent[f] = upto;
        return 0;
    };
    for (;;) {
        const cudaError_t q = cudaEventQuery(last);
        if (q != cudaSuccess && q != cudaErrorNotReady) return -5;
        if (q == cudaSuccess) break;
        CK(cudaMemcpyAsync(hp, d_prog, (size_t)n_frames * 8, cudaMemcpyDeviceToHost, so.poll));
        CK(cudaStreamSynchronize(so.poll));
        for (uint32_t f = 0; f < n_frames; f++) {
            const uint64_t avail = std::min<uint64_t>(hp[f], hf[f].out_size) & ~(uint64_t)4095;
            if (avail >= sent[f] + kPiece) { if ((rc = send(f, avail))) return rc; so.pieces++; }
        }
        struct timespec ts = { 0, 200000 }; nanosleep(&ts, nullptr);
    }
    CK(cudaStreamSynchronize(so.poll));
    // the kernels are done (k_finish wrote the per-item results into pinned memory): what is left of every good item
    for (uint32_t f = 0; f < n_frames; f++) {
        const Frame& fr = hf[f];
        if (outs[fr.item].status || fr.out_off + fr.out_size > outs[fr.item].dst_len) continue;
        if (fr.out_size > sent[f] && (rc = send(f, fr.out_size))) return rc;
    }
    so.done = true;
    return 0;
}

// Runs the whole pipeline for items [first, first + n) of c->h_items (Item records holding device
// pointers).  Results land in c->h_outs[first ..] (pinned).  Blocking on the context's stream.
int fzh_decode_run(FzCtx* ctx, int lane_idx, uint32_t first, uint32_t n, int flags, bool staggered)
{
    FzLane* c = &ctx->lane[lane_idx];
    cudaStream_t s = c->stream;
    const bool prof = flags & FZG_PROFILE;
    c->timing = fzg_timing_t{};
    struct Signal {                              // the next lane waits for this one's entropy stages: never leave it hanging
        FzCtx* ctx; FzLane* c; bool on;
        void fire() { if (on) { cudaEventRecord(c->ev_entropy, c->stream); c->entropy_epoch.store(ctx->epoch, std::memory_order_release); on = false; } }
        ~Signal() { fire(); }
    } signal{ ctx, c, staggered };
    if (n == 0) return 0;
    int rc;
    if ((rc = c->d_infos.reserve(n * sizeof(ItemInfo)))) return rc;
    if ((rc = c->d_bases.reserve(n * sizeof(ItemBase)))) return rc;
    if ((rc = c->d_outs.reserve(n * sizeof(ItemOut)))) return rc;
    if ((rc = c->d_totals.reserve(128))) return rc;
    // The per-call control data (item records in, totals and per-item results out) lives in pinned host memory that
    // the kernels access directly (UVA zero-copy).  A cudaMemcpyAsync would queue on the copy engines behind the
    // gigabyte-sized batch transfers of the neighbouring chunks (see run_batch) and stall the pipeline.
    const Item* d_items = (const Item*)ctx->h_items.p + first;
    ItemInfo* d_infos = (ItemInfo*)c->d_infos.p; ItemBase* d_bases = (ItemBase*)c->d_bases.p;
    ItemOut* d_outs = (ItemOut*)c->d_outs.p; uint64_t* d_totals = (uint64_t*)c->d_totals.p;
    uint32_t* d_tickets = (uint32_t*)(d_totals + 8);               // [0] sequences, [1] execute
    if ((rc = c->h_totals.reserve(64))) return rc;
    uint64_t* h_totals = (uint64_t*)c->h_totals.p;
    ItemOut* h_outs = (ItemOut*)ctx->h_outs.p + first;

    int ev = 0;
    auto mark = [&]() { if (prof || ev == 0) cudaEventRecord(c->ev[ev], s); ev++; };
    mark();                                                         // ev0: start
    const uint32_t tb = 128, gi = (n + tb - 1) / tb;
    k_count<<<gi, tb, 0, s>>>(d_items, d_infos, n, d_tickets); mark();
    k_scan<<<1, 512, 0, s>>>(d_infos, d_bases, h_totals, n); mark();
    CK(cudaStreamSynchronize(s));
    const uint64_t* tot = h_totals;
    const uint64_t n_frames = tot[0], n_blocks = tot[1], n_sj = tot[2], n_hj = tot[3], lit_bytes = tot[4], n_seq = tot[5];
    if (n_blocks >= (1ull << 31) || n_frames >= (1ull << 31)) return -22;
    if ((rc = c->d_frames.reserve((n_frames + 1) * sizeof(Frame)))) return rc;
    if ((rc = c->d_blocks.reserve((n_blocks + 1) * sizeof(Block)))) return rc;
    if ((rc = c->d_seq_jobs.reserve((n_sj + 1) * 4))) return rc;
    if ((rc = c->d_huf_jobs.reserve((2 * n_hj + 2) * 4))) return rc;      // the job list + the list of blocks deferred to the 2^12-cell launch
    if ((rc = c->d_lit.reserve(lit_bytes + 64))) return rc;
    if ((rc = c->d_seq.reserve((n_seq + 8) * 8))) return rc;
    if ((rc = c->d_seq_tabs.reserve((n_sj + 1) * (size_t)kJobTableBytes))) return rc;
    if ((rc = c->d_seq_hdrs.reserve((n_sj + 1) * sizeof(SeqJobHdr)))) return rc;
    Frame* d_frames = (Frame*)c->d_frames.p; Block* d_blocks = (Block*)c->d_blocks.p;
    uint32_t* d_sj = (uint32_t*)c->d_seq_jobs.p; uint32_t* d_hj = (uint32_t*)c->d_huf_jobs.p;
    uint64_t* d_seq = (uint64_t*)c->d_seq.p;
    uint8_t* d_tabs = (uint8_t*)c->d_seq_tabs.p; SeqJobHdr* d_hdrs = (SeqJobHdr*)c->d_seq_hdrs.p;

    k_fill<<<gi, tb, 0, s>>>(d_items, d_infos, d_bases, d_frames, d_blocks, d_sj, d_hj, (uint8_t*)c->d_lit.p, n); mark();
    int launches = 3;
    // Stagger the lanes: the entropy stages (bound by shared memory, few warps) of lane l start when lane l-1 has finished
    // its own and moved on to the LZ77 stage (bound by instruction issue, no shared memory), so the two kinds of work share
    // the SMs instead of two lanes fighting over the same resource.
    if (lane_idx > 0 && staggered) {
        FzLane& prev = ctx->lane[lane_idx - 1];
        while (prev.entropy_epoch.load(std::memory_order_acquire) < ctx->epoch) std::this_thread::yield();
        CK(cudaStreamWaitEvent(s, prev.ev_entropy, 0));
    }
    // A small batch fills neither the literal stage's SMs nor the sequence stage's: the two do not depend on each other,
    // so the literals run on a side stream beside tables / sequences / records (one file: 0.3 ms off the critical path).
    using L1 = LitCfg<kHufLogCommon>; using L2 = LitCfg<kHufLogMax>;
    const uint32_t lit_grid1 = (uint32_t)std::min<uint64_t>((n_hj + L1::groups - 1) / L1::groups, (uint64_t)g_sm_count);
    const uint32_t lit_grid2 = (uint32_t)std::min<uint64_t>((n_hj + L2::groups - 1) / L2::groups, (uint64_t)g_sm_count);
    const uint32_t seq_grid = (uint32_t)std::min<uint64_t>((n_sj + kSeqStreams - 1) / kSeqStreams, (uint64_t)g_sm_count);
    // (several lanes: the literal stage needs half an SM's registers and cannot share it with another lane's execute stage,
    //  so it waits on the side stream while tables / sequences / records of this lane run beside that stage)
    const bool side = n_hj && n_sj && (lit_grid1 + seq_grid <= (uint32_t)g_sm_count || staggered);
    cudaStream_t sl = side ? c->side : s;
    if (side) { CK(cudaEventRecord(c->ev_fork, s)); CK(cudaStreamWaitEvent(sl, c->ev_fork, 0)); }
    if (n_hj) {                                                      // tickets: [2] first launch, [3] deferred count, [4] second launch
        uint32_t* d_deferred = d_hj + n_hj + 1;
        k_literals<kHufLogCommon><<<lit_grid1, L1::threads, L1::smem, sl>>>(d_blocks, d_hj, (uint32_t)n_hj, nullptr, d_tickets + 2, d_deferred, d_tickets + 3);
        k_literals<kHufLogMax><<<lit_grid2, L2::threads, L2::smem, sl>>>(d_blocks, d_deferred, 0, d_tickets + 3, d_tickets + 4, nullptr, nullptr);
        launches += 2;
    }
    if (side) CK(cudaEventRecord(c->ev_join, sl));
    mark();
    if (n_sj) { k_seq_tables<<<(uint32_t)((n_sj + kTabThreads - 1) / kTabThreads), kTabThreads, 0, s>>>(d_blocks, d_sj, (uint32_t)n_sj, d_tabs, d_hdrs); launches++; }
    mark();
    if (n_sj) {
        k_sequences<<<seq_grid, kSeqThreads, kSeqSmem, s>>>(d_blocks, d_sj, (uint32_t)n_sj, d_seq, d_tickets, d_tabs, d_hdrs); launches++;
    }
    mark();
    if (n_sj) { k_records<<<(uint32_t)((n_sj + kRecWarps - 1) / kRecWarps), kRecWarps * 32, 0, s>>>(d_blocks, d_frames, d_sj, (uint32_t)n_sj, d_seq, d_tabs, d_hdrs); launches++; }
    mark();
    signal.fire();
    if (side) CK(cudaStreamWaitEvent(s, c->ev_join, 0));
    k_offsets<<<gi, tb, 0, s>>>(d_items, d_infos, d_bases, d_frames, d_blocks, d_outs, n); mark(); launches++;
    bool checksum_done = false;                                          // the CTA kernels with a checksum warp verify XXH64 themselves
    unsigned long long* frame_prog = nullptr;                            // per-frame flushed-bytes words, when this run keeps them (W = 32 + checksum CTAs)
    uint32_t n_tail = 0;                                                 // frames at the end of the batch executed (and hashed) by k_execute_cta<8>
    if (n_frames) {
        // Warps per frame, by batch shape (tools/exec_width_probe.py; ms for the whole pipeline on 1 MiB files, one warp
        // per frame -> chosen width: 1 file 8.1 -> 2.4, 64 files 9.6 -> 2.6, 148 files 10.4 -> 3.3, 592 files 10.8 -> 4.9,
        // 1184 files 11.7 -> 7.9; 16 x 64 MiB windowLog-23 frames 497 -> 59).  32 with a checksum CTA per frame while
        // every CTA can have an SM of its own, 8 (checksum warp inside the CTA) up to two waves of CTAs, 2 up to ~2400
        // frames; beyond that one warp per frame wins: k_execute has fewer instructions per sequence than the dataflow kernel.
        const int w_env = exec_warps_override();
        const uint64_t sms = (uint64_t)g_sm_count;
        // default: k_execute_tile (a CTA per frame, output-centric); threads per frame by batch shape: enough CTAs to fill the SMs
        // at 256 threads, else wider CTAs for the few frames there are.  FZG_EXEC_W selects the round-1 kernels instead.
        const int w = w_env ? w_env : (n_frames * 2 <= sms ? 32 : (n_frames <= sms * 8 ? 8 : (n_frames <= sms * 16 ? 2 : 1)));
        const int verify = (flags & FZG_NO_VERIFY_CHECKSUM) ? 0 : 1;
        auto cta = [&](auto wc) -> int {
            constexpr int W = decltype(wc)::value;
            uint32_t grid = (uint32_t)std::min<uint64_t>(n_frames, (uint64_t)g_sm_count * ExecCta<W>::ctas_per_sm);
            unsigned long long* gprog = nullptr;
            if (W == 32 && verify && n_frames * 2 <= (uint64_t)g_sm_count) {      // a checksum CTA per frame, every CTA on an SM of its own
                int r = c->d_prog.reserve(n_frames * 8);
                if (r) return r;
                gprog = (unsigned long long*)c->d_prog.p;
                CK(cudaMemsetAsync(gprog, 0, n_frames * 8, s));
                CK(cudaEventRecord(c->ev_prog, s));                  // k_offsets done (frames final), progress words zeroed
                frame_prog = gprog;
                grid = (uint32_t)n_frames * 2;
                checksum_done = true;
            }
            if (gprog) {
                // The checksum CTAs wait for the executing CTAs of the same grid: a COOPERATIVE launch makes the runtime guarantee that
                // every CTA of the grid is resident at once (it refuses the launch otherwise) instead of this code arguing it.
                Frame* a0 = d_frames; const Block* a1 = d_blocks; const Item* a2 = d_items; const ItemOut* a3 = d_outs; const uint64_t* a4 = d_seq;
                uint32_t a5 = (uint32_t)n_frames; uint32_t* a6 = d_tickets + 1; int a7 = verify; unsigned long long* a8 = gprog;
                void* args[] = { &a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8 };
                CK(cudaLaunchCooperativeKernel((const void*)k_execute_cta<W>, dim3(grid), dim3(ExecCta<W>::threads), args, 0, s));
            } else
                k_execute_cta<W><<<grid, ExecCta<W>::threads, 0, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, (uint32_t)n_frames, d_tickets + 1, verify, gprog);
            if (ExecCta<W>::hash) checksum_done = true;
            return 0;
        };
        auto tile = [&](auto tc) {
            constexpr int T = decltype(tc)::value;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(n_frames, (uint64_t)g_sm_count * TileCfg<T>::ctas_per_sm);
            k_execute_tile<T><<<grid, T, TileCfg<T>::smem, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, (uint32_t)n_frames, d_tickets + 1);
        };
        auto pass = [&](auto wc) {
            constexpr int W = decltype(wc)::value;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(n_frames, (uint64_t)g_sm_count * ExecPass<W>::ctas_per_sm);
            k_execute_pass<W><<<grid, ExecPass<W>::T, 0, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, (uint32_t)n_frames, d_tickets + 1);
        };
        if (w == 202) pass(std::integral_constant<int, 2>{});
        else if (w == 204) pass(std::integral_constant<int, 4>{});
        else if (w == 208) pass(std::integral_constant<int, 8>{});
        else if (w == 65) {
            const uint32_t grid = (uint32_t)std::min<uint64_t>((n_frames + kExecWarps - 1) / kExecWarps, (uint64_t)g_sm_count * (kExecCtasPerSm / 2));
            k_execute2<<<grid, kExecWarps * 32, 0, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, (uint32_t)n_frames, d_tickets + 1);
        }
        else if (w == -128) tile(std::integral_constant<int, 128>{});
        else if (w == -256) tile(std::integral_constant<int, 256>{});
        else if (w == -512) tile(std::integral_constant<int, 512>{});
        else if (w == -1024) tile(std::integral_constant<int, 1024>{});
        else if (w == 1 || w == 64) {
            // One warp per frame runs the batch in waves of 32 x SMs frames (equal-sized files finish in step), and a last wave with
            // few frames costs a whole frame time (5.7 ms per MiB) on a nearly empty GPU: 10 000 files = 2 waves + 528 frames took
            // 27.3 ms against 24.6 ms for 9 472.  So the frames beyond the last full wave, when they are few (up to 1/8 of the batch),
            // go to k_execute_cta<8> on the low-priority side stream, launched right after k_execute: its CTAs find no room until
            // k_execute's start to retire and then fill what the drain leaves idle (they hash their frames themselves; k_checksum
            // covers the rest).  Measured: 10 000 files 27.3 -> 26.0 ms with 625 frames moved (416 or 500, still three waves:
            // 27.4 / 27.2; 833, 1 250, 1 667: 26.4 / 26.9 / 27.8); 5 000 files 17.0 -> 13.4 ms; moving frames out of a batch of
            // exactly one wave (4 736) costs 4 %, out of a well-filled last wave (6 000, 7 200 files) changes nothing.
            static const bool tail_on = [] { const char* e = getenv("FZG_EXEC_TAIL"); return !e || atoi(e) != 0; }();
            if (tail_on && (!w_env || w_env == 64) && n_frames >= sms * 32) {
                const uint64_t rest = n_frames % (sms * 32);
                if (rest && rest <= n_frames / 8) n_tail = (uint32_t)(rest + n_frames / 128);       // a well-filled last wave is left alone
            }
            const uint32_t n1 = (uint32_t)n_frames - n_tail;
            if (n_tail) CK(cudaEventRecord(c->ev_fork, s));
            const uint32_t grid = (uint32_t)std::min<uint64_t>((n1 + kExecWarps - 1) / kExecWarps, (uint64_t)g_sm_count * kExecCtasPerSm);
            if (w == 64) k_execute<true><<<grid, kExecWarps * 32, 0, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, n1, d_tickets + 1);
            else k_execute<false><<<grid, kExecWarps * 32, 0, s>>>(d_frames, d_blocks, d_items, d_outs, d_seq, n1, d_tickets + 1);
            if (n_tail) {
                CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
                k_execute_cta<8><<<std::min<uint32_t>(n_tail, (uint32_t)sms * ExecCta<8>::ctas_per_sm), ExecCta<8>::threads, 0, c->side>>>(d_frames + n1, d_blocks, d_items, d_outs, d_seq, n_tail, d_tickets + 5, verify, nullptr);
                CK(cudaEventRecord(c->ev_join, c->side));
                CK(cudaStreamWaitEvent(s, c->ev_join, 0));
                launches++;
            }
        }
        else if (w == 2) rc = cta(std::integral_constant<int, 2>{});
        else if (w == 4) rc = cta(std::integral_constant<int, 4>{});
        else if (w == 8) rc = cta(std::integral_constant<int, 8>{});
        else if (w == 16) rc = cta(std::integral_constant<int, 16>{});
        else rc = cta(std::integral_constant<int, 32>{});
        if (rc) return rc;
        launches++;
    }
    mark();
    if (n_frames > n_tail && !(flags & FZG_NO_VERIFY_CHECKSUM) && !checksum_done) { k_checksum<<<(uint32_t)(((n_frames - n_tail) * 4 + 127) / 128), 128, 0, s>>>(d_frames, d_items, d_outs, (uint32_t)(n_frames - n_tail)); launches++; }
    mark();
    k_finish<<<gi, tb, 0, s>>>(d_infos, d_bases, d_frames, d_outs, h_outs, n); launches++;
    if (!prof) ev = 11;
    cudaEventRecord(c->ev[ev], s);                                   // last event
    if (ctx->so.dst && frame_prog && lane_idx == 0) {
        if ((rc = stream_out(ctx, c, first, (uint32_t)n_frames, d_frames, frame_prog, c->ev[ev]))) return rc;
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    c->timing.launches = launches;
    cudaEventElapsedTime(&c->timing.total_ms, c->ev[0], c->ev[ev]);
    if (prof) for (int k = 0; k < 11; k++) cudaEventElapsedTime(&c->timing.kernel_ms[k], c->ev[k], c->ev[k + 1]);
    return 0;
}
