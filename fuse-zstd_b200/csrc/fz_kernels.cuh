/*
 * fz_kernels.cuh -- warp-level stages of the decoder that are shared between the CUDA kernels
 * (fz_decode.cu) and the test-only host emulation (tests/emul), which replays them thread by thread.
 * The LZ77 execute pass is warp-cooperative CUDA and lives in fz_decode.cu only.
 */
#pragma once
#include "fz_core.cuh"

namespace fz {

struct ItemOut { uint64_t dst_len; int32_t status; int32_t fail; };

#ifdef __CUDACC__
static __constant__ SeqConsts c_seq_consts = { FZ_LL_BASE, FZ_ML_BASE, FZ_LL_BITS, FZ_ML_BITS, FZ_LL_DEF, FZ_OF_DEF, FZ_ML_DEF };
#endif

// ------------------------------------------------------------------ literals pass (four threads per block)
// Per block (group of four threads) in shared memory: the decode table (8 KB) and 1 KB that serves first as
// table-build scratch and then as the four bitstream rings.
FZ_HD constexpr uint32_t lit_group_bytes(int table_log) { return (1u << table_log) * 2 + 1024; }
#ifndef FZ_HUF_LOG_COMMON
#define FZ_HUF_LOG_COMMON 11             // (tests build a variant with 9 so that ordinary frames exercise the second pass)
#endif
constexpr int kHufLogCommon = FZ_HUF_LOG_COMMON;   // 11: what libzstd's encoder emits at most; deeper trees (the format allows 12) take a second pass
struct LitScratch {          // overlays the 1 KB ring area while the table is being built
    uint8_t w[256]; uint32_t ft[64]; uint16_t cnt[64]; uint32_t start[kHufLogMax + 2], rank[kHufLogMax + 2];
};

// thread 0 of the group: tree description (possibly from an earlier block: Treeless) -> table of 1 << log cells.
// log == -1 reports a malformed description, log == -2 a tree deeper than max_log (no table is written).
FZ_HD void lit_build(const Block* blocks, const Block& b, uint16_t* table, LitScratch& sc, int max_log, int& log, uint32_t& used)
{
    const Block& sb = blocks[b.huf_src];
    int nw; HufInfo hi; hi.log = 0; hi.used = 0;
    if (huf_read_weights(sb.src + sb.lit_hdr, sb.lit_csize, sc.w, nw, hi, sc.ft, sc.cnt) != 0) { log = -1; used = 0; return; }
    if (hi.log > max_log) { log = -2; used = 0; return; }
    huf_fill_table(table, sc.w, nw, hi.log, sc.start, sc.rank);
    log = hi.log; used = hi.used;
}

// What thread `sub` (0..3) of the group has to do for block b: nothing, an RLE fill, or one Huffman stream.
struct LitWork { const uint8_t* p; uint8_t* out; uint32_t n, n_out; int kind; };   // kind: 0 none, 1 rle, 2 huffman stream, -1 corrupt
FZ_HD LitWork lit_plan(const Block& b, uint32_t sub, int log, uint32_t used)
{
    LitWork wk{ nullptr, const_cast<uint8_t*>(b.lit), 0, 0, 0 };
    if (b.lit_type == LT_RLE) { wk.kind = 1; wk.p = b.src + b.lit_hdr; wk.n_out = b.lit_regen; return wk; }
    if (log < 0) { wk.kind = -1; return wk; }
    const uint32_t skip = b.lit_type == LT_HUF ? used : 0;
    if (skip > b.lit_csize) { wk.kind = -1; return wk; }
    const uint8_t* p = b.src + b.lit_hdr + skip;
    const uint32_t n = b.lit_csize - skip, regen = b.lit_regen;
    if (b.lit_streams == 1) { if (sub == 0) { wk.kind = 2; wk.p = p; wk.n = n; wk.n_out = regen; } return wk; }
    if (n < 10) { wk.kind = -1; return wk; }
    const uint32_t l1 = p[0] | ((uint32_t)p[1] << 8), l2 = p[2] | ((uint32_t)p[3] << 8), l3 = p[4] | ((uint32_t)p[5] << 8);
    const uint32_t seg = (regen + 3) / 4;
    if (6 + l1 + l2 + l3 > n || seg * 3 > regen) { wk.kind = -1; return wk; }
    const uint32_t l4 = n - 6 - l1 - l2 - l3;
    const uint32_t start = sub == 0 ? 0 : (sub == 1 ? l1 : (sub == 2 ? l1 + l2 : l1 + l2 + l3));
    wk.kind = 2; wk.p = p + 6 + start; wk.n = sub == 0 ? l1 : (sub == 1 ? l2 : (sub == 2 ? l3 : l4));
    wk.out += sub * seg; wk.n_out = sub == 3 ? regen - 3 * seg : seg;
    return wk;
}

// Executes the plan; ring = this thread's 256 bytes; bound / mask as in huf_decode_stream.  Returns 0 or 1 (corrupt).
FZ_HD int lit_run(const LitWork& wk, uint32_t sub, const uint16_t* table, int log, uint8_t* ring, uint32_t bound, uint32_t mask)
{
    if (wk.kind == 1) { const uint8_t v = wk.p[0]; for (uint32_t i = sub; i < wk.n_out; i += 4) wk.out[i] = v; }
    const int r = huf_decode_stream(table, log, wk.p, wk.n, wk.out, wk.n_out, ring, bound, mask, wk.kind == 2);
    return wk.kind == 2 ? (r != 0) : (wk.kind < 0);
}

// ------------------------------------------------------------------ sequences pass, stage A (one thread per block)
// mem = kChainBytes of shared memory for this stream; bound / mask: see decode_sequences_chain.
FZ_HD void seq_chain_thread(Block& b, const uint8_t* gtab, const SeqJobHdr& h, uint8_t* mem, uint64_t* seqs, uint32_t bound, uint32_t mask)
{
    const int st = decode_sequences_chain(b, gtab, h, mem, seqs + b.seq_base, bound, mask);
    if (st && !b.status) b.status = st;
}

// ------------------------------------------------------------------ checksum pass (four threads per frame)
#ifndef FZ_XX_DEPTH
#define FZ_XX_DEPTH 32
#endif
constexpr int kXxDepth = FZ_XX_DEPTH;            // loads in flight per accumulator chain of k_checksum (see xx_lane)
FZ_HD uint64_t xx_lane(const uint8_t* p, uint64_t len, uint32_t j)   // accumulator j over all 32-byte stripes
{
    uint64_t acc = j == 0 ? XP1 + XP2 : (j == 1 ? XP2 : (j == 2 ? 0 : 0 - XP1));
    const uint64_t stripes = len >> 5;
    const uint8_t* q = p + 8 * j;
    if (((uintptr_t)p & 7) == 0) {
        // a serial multiply chain fed from HBM: two batches of kXxDepth loads, the next one in flight while this one is consumed
        uint64_t s = 0;
        uint64_t cur[kXxDepth], nxt[kXxDepth];
        const bool any = stripes >= kXxDepth;
        if (any) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < kXxDepth; k++) cur[k] = *(const uint64_t*)(q + 32 * k);
        }
        for (; s + 2 * kXxDepth <= stripes; s += kXxDepth) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < kXxDepth; k++) nxt[k] = *(const uint64_t*)(q + 32 * (s + kXxDepth + k));
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < kXxDepth; k++) acc = xx_round(acc, cur[k]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < kXxDepth; k++) cur[k] = nxt[k];
        }
        if (any) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < kXxDepth; k++) acc = xx_round(acc, cur[k]);
            s += kXxDepth;
        }
        for (; s < stripes; s++) acc = xx_round(acc, *(const uint64_t*)(q + 32 * s));
    } else {
        for (uint64_t s = 0; s < stripes; s++) acc = xx_round(acc, rd64u(q + 32 * s));
    }
    return acc;
}
FZ_HD uint64_t xx_combine(uint64_t v1, uint64_t v2, uint64_t v3, uint64_t v4, const uint8_t* p, uint64_t len)
{
    uint64_t h;
    if (len >= 32) {
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
    } else h = XP5;
    const uint64_t tail = len & 31;
    return xx_finish(h, p + (len - tail), tail, len);
}

// ------------------------------------------------------------------ offsets pass (one thread per item)
// After the literal and sequence passes every block knows its regenerated size.  Assigns each
// block / frame its position in the item's dst, folds the first error (in stream order) into the
// item status and checks Frame_Content_Size and the destination capacity.
FZ_HD void offsets_item(const Item& it, const ItemInfo& info, const ItemBase& base, Frame* frames, Block* blocks,
                        ItemOut& out)
{
    uint64_t pos = 0; int status = 0;
    for (uint32_t f = 0; f < info.n_frames && !status; f++) {
        Frame& fr = frames[base.frame + f];
        fr.out_off = pos;
        uint64_t fsize = 0;
        uint32_t r0 = 1, r1 = 4, r2 = 8;              // RFC 8878 3.1.1.5: history at the start of a frame
        for (uint32_t k = 0; k < fr.n_blocks; k++) {
            Block& b = blocks[fr.first_block + k];
            if (b.status) { status = b.status; break; }
            if (b.rsize > fr.block_max) { status = FZG_E_CORRUPT; break; }
            if (it.dst_cap - pos < b.rsize) { status = FZG_E_DSTSIZE; break; }
            b.out_off = pos; pos += b.rsize; fsize += b.rsize;
            if (b.type == BT_COMPRESSED && b.nseq) {
                b.rep_in[0] = r0; b.rep_in[1] = r1; b.rep_in[2] = r2;
                const uint32_t n0 = off_resolve(b.rep_out[0], r0, r1, r2), n1 = off_resolve(b.rep_out[1], r0, r1, r2),
                               n2 = off_resolve(b.rep_out[2], r0, r1, r2);
                r0 = n0; r1 = n1; r2 = n2;
            }
        }
        if (status) break;
        fr.out_size = fsize;
        if (fr.has_fcs && fr.fcs != fsize) { status = FZG_E_FCS; break; }
    }
    if (!status) status = info.walk_status;
    out.dst_len = status ? 0 : pos; out.status = status; out.fail = status != 0;
}

// finish pass (one thread per item): errors found while executing / checksumming, in frame order
FZ_HD void finish_item(const ItemInfo& info, const ItemBase& base, const Frame* frames, ItemOut& o)
{
    if (o.fail) return;
    for (uint32_t f = 0; f < info.n_frames; f++) {
        int st = frames[base.frame + f].status;
        if (st) { o.status = st; o.fail = 1; o.dst_len = 0; return; }
    }
}

}  // namespace fz
