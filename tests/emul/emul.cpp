/*
 * emul.cpp -- TEST ONLY.  Compiles the decoder's per-thread device code (fz_core.cuh,
 * fz_kernels.cuh) with g++ and replays fz_decode.cu's launch sequence thread by thread on the
 * host, so that descriptor/pipeline logic can be checked against the oracle on a machine with
 * no GPU.  It is never linked into libfzgpu.so and is not a fallback: the product path is CUDA
 * only.  Warp-level stages run with a one-lane "warp" policy.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../fuse-zstd_b200/csrc/fz_kernels.cuh"
#include "../../fuse-zstd_b200/csrc/fz_enc_core.cuh"

using namespace fz;

static uint64_t g_far_records = 0;     // RAW records in the far form (extra bits > 32) seen so far
extern "C" uint64_t fze_far_records(void) { return g_far_records; }
// table logs seen so far: [0] LL, [1] OF, [2] ML, 16 counters each (what a stream of stage A really needs of its shared memory)
static uint64_t g_table_logs[3][16];
extern "C" void fze_table_logs(uint64_t* out) { for (int t = 0; t < 3; t++) for (int l = 0; l < 16; l++) out[t * 16 + l] = g_table_logs[t][l]; }
static const SeqConsts kConsts = { FZ_LL_BASE, FZ_ML_BASE, FZ_LL_BITS, FZ_ML_BITS, FZ_LL_DEF, FZ_OF_DEF, FZ_ML_DEF };

// TEST-ONLY serial restatement of stage B (k_records in fz_decode.cu is warp-parallel CUDA): RAW records ->
// positional records + span index + block totals, through the same helpers (raw_unpack, rep_update, rec_pack).
static void records_serial(const Block* blocks, Block& b, uint32_t block_max, uint64_t* seqs, const uint8_t* tab, const SeqJobHdr& h)
{
    if (b.status) return;
    uint64_t* sq = seqs + b.seq_base;
    const uint8_t* yLL = tab + kChainCellBytes; const uint8_t* yML = yLL + 512;
    const uint8_t* bits = b.src + h.bits_off;
    uint32_t rep0 = off_sym(0), rep1 = off_sym(1), rep2 = off_sym(2), E = 0, LE = 0;
    for (uint32_t i = 0; i < b.nseq; i++) {
        uint32_t ll, ml, ofv;
        bool ok = raw_unpack(sq[i], kConsts, yLL, yML, bits, ll, ml, ofv);
        const uint32_t off = rep_update(ofv, ll == 0, rep0, rep1, rep2);
        LE += ll; E += ll + ml;
        if (!ok || (ofv > 3 && off > kOffMax) || LE > b.lit_regen || E > block_max) { b.status = FZG_E_CORRUPT; return; }
        sq[i] = rec_pack(E, LE, off);
    }
    const uint32_t rsize = E + (b.lit_regen - LE);
    if (rsize > block_max) { b.status = FZG_E_CORRUPT; return; }
    b.rsize = rsize; b.rep_out[0] = rep0; b.rep_out[1] = rep1; b.rep_out[2] = rep2;
}

// TEST-ONLY serial executor over the records the sequence pass emits (the product's execute pass is the
// warp-cooperative CUDA kernel k_execute; this checks the record / span / repeat-offset logic on the CPU).
static void exec_frame_serial(Frame& fr, const Block* blocks, const Item& it, const uint64_t* seqs)
{
    uint8_t* const fbase = it.dst + fr.out_off;
    uint64_t done = 0;
    for (uint32_t k = 0; k < fr.n_blocks; k++) {
        const Block& b = blocks[fr.first_block + k];
        uint8_t* out = fbase + done;
        if (b.type == BT_RAW) memcpy(out, b.src, b.rsize);
        else if (b.type == BT_RLE) memset(out, b.src[0], b.rsize);
        else {
            const uint64_t* sq = seqs + b.seq_base;
            uint32_t S = 0, LEp = 0;
            for (uint32_t i = 0; i < b.nseq; i++) {
                const uint32_t E = rec_e(sq[i]), LE = rec_le(sq[i]);
                const uint32_t off = off_resolve(rec_off(sq[i]), b.rep_in[0], b.rep_in[1], b.rep_in[2]);
                const uint32_t ll = LE - LEp, M = S + ll;
                memcpy(out + S, b.lit + LEp, ll);
                if ((uint64_t)off > done + M) { fr.status = FZG_E_CORRUPT; return; }
                for (uint32_t q = M; q < E; q++) out[q] = out[(int64_t)q - off];
                S = E; LEp = LE;
            }
            memcpy(out + S, b.lit + LEp, b.lit_regen - LEp);
        }
        done += b.rsize;
    }
}

extern "C" int fze_decode_batch(size_t n, const void* const* src, const size_t* src_len, void* const* dst,
                                const size_t* dst_cap, size_t* dst_len, int* status, int flags)
{
    std::vector<Item> items(n);
    for (size_t i = 0; i < n; i++) items[i] = Item{ (const uint8_t*)src[i], src_len[i], (uint8_t*)dst[i], dst_cap[i] };
    // count
    std::vector<ItemInfo> infos(n);
    for (size_t i = 0; i < n; i++) walk_item<false>((uint32_t)i, items[i], infos[i], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    // scan
    std::vector<ItemBase> bases(n);
    ItemBase run{ 0, 0, 0, 0, 0, 0 };
    for (size_t i = 0; i < n; i++) {
        bases[i] = run;
        run.frame += infos[i].n_frames; run.block += infos[i].n_blocks; run.seq_job += infos[i].n_seq_jobs;
        run.huf_job += infos[i].n_huf_jobs; run.lit += infos[i].lit_bytes; run.seq += infos[i].n_seq;
    }
    std::vector<Frame> frames(run.frame + 1);
    std::vector<Block> blocks(run.block + 1);
    std::vector<uint32_t> seq_jobs(run.seq_job + 1), huf_jobs(run.huf_job + 1);
    std::vector<uint8_t> lit(run.lit + 64);
    std::vector<uint64_t> seqs(run.seq + 8);
    // fill
    for (size_t i = 0; i < n; i++) {
        ItemInfo tmp;
        walk_item<true>((uint32_t)i, items[i], tmp, &bases[i], frames.data(), blocks.data(), seq_jobs.data(), huf_jobs.data(), lit.data());
        if (memcmp(&tmp, &infos[i], sizeof tmp) != 0) return -1000;   // the two passes must agree
    }
    // literals: group of four threads per job
    std::vector<uint16_t> table(1 << kHufLogMax);
    alignas(16) static uint8_t ring[256]; static LitScratch sc;
    for (uint32_t j = 0; j < run.huf_job; j++) {
        Block& b = blocks[huf_jobs[j]];
        int log = 0; uint32_t used = 0;
        if (b.lit_type >= LT_HUF) lit_build(blocks.data(), b, table.data(), sc, kHufLogMax, log, used);
        for (uint32_t sub = 0; sub < 4; sub++) {
            const LitWork wk = lit_plan(b, sub, log, used);
            if (lit_run(wk, sub, table.data(), log, ring, wk.n_out / 4, 1)) b.status = FZG_E_CORRUPT;
        }
    }
    // sequences
    alignas(16) static uint8_t chain_mem[kChainBytes];
    for (uint32_t j = 0; j < run.seq_job; j++) {
        Block& b = blocks[seq_jobs[j]];
        alignas(16) static uint8_t tab[kJobTableBytes]; SeqJobHdr h;
        static uint8_t w_sym[512]; static int16_t w_norm[64]; static uint16_t w_cnt[64];
        const TabWork tw{ { w_sym, 1 }, { w_norm, 1 }, { w_cnt, 1 } };
        seq_tables_thread(blocks.data(), b, kConsts, tab, h, tw);
        if (h.bad && !b.status) b.status = FZG_E_CORRUPT;
        if (!h.bad) { g_table_logs[0][h.logLL & 15]++; g_table_logs[1][h.logOF & 15]++; g_table_logs[2][h.logML & 15]++; }
        seq_chain_thread(b, tab, h, chain_mem, seqs.data(), b.nseq - 1, 1);
        if (!b.status) for (uint32_t i = 0; i < b.nseq; i++) g_far_records += seqs[b.seq_base + i] >> 63;
        records_serial(blocks.data(), b, frames[b.frame].block_max, seqs.data(), tab, h);
    }
    // offsets
    std::vector<ItemOut> outs(n);
    for (size_t i = 0; i < n; i++) offsets_item(items[i], infos[i], bases[i], frames.data(), blocks.data(), outs[i]);
    // execute
    for (uint32_t f = 0; f < run.frame; f++) {
        if (outs[frames[f].item].fail) continue;
        exec_frame_serial(frames[f], blocks.data(), items[frames[f].item], seqs.data());
    }
    // checksum
    if (!(flags & FZG_NO_VERIFY_CHECKSUM))
        for (uint32_t f = 0; f < run.frame; f++) {
            Frame& fr = frames[f];
            if (!fr.has_checksum || fr.status || outs[fr.item].fail) continue;
            const uint8_t* p = items[fr.item].dst + fr.out_off;
            uint64_t v[4];
            for (uint32_t j = 0; j < 4; j++) v[j] = xx_lane(p, fr.out_size, j);
            if ((uint32_t)xx_combine(v[0], v[1], v[2], v[3], p, fr.out_size) != fr.checksum) fr.status = FZG_E_CHECKSUM;
        }
    // finish
    for (size_t i = 0; i < n; i++) {
        finish_item(infos[i], bases[i], frames.data(), outs[i]);
        dst_len[i] = outs[i].dst_len; status[i] = outs[i].status;
    }
    return 0;
}

/* intermediate stages for one single-frame item: records + literals, for stage-level comparison with
 * the oracle trace */
extern "C" int fze_trace(const void* src, size_t src_len, uint64_t* seq_out, size_t seq_cap, size_t* n_seq,
                         uint8_t* lit_out, size_t lit_cap, size_t* n_lit)
{
    Item it{ (const uint8_t*)src, src_len, nullptr, 0 };
    ItemInfo info; ItemBase base{ 0, 0, 0, 0, 0, 0 };
    walk_item<false>(0, it, info, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (info.walk_status) return info.walk_status;
    std::vector<Frame> frames(info.n_frames + 1); std::vector<Block> blocks(info.n_blocks + 1);
    std::vector<uint32_t> sj(info.n_seq_jobs + 1), hj(info.n_huf_jobs + 1);
    std::vector<uint8_t> lit(info.lit_bytes + 64); std::vector<uint64_t> seqs(info.n_seq + 8);
    ItemInfo tmp;
    walk_item<true>(0, it, tmp, &base, frames.data(), blocks.data(), sj.data(), hj.data(), lit.data());
    std::vector<uint16_t> table(1 << kHufLogMax);
    alignas(16) static uint8_t ring[256]; static LitScratch sc;
    for (uint32_t j = 0; j < info.n_huf_jobs; j++) {
        Block& b = blocks[hj[j]]; int log = 0; uint32_t used = 0;
        if (b.lit_type >= LT_HUF) lit_build(blocks.data(), b, table.data(), sc, kHufLogMax, log, used);
        for (uint32_t sub = 0; sub < 4; sub++) {
            const LitWork wk = lit_plan(b, sub, log, used);
            if (lit_run(wk, sub, table.data(), log, ring, wk.n_out / 4, 1)) return FZG_E_CORRUPT;
        }
    }
    alignas(16) static uint8_t chain_mem[kChainBytes];
    for (uint32_t j = 0; j < info.n_seq_jobs; j++) {
        Block& b = blocks[sj[j]];
        alignas(16) static uint8_t tab[kJobTableBytes]; SeqJobHdr h;
        static uint8_t w_sym[512]; static int16_t w_norm[64]; static uint16_t w_cnt[64];
        const TabWork tw{ { w_sym, 1 }, { w_norm, 1 }, { w_cnt, 1 } };
        seq_tables_thread(blocks.data(), b, kConsts, tab, h, tw);
        if (h.bad && !b.status) b.status = FZG_E_CORRUPT;
        seq_chain_thread(b, tab, h, chain_mem, seqs.data(), b.nseq - 1, 1);
        records_serial(blocks.data(), b, frames[b.frame].block_max, seqs.data(), tab, h);
    }
    it.dst_cap = ~0ull;
    ItemOut io; offsets_item(it, info, base, frames.data(), blocks.data(), io);      // resolves every block's starting history
    size_t ns = 0, nl = 0;
    for (uint32_t k = 0; k < info.n_blocks; k++) {
        const Block& b = blocks[k];
        if (b.status) return b.status;
        if (b.type != BT_COMPRESSED) continue;
        uint32_t S = 0, LEp = 0;
        for (uint32_t i = 0; i < b.nseq; i++, ns++) {           // (ll : 17 | ml : 18 | resolved distance : 29)
            const uint64_t r = seqs[b.seq_base + i];
            const uint32_t ll = rec_le(r) - LEp, ml = rec_e(r) - S - ll;
            const uint32_t off = off_resolve(rec_off(r), b.rep_in[0], b.rep_in[1], b.rep_in[2]);
            if (ns < seq_cap) seq_out[ns] = (uint64_t)ll | ((uint64_t)ml << 17) | ((uint64_t)off << 35);
            S = rec_e(r); LEp = rec_le(r);
        }
        for (uint32_t i = 0; i < b.lit_regen; i++, nl++) if (nl < lit_cap) lit_out[nl] = b.lit[i];
    }
    *n_seq = ns; *n_lit = nl;
    return 0;
}

/* Encoder building blocks against the decoder's: normalise the histogram of syms[0..n), write the FSE table description,
 * FSE-encode the symbols with the encoding table, then read the description back (read_ncount), rebuild the DECODING
 * table (build_fse_table) and decode.  Returns 0 when everything round-trips, a negative stage code otherwise. */
extern "C" int fze_fse_roundtrip(const uint8_t* syms, size_t n, int n_sym, int max_log)
{
    std::vector<uint32_t> count(n_sym, 0);
    int present = 0;
    for (size_t i = 0; i < n; i++) { if (syms[i] >= n_sym) return -100; if (!count[syms[i]]++) present++; }
    const int log = enc_table_log((uint32_t)n, present, max_log);
    std::vector<int16_t> norm(64, 0);
    if (enc_normalize(count.data(), n_sym, (uint32_t)n, log, norm.data()) != 0) return -1;
    int sum = 0; for (int s = 0; s < n_sym; s++) { if (count[s] && norm[s] < 1) return -2; sum += norm[s]; }
    if (sum != (1 << log)) return -3;
    uint8_t desc[160] = { 0 };
    const int dlen = enc_write_ncount(desc, norm.data(), n_sym, log);
    if (dlen < 0) return -4;
    int16_t back[64]; int ns2 = 0, log2 = 0;
    const int used = read_ncount(desc, (uint32_t)dlen, n_sym - 1, max_log, back, ns2, log2);
    if (used != dlen || log2 != log) return -5;
    for (int s = 0; s < n_sym; s++) if ((s < ns2 ? back[s] : 0) != norm[s]) return -6;
    // encode (single state, symbols last to first, like the sequence coder does per state)
    std::vector<uint16_t> state(1 << log); std::vector<uint32_t> dnb(n_sym); std::vector<int32_t> dfs(n_sym);
    std::vector<uint8_t> tmp(1 << log); std::vector<uint16_t> cumul(n_sym + 2);
    enc_build_ctable(state.data(), dnb.data(), dfs.data(), norm.data(), n_sym, log, tmp.data(), cumul.data());
    std::vector<uint8_t> bs(n * 2 + 16, 0); uint64_t acc = 0; uint32_t nb_acc = 0; size_t bp = 0;
    auto add = [&](uint32_t v, uint32_t nb) { acc |= (uint64_t)v << nb_acc; nb_acc += nb; while (nb_acc >= 8) { bs[bp++] = (uint8_t)acc; acc >>= 8; nb_acc -= 8; } };
    uint32_t st;
    { const uint32_t sy = syms[n - 1]; const uint32_t nbo = (dnb[sy] + (1u << 15)) >> 16; const uint32_t v = (nbo << 16) - dnb[sy]; st = state[(v >> nbo) + dfs[sy]]; }
    for (size_t i = n - 1; i-- > 0;) { const uint32_t sy = syms[i]; const uint32_t nbo = (st + dnb[sy]) >> 16; add(st & ((1u << nbo) - 1), nbo); st = state[(st >> nbo) + dfs[sy]]; }
    add(st - (1u << log), (uint32_t)log); add(1, 1); if (nb_acc) { bs[bp++] = (uint8_t)acc; }
    // decode
    std::vector<uint32_t> dt(1 << log); uint16_t cnt[64]; static const uint8_t no_extra[64] = { 0 };
    if (build_fse_table(dt.data(), back, ns2, log2, no_extra, cnt) != 0) return -7;
    BackBits br; if (br.init(bs.data(), (uint32_t)bp) != 0) return -8;
    br.refill(); uint32_t ds = br.read((uint32_t)log);
    for (size_t i = 0; i < n; i++) {
        if (br.avail <= 32) br.refill();
        const uint32_t c = dt[ds];
        if (cell_sym(c) != syms[i]) return -9;
        if (i + 1 < n) ds = cell_base(c) + br.read(cell_nb(c));
    }
    return br.left == 0 ? 0 : -10;
}
