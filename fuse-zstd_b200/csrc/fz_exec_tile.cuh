/*
 * fz_exec_tile.cuh -- the LZ77 execute stage, output-centric: k_execute_tile<T>.
 *
 * Replaces the inner copy loop of zstd::stream::copy_decode (/root/reference/src/main.rs:463-467): literal runs and
 * matches of every sequence -> the plain bytes of the file.
 *
 * Why this shape (measured on the config-2 corpus, profiles/r02_match_histogram.json): matches are short (mean 8.9
 * bytes, 99.9 % <= 25) and far (median distance 32 KB; only 4.7 % below 512 bytes, 21 % below 4 KB, 79 % below
 * 128 KB), 76 % of the sequences carry no literals.  So
 *   - almost nothing inside a few KB of output depends on anything else inside it: a CTA can produce a ROUND of T x 16
 *     output bytes with all its threads at once, no polling, no per-byte bookkeeping;
 *   - what a frame needs from the memory system is its own recent output.  A CTA per frame and four CTAs per SM keep
 *     600-1 000 frames in flight instead of 4 736 (k_execute: one per warp, 15 % L2 hit rate on its window reads,
 *     123 GB of DRAM traffic per 14 GB of algorithmic bytes).
 *
 * A thread owns one 16-byte-aligned GRANULE of the round.  Per round the CTA first expands the round's records (at most 2 T,
 * two per thread) into execute-friendly 16-byte entries { M, E, literal source - S, resolved distance } and MARKS, for every
 * record, the first granule whose first byte it produces; a running maximum over the marks (warp shuffles + one barrier)
 * gives every granule the sequence it starts in -- no search.  The thread then walks the sequences forward, piece by piece: a
 * piece is <= 8 bytes of one literal run or one match that fall inside the granule, an unaligned 8-byte load (two aligned
 * loads + funnel shift) merged into the granule's four registers; a finished granule leaves with one 16-byte store to HBM and
 * one to the WINDOW: the frame's most recent 16-64 KB live in a shared-memory ring indexed by the low bits of the global
 * address, so that a near match source is an LDS (bank conflicts only) instead of a scattered global load.
 *
 * A match piece copies from `off` bytes back.  Zstandard's overlap rule (offset < length: the copy reads what it
 * has just written) is periodicity: x[p] = x[p - k off] for as long as p - (k - 1) off >= M (M = start of the match), so
 * every piece is re-aimed at the last period BEFORE the match, [M - off, M) -- a piece never depends on its own
 * match.  Source below the round: window or HBM at once (the previous rounds are behind a barrier).  Source inside the
 * round (9-14 % of the pieces at 2-4 KB rounds): the piece becomes a HOLE (source position, length, place in the granule),
 * the walk goes on, and the holes are filled in later PASSES of the round, separated by __syncthreads_or: after every pass
 * a thread publishes its (partial) granule in the window together with a 16-bit mask of the bytes that are there; a hole
 * is filled as soon as the masks of the PREVIOUS pass cover its source bytes (double-buffered masks: the barrier orders
 * the stores and the loads, no fences).  The lowest unfinished granule can always finish, so a round takes at most T
 * passes; ordinary text takes 3-4.
 *
 * The positional records reach shared memory through TMA bulk copies (cp.async.bulk + mbarrier, UBLKCP.S.G in SASS):
 * chunks of T / 2 records, a ring of 16 chunks, issued two rounds ahead by thread 0, so a round never waits for its
 * records.  Record index nseq of a block is the TAIL record written by k_records (E = block size, LE = all
 * literals): the literals after the last sequence are one more literal run and need no code of their own.
 *
 * Measured (profiles/r02_notes.md section 2): 4.9-6.5 MB of DRAM reads per 1 MiB file against 11.7 for k_execute, but
 * 1 468 warp-instructions per 512 bytes of output against 918 -- 45 ms for config 2 against 27.  Selectable
 * (FZG_EXEC_W=t128 .. t1024) and covered by the decode tests, not the default executor.
 */
#pragma once
#include "fz_kernels.cuh"

namespace fz {

__device__ __forceinline__ uint64_t funnel8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t byte_shift)
{
    if (byte_shift >= 4) { x0 = x1; x1 = x2; x2 = x3; }
    const uint32_t r = (byte_shift & 3) * 8;
    const uint32_t lo = __funnelshift_r(x0, x1, r), hi = __funnelshift_r(x1, x2, r);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
// CTA-cooperative copy, any alignment, any size
__device__ __forceinline__ void group_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t tid, uint32_t nthr)
{
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const uint32_t nv = n >> 4;
        for (uint32_t i = tid; i < nv; i += nthr) ((uint4*)dst)[i] = ((const uint4*)src)[i];
        for (uint32_t i = (nv << 4) + tid; i < n; i += nthr) dst[i] = src[i];
    } else {
        for (uint32_t i = tid; i < n; i += nthr) dst[i] = src[i];
    }
}

template <int T> struct TileCfg {
    static constexpr uint32_t threads = T;
    static constexpr uint32_t round_bytes = 16u * T;
    static constexpr uint32_t rmax = 2u * T;                                     // records a round may use (>= 3 bytes each)
    static constexpr uint32_t chunk = T / 2;                                     // records per bulk copy
    static constexpr uint32_t slots = 16;                                        // chunks in the ring: [g - 1, g + 2 rmax] is 8 chunks + 3
    static constexpr uint32_t ring = slots * chunk;                              // records in the ring (8 T)
    static constexpr uint32_t window = T <= 128 ? 16384u : (T <= 512 ? 32768u : 65536u);               // bytes of the frame's most recent output kept in shared memory (power of two)
    // dynamic shared memory: record ring | expanded records of the round | window
    static constexpr uint32_t xr_off = ring * 8u, win_off = xr_off + (rmax + 2u) * 16u, smem = win_off + window;
    static constexpr int ctas_per_sm = (int)((227u * 1024u) / (smem + 2048u + 6u * T)) < (2048 / T) ? (int)((227u * 1024u) / (smem + 2048u + 6u * T)) : (2048 / T);
};

// ---- mbarrier / bulk copy (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

// nb (1..8) bytes at any alignment, given the two aligned 8-byte words that hold them as GENERIC pointers: the bytes live in
// global memory (literals; the far window in L2 / HBM) or in the CTA's shared-memory window, whose second word may wrap.
__device__ __forceinline__ uint64_t ld8_pair(const void* p0, const void* p1, uint32_t sh, uint32_t nb)
{
    uint32_t x0, x1, x2 = 0, x3 = 0;
    asm volatile("ld.v2.u32 {%0, %1}, [%2];" : "=r"(x0), "=r"(x1) : "l"(p0) : "memory");
    if (sh + nb > 8) asm volatile("ld.v2.u32 {%0, %1}, [%2];" : "=r"(x2), "=r"(x3) : "l"(p1) : "memory");
    return funnel8(x0, x1, x2, x3, sh);
}

struct TileOut { uint64_t lo, hi; };
// the n (1..8) low bytes of v -> bytes [d, d + n) of the granule (d + n <= 16)
__device__ __forceinline__ void tile_merge(TileOut& o, uint64_t v, uint32_t n, uint32_t d)
{
    if (n < 8) v &= (1ull << (8 * n)) - 1ull;
    if (d < 8) { o.lo |= v << (8 * d); if (d) o.hi |= v >> (64 - 8 * d); }
    else o.hi |= v << (8 * (d - 8));
}
__device__ __forceinline__ uint32_t tile_byte(const TileOut& o, uint32_t k) { return (uint32_t)((k < 8 ? o.lo >> (8 * k) : o.hi >> (8 * (k - 8))) & 0xFF); }

// The CTA's shared-memory window: the most recent `window` bytes of the frame, a ring indexed by the low bits of the
// GLOBAL address of an output byte (granules are 16-byte aligned in both).
struct TileWin {
    uint8_t* p;               // generic pointer to the ring
    uint32_t mask;            // window - 1
    __device__ __forceinline__ uint8_t* at(const uint8_t* g) const { return p + ((uint32_t)(uintptr_t)g & mask); }
};

// One compressed block with sequences, by the whole CTA.  g0 = the block's first output byte, done = bytes of the frame
// before it, cg = running chunk number of the CTA's record ring (see k_execute_tile), win_from = block-relative position
// from which on the shared-memory window holds the frame's bytes (<= 0: since some earlier block).
struct TileSm {                                   // static shared memory of the kernel
    uint64_t* ring; uint4* xr; uint16_t* mark; uint16_t* have /*[2][T + 1]*/; uint32_t* wtot; uint32_t* next /*[2]*/;
    uint32_t ring_sm, bar_sm;
};

template <int T>
__device__ __forceinline__ void tile_block(const TileSm& sm, const TileWin& win, int32_t win_from, const Block& b, const uint64_t* __restrict__ sq,
                                           uint8_t* g0, uint64_t done, uint32_t& cg, int& status, const uint32_t tid)
{
    using C = TileCfg<T>;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t nseq = b.nseq, rsize = b.rsize, nrec = nseq + 1;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    const uint32_t n_chunks = (nrec + C::chunk - 1) / C::chunk;
    const uint32_t rbase = cg * C::chunk;                                // ring position of record 0
    const uint32_t reach = (uint32_t)(done < (1ull << 28) ? done : (1ull << 28));   // off <= 2^27: anything beyond is always inside the frame
    const uint32_t a0 = (uint32_t)((uintptr_t)g0 & 15);
    auto rec = [&](uint32_t idx) -> uint64_t { return sm.ring[(rbase + idx) & (C::ring - 1)]; };
    uint16_t* const have0 = sm.have; uint16_t* const have1 = sm.have + T + 1;
    uint32_t issued = 0;                                                 // chunks of this block in flight or landed (CTA-uniform)
    uint32_t g = 0, gS = 0;                                              // first record with E > gS; start of the round (CTA-uniform)
    uint32_t par = 0;                                                    // sm.next is double-buffered: a fast thread may finish the next round's walk before a slow one has read this round's
    while (gS < rsize) {
        // ---- records: chunks [c_lo, c_need] are read by this round, everything up to c_top is asked for now
        const uint32_t c_lo = g ? (g - 1) / C::chunk : 0;
        const uint32_t c_need = min(n_chunks - 1, (g + C::rmax) / C::chunk);
        const uint32_t c_top = min(min(n_chunks - 1, (g + 2 * C::rmax) / C::chunk), c_lo + C::slots - 1);
        if (tid == 0) {
            for (uint32_t c = issued; c <= c_top; c++) {
                const uint32_t slot = (cg + c) & (C::slots - 1);
                const uint32_t n = min(C::chunk, nrec - c * C::chunk);
                const uint32_t bytes = ((n + 1) & ~1u) * 8;              // 16-byte units: the pad record exists (walk_item)
                mbar_expect_tx(sm.bar_sm + 8 * slot, bytes);
                bulk_g2s(sm.ring_sm + slot * C::chunk * 8, sq + (size_t)c * C::chunk, bytes, sm.bar_sm + 8 * slot);
            }
        }
        issued = max(issued, c_top + 1);
        for (uint32_t c = c_lo; c <= c_need; c++) mbar_wait(sm.bar_sm + 8 * ((cg + c) & (C::slots - 1)), ((cg + c) / C::slots) & 1u);
        // ---- geometry of the round: granules are 16-byte aligned in memory; it ends on a granule boundary or at the block's end
        const uint32_t last = min(g + C::rmax - 1, nseq);
        const uint32_t a = (a0 + gS) & 15;
        const int32_t base = (int32_t)gS - (int32_t)a;                   // position of granule 0's first byte
        uint32_t gE = min((uint32_t)(base + (int32_t)C::round_bytes), rec_e(rec(last)));
        if (gE < rsize) gE -= (a0 + gE) & 15;
        // the ring slots of positions below base + round_bytes - window are overwritten during this round
        const int32_t safe_lo = max(win_from, base + (int32_t)C::round_bytes - (int32_t)C::window);
        const int32_t P = base + 16 * (int32_t)tid;
        const uint32_t lo = (uint32_t)max(P, (int32_t)gS), hi = (uint32_t)min(P + 16, (int32_t)gE);
        const bool active = P + 16 > (int32_t)gS && P < (int32_t)gE;
        // ---- the round's records, expanded once: { M, E, literal source - S, distance } and marked at the first granule they own
        sm.mark[tid] = 0; have0[tid] = 0; have1[tid] = 0;
        __syncthreads();                                                 // (also: the previous round's reads of xr / mark are done)
        const uint32_t nr = last - g + 1;
#pragma unroll
        for (uint32_t k = 0; k < 2; k++) {
            const uint32_t r = tid + k * T;
            if (r < nr) {
                const uint64_t q = rec(g + r), qp = g + r ? rec(g + r - 1) : 0ull;
                const uint32_t E = rec_e(q), Ep = rec_e(qp), LEp = rec_le(qp), M = Ep + (rec_le(q) - LEp);
                uint32_t off = off_resolve(rec_off(q), in0, in1, in2);
                if (r < nr - 1 || g + r < nseq) { if (off == 0 || off > reach + M) { off = 0; status = FZG_E_CORRUPT; } }     // reaches before the frame start (the tail record has no match)
                sm.xr[r] = make_uint4(M, E, LEp - Ep, off);
                // granules whose first byte this sequence produces: P_G in [Ep, E); record g owns granule 0 whatever its start
                const int32_t G0 = r ? (int32_t)(Ep - (uint32_t)base + 15u) >> 4 : 0, G1 = ((int32_t)(E - (uint32_t)base + 15u) >> 4) - 1;
                if (G0 <= G1 && G0 < T) sm.mark[G0] = (uint16_t)(r + 1);
            }
        }
        __syncthreads();
        // ---- the first sequence that reaches beyond the granule's first byte: running maximum of the marks
        uint32_t i;
        {
            uint32_t v = sm.mark[tid];
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, dd); if (lane >= (uint32_t)dd) v = max(v, t); }
            if (lane == 31) sm.wtot[warp] = v;
            __syncthreads();
            uint32_t carry = 0;
            for (uint32_t w = 0; w < warp; w++) carry = max(carry, sm.wtot[w]);
            i = max(v, carry) - 1;                                       // >= 0: granule 0 is marked
        }
        uint32_t cur = lo;
        TileOut o{ 0, 0 };
        uint32_t mine = 0;                                               // bytes of the granule that are in `o`
        uint32_t h0 = 0, h1 = 0, h2 = 0;                                 // holes: (s - base) [0:15) | d [15:19) | n [19:23), 0 = free slot
        const uint32_t full = active ? ((1u << (hi - (uint32_t)P)) - 1u) & ~((1u << (lo - (uint32_t)P)) - 1u) : 0u;
        bool pending = active, linger = false;
        for (uint32_t pass = 1;; pass++) {
            uint16_t* const have_w = (pass & 1) ? have1 : have0;         // written in this pass, read in the next
            const uint16_t* const have_r = (pass & 1) ? have0 : have1;
            if (linger) { have_w[tid] = (uint16_t)full; linger = false; }
            if (pending) {
                // ---- holes of the earlier passes: their sources lie inside the round, i.e. in the window once stored
                if (h0 | h1 | h2) {
#pragma unroll 1
                    for (int k = 0; k < 3; k++) {
                        uint32_t h = h0;
                        if (h) {
                            const uint32_t sb = h & 0x7FFFu, d = (h >> 15) & 15u, n = h >> 19;
                            const uint32_t hv = (uint32_t)have_r[sb >> 4] | ((uint32_t)have_r[(sb >> 4) + 1] << 16);
                            const uint32_t need = ((1u << n) - 1u) << (sb & 15u);
                            if ((hv & need) == need) {
                                const uint8_t* gp = g0 + base + (int32_t)sb;
                                const uint32_t sh = (uint32_t)((uintptr_t)gp & 7);
                                tile_merge(o, ld8_pair(win.at(gp - sh), win.at(gp - sh + 8), sh, n), n, d);
                                mine |= ((1u << n) - 1u) << d;
                                h = 0;
                            }
                        }
                        h0 = h1; h1 = h2; h2 = h;
                    }
                    // compact: free slots last
                    if (!h0) { h0 = h1; h1 = h2; h2 = 0; }
                    if (!h0) { h0 = h1; h1 = 0; }
                    if (!h1) { h1 = h2; h2 = 0; }
                }
                // ---- the walk
                while (cur < hi && !h2) {
                    const uint4 x = sm.xr[i];                                 // { M, E, literal source - S, distance }
                    const uint32_t d = cur - (uint32_t)P;
                    const bool is_lit = cur < x.x;
                    uint32_t n = min(min(is_lit ? x.x : x.y, hi) - cur, 8u);
                    const uint8_t* src = lit + (uint32_t)(x.z + cur);         // literal run: lit[LEp + (cur - S)] (x.z = LEp - S, modulo 2^32)
                    bool in_win = false, hole = false;
                    if (!is_lit) {
                        const uint32_t off = x.w;
                        // the last period before the match: s in [M - off, M), the piece ends at M at the latest
                        const uint32_t into = cur - x.x;
                        uint32_t back = off;
                        if (into >= off && off) back = (into / off + 1u) * off;
                        const int32_t s = (int32_t)cur - (int32_t)back;
                        n = min(n, x.x - (uint32_t)s);
                        if (s < (int32_t)gS && s + (int32_t)n > (int32_t)gS) n = gS - (uint32_t)s;        // a piece lies below the round or inside it
                        src = g0 + s;
                        in_win = s >= safe_lo;
                        hole = s >= (int32_t)gS;
                        if (off == 0) { hole = false; in_win = false; src = (const uint8_t*)sm.xr; n = min(min(x.y, hi) - cur, 8u); }   // corrupt (the frame fails): any readable bytes
                        if (hole) {
                            const uint32_t h = (uint32_t)(s - base) | (d << 15) | (n << 19);
                            if (!h0) h0 = h; else if (!h1) h1 = h; else h2 = h;
                        }
                    }
                    if (!hole) {
                        const uint32_t sh = (uint32_t)((uintptr_t)src & 7);
                        const uint8_t* al = src - sh;
                        const void* p0 = in_win ? (const void*)win.at(al) : (const void*)al;
                        const void* p1 = in_win ? (const void*)win.at(al + 8) : (const void*)(al + 8);
                        tile_merge(o, ld8_pair(p0, p1, sh, n), n, d);
                        mine |= ((1u << n) - 1u) << d;
                    }
                    cur += n;
                    if (cur == x.y) i++;
                }
                // ---- publish what the granule has; a complete one leaves for global memory
                uint8_t* gp = g0 + P; uint8_t* wp = win.at(gp);
                const uint4 v = make_uint4((uint32_t)o.lo, (uint32_t)(o.lo >> 32), (uint32_t)o.hi, (uint32_t)(o.hi >> 32));
                if (hi - lo == 16) *(uint4*)wp = v;
                else for (uint32_t k = lo - (uint32_t)P; k < hi - (uint32_t)P; k++) wp[k] = (uint8_t)tile_byte(o, k);          // a block's first / last granule
                have_w[tid] = (uint16_t)mine;
                if (mine == full) {
                    if (hi - lo == 16) *(uint4*)gp = v;
                    else for (uint32_t k = lo - (uint32_t)P; k < hi - (uint32_t)P; k++) gp[k] = (uint8_t)tile_byte(o, k);
                    if (hi == gE) sm.next[par] = i;
                    pending = false; linger = true;
                }
            }
            if (!__syncthreads_or(pending)) break;
        }
        g = g + sm.next[par]; gS = gE; par ^= 1u;
    }
    cg += n_chunks;
}

template <int T>
__global__ void __launch_bounds__(T, TileCfg<T>::ctas_per_sm) k_execute_tile(Frame* frames, const Block* blocks, const Item* items, const ItemOut* outs,
                                                                            const uint64_t* seqs, uint32_t n_frames, uint32_t* ticket)
{
    using C = TileCfg<T>;
    extern __shared__ __align__(128) uint8_t tile_dyn[];               // record ring | expanded records | window
    __shared__ __align__(8) uint64_t s_bar[C::slots];
    __shared__ uint16_t s_mark[T], s_have[2 * (T + 1)];
    __shared__ uint32_t s_wtot[32], s_next[2], s_f;
    const uint32_t tid = threadIdx.x;
    TileSm sm;
    sm.ring = (uint64_t*)tile_dyn; sm.xr = (uint4*)(tile_dyn + C::xr_off); sm.mark = s_mark; sm.have = s_have; sm.wtot = s_wtot; sm.next = s_next;
    sm.ring_sm = (uint32_t)__cvta_generic_to_shared(tile_dyn); sm.bar_sm = (uint32_t)__cvta_generic_to_shared(s_bar);
    const TileWin win{ tile_dyn + C::win_off, C::window - 1 };
    if (tid == 0) {
        for (uint32_t s = 0; s < C::slots; s++) mbar_init(sm.bar_sm + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_have[T] = 0; s_have[2 * T + 1] = 0;                           // the entry a hole's second granule may read past the round
    }
    uint32_t cg = 0;                                   // chunks ever issued into the ring: slot = cg % slots, phase = cg / slots
    for (;;) {
        __syncthreads();
        if (tid == 0) s_f = atomicAdd(ticket, 1);
        __syncthreads();
        const uint32_t f = s_f;
        if (f >= n_frames) return;
        Frame& fr = frames[f];
        if (outs[fr.item].fail) continue;
        uint8_t* const fbase = items[fr.item].dst + fr.out_off;
        uint64_t done = 0, win_done = 0;               // win_done: frame position from which on the window mirrors the output
        int status = 0;
        for (uint32_t kb = 0; kb < fr.n_blocks; kb++) {
            const Block& b = blocks[fr.first_block + kb];
            uint8_t* const g0 = fbase + done;
            const uint32_t rsize = b.rsize;
            const bool seqs_block = b.type == BT_COMPRESSED && b.nseq != 0;
            if (b.type == BT_RAW) group_copy(g0, b.src, rsize, tid, T);
            else if (b.type == BT_RLE) {
                const uint8_t v = b.src[0];
                for (uint32_t i = tid; i < rsize; i += T) g0[i] = v;
            } else if (b.nseq == 0) group_copy(g0, b.lit, rsize, tid, T);
            else {
                const uint64_t back = done - win_done;                   // bytes of the frame before this block that the window mirrors
                const int32_t win_from = back < (uint64_t)C::window ? -(int32_t)back : -(int32_t)C::window;
                tile_block<T>(sm, win, win_from, b, seqs + b.seq_base, g0, done, cg, status, tid);
            }
            __syncthreads();                           // later blocks read this one back (the window)
            done += rsize;
            if (!seqs_block) win_done = done;          // written to global memory only
        }
        if (__syncthreads_or(status) && tid == 0) fr.status = FZG_E_CORRUPT;
    }
}

}  // namespace fz
