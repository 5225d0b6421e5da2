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
 *     ~600 frames in flight, i.e. the last ~200 KB of every frame stays in the 126 MB L2 (round 1 kept 4 736 frames
 *     in flight, one per warp: 26 KB of L2 each, and fetched almost every match from DRAM: 123 GB per 14 GB).
 *
 * A thread owns one 16-byte-aligned GRANULE of the round.  It finds the first sequence that reaches into the granule
 * (binary search over the E fields of the positional records), then walks the sequences forward, piece by piece:
 * a piece is <= 8 bytes of one literal run or one match that fall inside the granule.  A piece is an unaligned
 * 8-byte load (two aligned loads + funnel shift) merged into the granule's four registers; a finished granule leaves
 * with one 16-byte store.  No byte stores, no shared-memory stage: the output buffer (L1 / L2) is the window, for
 * sources inside the round as well.
 *
 * A match piece copies from `off` bytes back.  Zstandard's overlap rule (offset < length: the copy reads what it
 * has just written) is periodicity: x[p] = x[p - k off] for as long as p - (k - 1) off >= M (M = start of the match), so
 * every piece is re-aimed at the last period BEFORE the match, [M - off, M) -- a piece never depends on its own
 * match.  Source below the round: plain load (the previous rounds are behind a barrier).  Source inside the round
 * (9-14 % of the pieces at 2-4 KB rounds): the piece becomes a HOLE (source position, length, place in the granule),
 * the walk goes on, and the holes are filled in later PASSES of the round, separated by __syncthreads_or: a granule
 * publishes the number of the pass it was stored in, a hole accepts sources stored in EARLIER passes only (so the
 * barrier orders the global stores and the loads, no fences).  The lowest unfinished granule can always finish
 * (everything below it is stored), so a round takes at most T passes; ordinary text takes 2-3.
 *
 * The positional records reach shared memory through TMA bulk copies (cp.async.bulk + mbarrier, UBLKCP in SASS):
 * chunks of 256 records, a ring of 8+ chunks, issued two rounds ahead by thread 0, so a round never waits for its
 * records.  Record index nseq of a block is the TAIL record written by k_records (E = block size, LE = all
 * literals): the literals after the last sequence are one more literal run and need no code of their own.
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

constexpr uint32_t kTileChunk = 256;                        // records per bulk copy (2 KB)

template <int T> struct TileCfg {
    static constexpr uint32_t threads = T;
    static constexpr uint32_t round_bytes = 16u * T;
    static constexpr uint32_t rmax = 2u * T;                                     // records a round may use (>= 3 bytes each)
    static constexpr uint32_t need = 2u * rmax / kTileChunk + 3u;               // chunks alive at once: [g - 1, g + 2 rmax]
    static constexpr uint32_t slots = need <= 8 ? 8 : (need <= 16 ? 16 : 32);   // power of two
    static constexpr uint32_t ring = slots * kTileChunk;                         // records in the ring
    static constexpr uint32_t window = 128u * T;                                 // bytes of the frame's most recent output kept in shared memory (power of two)
    static constexpr uint32_t smem = ring * 8u + window;                         // dynamic shared memory: record ring | window
    static constexpr int ctas_per_sm = T <= 128 ? 8 : (T <= 256 ? 4 : (T <= 512 ? 2 : 1));
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
// exact-size L2 prefetch (16-byte units): the next round's far match sources
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }

// nb (1..8) bytes at any alignment.  The bytes live either in global memory (the far window, L2 / HBM) or in the CTA's
// shared-memory window, a ring indexed by the low bits of the global address -- so the two aligned 8-byte words that hold
// them are given as two generic pointers (the second word of a ring source may wrap).
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
// GLOBAL address of an output byte (granules are 16-byte aligned in both).  Scattered loads are what bounds this stage: a
// warp-wide load of 32 random global addresses occupies the L1 pipeline for 32-64 cycles, the same load from shared
// memory for 2-4 (bank conflicts only), and 50-75 % of the match sources lie within the last 32-128 KB.
struct TileWin {
    uint8_t* p;               // generic pointer to the ring
    uint32_t mask;            // window - 1
    __device__ __forceinline__ uint8_t* at(const uint8_t* g) const { return p + ((uint32_t)(uintptr_t)g & mask); }
};
// n (1..8) bytes of the frame at global address g: from the ring (in_win) or from global memory
__device__ __forceinline__ uint64_t tile_load(const TileWin& w, const uint8_t* g, uint32_t n, bool in_win)
{
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7);
    const uint8_t* a = g - sh;
    const void* p0 = in_win ? (const void*)w.at(a) : (const void*)a;
    const void* p1 = in_win ? (const void*)w.at(a + 8) : (const void*)(a + 8);
    return ld8_pair(p0, p1, sh, n);
}

// A hole whose source reaches into the thread's own granule (distances below ~24: rare): byte by byte, own bytes
// from the registers (if no earlier hole still covers them), the others from the ring if their granule is stored.
// Returns false when a byte is not there yet.
__device__ __noinline__ bool tile_fill_bytes(TileOut& o, int32_t s, uint32_t n, uint32_t d, int32_t P, int32_t base, uint32_t pass,
                                             const uint16_t* flag, const uint8_t* g0, const TileWin& w, int32_t safe_lo,
                                             uint32_t unfilled /* bytes of the granule still covered by holes */)
{
    uint64_t v = 0;
    for (uint32_t k = 0; k < n; k++) {
        const int32_t q = s + (int32_t)k;
        uint32_t byte;
        if (q >= P) {
            const uint32_t at = (uint32_t)(q - P);
            if ((unfilled >> at) & 1u) return false;
            byte = tile_byte(o, at);
        } else {
            const int32_t j = (q - base) >> 4;
            if (j >= 0) { const uint32_t f = flag[j]; if (f == 0 || f >= pass) return false; }
            byte = q >= safe_lo ? *(volatile const uint8_t*)w.at(g0 + q) : *(volatile const uint8_t*)(g0 + q);
        }
        v |= (uint64_t)byte << (8 * k);
    }
    tile_merge(o, v, n, d);
    return true;
}

// One compressed block with sequences, by the whole CTA.  g0 = the block's first output byte, done = bytes of the frame
// before it, cg = running chunk number of the CTA's record ring (see k_execute_tile), win_from = block-relative position
// from which on the shared-memory window holds the frame's bytes (<= 0: since some earlier block).
template <int T>
__device__ __forceinline__ void tile_block(uint64_t* ring, const uint32_t ring_sm, const uint32_t bar_sm, uint16_t* flag, uint32_t* s_next /*[2]*/,
                                           const TileWin& win, int32_t win_from, const Block& b, const uint64_t* __restrict__ sq, uint8_t* g0,
                                           uint64_t done, uint32_t& cg, int& status, const uint32_t tid)
{
    using C = TileCfg<T>;
    const uint32_t nseq = b.nseq, rsize = b.rsize, nrec = nseq + 1;
    const uint8_t* __restrict__ lit = b.lit;
    const uint32_t in0 = b.rep_in[0], in1 = b.rep_in[1], in2 = b.rep_in[2];
    const uint32_t n_chunks = (nrec + kTileChunk - 1) / kTileChunk;
    const uint32_t rbase = cg * kTileChunk;                              // ring position of record 0
    const uint32_t reach = (uint32_t)(done < (1ull << 28) ? done : (1ull << 28));   // off <= 2^27: anything beyond is always inside the frame
    const uint32_t a0 = (uint32_t)((uintptr_t)g0 & 15);
    auto rec = [&](uint32_t idx) -> uint2 { return *(const uint2*)(ring + ((rbase + idx) & (C::ring - 1))); };
    uint32_t issued = 0;                                                 // chunks of this block in flight or landed (CTA-uniform)
    uint32_t g = 0, gS = 0;                                              // first record with E > gS; start of the round (CTA-uniform)
    uint32_t par = 0;                                                    // s_next is double-buffered: a fast thread may finish the next round's walk before a slow one has read this round's
    while (gS < rsize) {
        // ---- records: chunks [c_lo, c_need] are read by this round, everything up to c_top is asked for now
        const uint32_t c_lo = g ? (g - 1) / kTileChunk : 0;
        const uint32_t c_need = min(n_chunks - 1, (g + C::rmax) / kTileChunk);
        const uint32_t c_top = min(min(n_chunks - 1, (g + 2 * C::rmax) / kTileChunk), c_lo + C::slots - 1);
        if (tid == 0) {
            for (uint32_t c = issued; c <= c_top; c++) {
                const uint32_t slot = (cg + c) & (C::slots - 1);
                const uint32_t n = min(kTileChunk, nrec - c * kTileChunk);
                const uint32_t bytes = ((n + 1) & ~1u) * 8;              // 16-byte units: the pad record exists (walk_item)
                mbar_expect_tx(bar_sm + 8 * slot, bytes);
                bulk_g2s(ring_sm + slot * kTileChunk * 8, sq + (size_t)c * kTileChunk, bytes, bar_sm + 8 * slot);
            }
        }
        issued = max(issued, c_top + 1);
        for (uint32_t c = c_lo; c <= c_need; c++) mbar_wait(bar_sm + 8 * ((cg + c) & (C::slots - 1)), ((cg + c) / C::slots) & 1u);
        // ---- geometry of the round: granules are 16-byte aligned in memory; it ends on a granule boundary or at the block's end
        const uint32_t last = min(g + C::rmax - 1, nseq);
        const uint32_t a = (a0 + gS) & 15;
        const int32_t base = (int32_t)gS - (int32_t)a;                   // position of granule 0's first byte
        uint32_t gE = min((uint32_t)(base + (int32_t)C::round_bytes), rec_e(rec(last).x));
        if (gE < rsize) gE -= (a0 + gE) & 15;
        // the ring slots of positions below base + round_bytes - window are overwritten during this round
        const int32_t safe_lo = max(win_from, base + (int32_t)C::round_bytes - (int32_t)C::window);
        const int32_t P = base + 16 * (int32_t)tid;
        const uint32_t lo = (uint32_t)max(P, (int32_t)gS), hi = (uint32_t)min(P + 16, (int32_t)gE);
        const bool active = P + 16 > (int32_t)gS && P < (int32_t)gE;
        flag[tid] = 0;
        // ---- the first sequence that reaches beyond lo
        uint32_t i = g;
        {
            uint32_t cnt = 0;
#pragma unroll
            for (uint32_t step = C::rmax / 2; step; step >>= 1) {
                const uint32_t c = cnt + step;
                if (g + c - 1 <= last && rec_e(rec(g + c - 1).x) <= lo) cnt = c;
            }
            i = g + cnt;
        }
        uint32_t Eprev = 0, LEprev = 0;
        if (i) { const uint2 r = rec(i - 1); const uint64_t q = (uint64_t)r.x | ((uint64_t)r.y << 32); Eprev = rec_e(q); LEprev = rec_le(q); }
        uint32_t cur = lo;
        TileOut o{ 0, 0 };
        int32_t hs0 = 0, hs1 = 0, hs2 = 0; uint32_t hm0 = 0, hm1 = 0, hm2 = 0;      // holes: source position, (d | n << 4), 0: free
        bool pending = active;
        for (uint32_t pass = 1;; pass++) {
            if (pending) {
                // ---- holes of the earlier passes: their sources lie inside the round, i.e. in the ring once stored
                if (hm0 | hm1 | hm2) {
                    auto fill = [&](int32_t s, uint32_t& m) {
                        if (!m) return;
                        const uint32_t d = m & 15u, n = m >> 4;
                        if (s + (int32_t)n > P) {                             // reaches into this very granule
                            uint32_t unfilled = 0;
                            if (hm0) unfilled |= ((1u << (hm0 >> 4)) - 1u) << (hm0 & 15u);
                            if (hm1) unfilled |= ((1u << (hm1 >> 4)) - 1u) << (hm1 & 15u);
                            if (hm2) unfilled |= ((1u << (hm2 >> 4)) - 1u) << (hm2 & 15u);
                            if (tile_fill_bytes(o, s, n, d, P, base, pass, flag, g0, win, safe_lo, unfilled)) m = 0;
                            return;
                        }
                        const int32_t j0 = (s - base) >> 4, j1 = (s + (int32_t)n - 1 - base) >> 4;
                        bool ok = true;
                        if (j0 >= 0) { const uint32_t f = flag[j0]; ok = f != 0 && f < pass; }
                        if (j1 >= 0 && j1 != j0) { const uint32_t f = flag[j1]; ok = ok && f != 0 && f < pass; }
                        if (ok) { tile_merge(o, tile_load(win, g0 + s, n, s >= safe_lo), n, d); m = 0; }
                    };
                    fill(hs0, hm0); fill(hs1, hm1); fill(hs2, hm2);
                }
                // ---- the walk
                while (cur < hi) {
                    const uint2 rr = rec(i);
                    const uint64_t r = (uint64_t)rr.x | ((uint64_t)rr.y << 32);
                    const uint32_t E = rec_e(r), LE = rec_le(r);
                    const uint32_t M = Eprev + (LE - LEprev);
                    const uint32_t d = cur - (uint32_t)P;
                    uint32_t n;
                    if (cur < M) {                                            // literal run
                        n = min(min(M, hi) - cur, 8u);
                        const uint8_t* src = lit + LEprev + (cur - Eprev);
                        const uintptr_t al = (uintptr_t)src & ~(uintptr_t)7;
                        const uint32_t sh = (uint32_t)((uintptr_t)src & 7);
                        const uint2 w0 = __ldg((const uint2*)al);
                        uint2 w1 = make_uint2(0, 0);
                        if (sh + n > 8) w1 = __ldg((const uint2*)(al + 8));
                        tile_merge(o, funnel8(w0.x, w0.y, w1.x, w1.y, sh), n, d);
                    } else {                                                  // match
                        n = min(min(E, hi) - cur, 8u);
                        uint32_t off = off_resolve(rec_off(r), in0, in1, in2);
                        if (off == 0 || off > reach + M) { off = 0; status = FZG_E_CORRUPT; }      // reaches before the frame start
                        if (off) {
                            // the last period before the match: s in [M - off, M), the piece ends at M at the latest
                            const uint32_t into = cur - M;
                            const uint32_t k = into < off ? 1u : into / off + 1u;
                            const int32_t s = (int32_t)cur - (int32_t)(k * off);
                            n = min(n, (uint32_t)((int32_t)M - s));
                            if (s + (int32_t)n <= (int32_t)gS) tile_merge(o, tile_load(win, g0 + s, n, s >= safe_lo), n, d);   // below the round
                            else {
                                const uint32_t m = d | (n << 4);
                                if (!hm0) { hs0 = s; hm0 = m; } else if (!hm1) { hs1 = s; hm1 = m; } else if (!hm2) { hs2 = s; hm2 = m; }
                                else break;                                   // no free slot: the walk waits for a pass
                            }
                        }                                                     // corrupt: zeros
                    }
                    cur += n;
                    if (cur == E) { i++; Eprev = E; LEprev = LE; }
                }
                if (cur == hi && !(hm0 | hm1 | hm2)) {
                    uint8_t* gp = g0 + P; uint8_t* wp = win.at(gp);
                    if (hi - lo == 16) {
                        const uint4 v = make_uint4((uint32_t)o.lo, (uint32_t)(o.lo >> 32), (uint32_t)o.hi, (uint32_t)(o.hi >> 32));
                        *(uint4*)gp = v; *(uint4*)wp = v;
                    } else for (uint32_t k = lo - (uint32_t)P; k < hi - (uint32_t)P; k++) { const uint8_t v = (uint8_t)tile_byte(o, k); gp[k] = v; wp[k] = v; }   // a block's first / last granule
                    flag[tid] = (uint16_t)pass;
                    if (hi == gE) s_next[par] = i;
                    pending = false;
                }
            }
            if (!__syncthreads_or(pending)) break;
        }
        g = s_next[par]; gS = gE; par ^= 1u;
    }
    cg += n_chunks;
}

template <int T>
__global__ void __launch_bounds__(T, TileCfg<T>::ctas_per_sm) k_execute_tile(Frame* frames, const Block* blocks, const Item* items, const ItemOut* outs,
                                                                            const uint64_t* seqs, uint32_t n_frames, uint32_t* ticket)
{
    using C = TileCfg<T>;
    extern __shared__ __align__(128) uint64_t tile_ring[];             // record ring (C::ring records) | window (C::window bytes)
    __shared__ __align__(8) uint64_t s_bar[C::slots];
    __shared__ uint16_t s_flag[T];
    __shared__ uint32_t s_next[2], s_f;
    const uint32_t tid = threadIdx.x;
    const uint32_t ring_sm = (uint32_t)__cvta_generic_to_shared(tile_ring), bar_sm = (uint32_t)__cvta_generic_to_shared(s_bar);
    const TileWin win{ (uint8_t*)(tile_ring + C::ring), C::window - 1 };
    if (tid == 0) {
        for (uint32_t s = 0; s < C::slots; s++) mbar_init(bar_sm + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
                tile_block<T>(tile_ring, ring_sm, bar_sm, s_flag, s_next, win, win_from, b, seqs + b.seq_base, g0, done, cg, status, tid);
            }
            __syncthreads();                           // later blocks read this one back (the window)
            done += rsize;
            if (!seqs_block) win_done = done;          // written to global memory only
        }
        if (__syncthreads_or(status) && tid == 0) fr.status = FZG_E_CORRUPT;
    }
}

}  // namespace fz
