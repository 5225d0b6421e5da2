/*
 * fz_encode.cu -- sm_100a zstd encoder: replaces the Encoder flow of
 * /root/reference/src/main.rs:781-791 (Encoder::new(level) + set_pledged_src_size + include_checksum(true) +
 * io::copy + finish) for a whole batch of files per call.
 *
 * Output format: every file becomes a concatenation of INDEPENDENT frames of <= `chunk_size` bytes of input (default
 * 1 MiB), each with Frame_Content_Size, Single_Segment and the XXH64 content checksum (what the reference's writer
 * sets), holding blocks of <= 128 KiB (Compressed, or Raw when that is not smaller).  The blocks of a frame SHARE THE
 * WINDOW: a match may reach up to 64 KiB back, into the previous blocks of its frame.  The unit of parallel work is still
 * the block ("chunk"): its matcher first seeds its hash tables from the 64 KiB of the frame before it, and it spends repeat
 * codes only once its own first three offsets have defined the history (a block does not know the history the previous
 * block ends with -- blocks are matched in parallel -- and after three plain offsets it no longer matters).  The
 * reference's decoder (zstd::stream::copy_decode, /root/reference/src/main.rs:463) reads concatenated frames as one file.
 * `level` is accepted as the reference passes it (0..19, 0 => default) and selects the one strategy
 * implemented here, which sits below libzstd level 3 in ratio.
 *
 * Stages (one kernel each, all on the context's stream); unit = chunk = frame:
 *   k_enc_match    warp/chunk    greedy LZ77: 32 positions per step, 4-byte hash, 16 KB shared-memory table
 *                                per warp (+ __match_any_sync for candidates inside the step), matches verified
 *                                and extended 8 bytes at a time; emits sequences + the literal run bytes
 *   k_enc_lit      warp/chunk    literals: histogram -> length-limited (11 bit) code lengths with an exact
 *                                Kraft sum -> canonical codes in the decoder's order -> direct-weight tree
 *                                description -> 4 Huffman streams (sizes computed first, written in place)
 *   k_enc_seq      warp/chunk    sequences: histograms -> per-block FSE tables (normalised counts, table description,
 *                                encoding table; Predefined for few sequences, RLE for one symbol) -> bitstream
 *   k_enc_place    thread/item   frame sizes -> offsets inside the item's dst, capacity check
 *   k_enc_write    warp/chunk    frame header + block header + sections (or the raw bytes) + XXH64 trailer
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fz_host.h"
#include "fz_kernels.cuh"
#include "fz_enc_core.cuh"

namespace fz {

constexpr uint32_t kEncChunkMax = 128u * 1024u;
constexpr uint32_t kEncMinMatch = 4;
#ifndef FZ_ENC_HASHLOG
#define FZ_ENC_HASHLOG 13
#endif
constexpr uint32_t kEncHashLog = FZ_ENC_HASHLOG;
constexpr uint32_t kEncMaxOff = 65535;                       // the table keeps the low 16 bits of a position
#ifndef FZ_ENC_MATCH_WARPS
#define FZ_ENC_MATCH_WARPS 4
#endif
constexpr uint32_t kEncMatchWarps = FZ_ENC_MATCH_WARPS;      // warps (= chunks in flight) per CTA in k_enc_match
#ifndef FZ_ENC_LAZY
#define FZ_ENC_LAZY 1
#endif
// Matcher modes, chosen by the level the reference passes (src/main.rs:781-785; 0 = libzstd's default = 3):
//   0  levels 1-2   the shared-memory table alone (4-byte hash, 2^13 16-bit entries per warp, 64 KiB of reach)
//   1  levels 0, 3  + a SECOND table per warp in global memory (i.e. in L2), keyed by an 8-byte hash: 2^16 16-bit entries (the low
//                   bits of a frame position: 64 KiB of reach); the shared table then hashes 5 bytes
//   2  levels 4-19  the second table holds whole frame positions in 32 bits and reaches 256 KiB back, across the blocks of a frame
// Measured on 256 x 4 MiB JSON (libzstd level 3: ratio 3.015), GB/s of input / ratio / bytes against libzstd level 3:
//   mode 0: 14.9 / 2.42 / x1.246;   mode 1: 10.1 / 2.716 / x1.110;   mode 2: 7.3 / 2.762 / x1.092
//   (mode 1 with 2^15 / 2^17 entries: 10.8 / 2.682 and 9.6 / 2.736; mode 2 with 2^15 entries: 8.2 / 2.695, with 1 MiB of reach: 6.8 /
//   2.695 -- what buys ratio is slots, not reach: 65 % of libzstd's own matches lie within 64 KiB, 90 % within 256 KiB,
//   profiles/r02_match_histogram.json.  Every 2- or 4-byte table access is a 32-byte sector transaction in L2, which is what
//   bounds the stage.)
#ifndef FZ_ENC_PREFETCH
#define FZ_ENC_PREFETCH 1
#endif
template <int MODE> struct EncMode {
    static constexpr bool has_long = MODE != 0, wide = MODE == 2;
    static constexpr uint32_t long_log = 16;
    static constexpr uint32_t long_bytes = has_long ? (wide ? 4u : 2u) << long_log : 0u;      // the long table of one warp
    static constexpr uint32_t long_window = wide ? 256u * 1024u : 65535u;
    static constexpr uint32_t window = has_long ? long_window : 65535u;                      // farthest candidate of any table
    __device__ static __forceinline__ uint32_t put(uint32_t P) { return wide ? P + 1 : (P & 0xFFFFu); }
    __device__ static __forceinline__ int32_t get(uint32_t el, uint32_t P)
    {
        if (wide) return (int32_t)el - 1;                              // 0 = empty
        int32_t c = (int32_t)((P & ~0xFFFFu) | el); if (c >= (int32_t)P) c -= 65536; return c;
    }
    __device__ static __forceinline__ uint32_t load(const uint8_t* t, uint32_t h) { return wide ? ((const uint32_t*)t)[h] : (uint32_t)((const uint16_t*)t)[h]; }
    __device__ static __forceinline__ void store(uint8_t* t, uint32_t h, uint32_t P) { if (wide) ((uint32_t*)t)[h] = put(P); else ((uint16_t*)t)[h] = (uint16_t)put(P); }
};
constexpr uint32_t kHufMaxLen = 11;

struct EncChunk {
    const uint8_t* src;      // chunk input
    uint8_t* scratch;        // per-chunk scratch: literals | sequences | literals section | sequences section
    uint32_t size;           // input bytes (<= 128 KiB)
    uint32_t item;           // index of the owning item in this launch
    uint32_t nseq, nlit;     // from k_enc_match
    uint32_t lit_sec, seq_sec;   // section sizes in bytes (0 lit_sec = not compressible -> raw block)
    uint32_t frame_size;     // bytes this chunk occupies in the output
    uint32_t raw;            // 1: the block is stored Raw (set by k_enc_place)
    uint64_t out_off;        // offset of this block's bytes (frame header included for a frame's first block) inside the item's dst
    uint32_t foff;           // offset of the chunk inside its frame: src - foff is the frame's first byte
    uint32_t fsize;          // content size of the frame
    uint32_t nblk;           // blocks in the frame (meaningful in the frame's first chunk)
    uint32_t pad;
};

// scratch layout per chunk (offsets from EncChunk::scratch); sized for the worst case
constexpr uint32_t kScrLit = 0;                                          // literal bytes, <= 128 KiB
constexpr uint32_t kScrSeq = kEncChunkMax + 64;                          // 8-byte sequence records, <= size/4 + 1
constexpr uint32_t kScrLitSec = kScrSeq + (kEncChunkMax / kEncMinMatch + 8) * 8;      // literals section
constexpr uint32_t kScrSeqSec = kScrLitSec + kEncChunkMax + kEncChunkMax / 2 + 1024;  // sequences section
constexpr uint32_t kScrBytes = kScrSeqSec + kEncChunkMax + kEncChunkMax / 2 + 1024;

FZ_HD uint64_t eseq_pack(uint32_t ll, uint32_t ml, uint32_t off) { return (uint64_t)ll | ((uint64_t)ml << 18) | ((uint64_t)off << 36); }
FZ_HD uint32_t eseq_ll(uint64_t r) { return (uint32_t)r & 0x3FFFFu; }
FZ_HD uint32_t eseq_ml(uint64_t r) { return (uint32_t)(r >> 18) & 0x3FFFFu; }
FZ_HD uint32_t eseq_off(uint64_t r) { return (uint32_t)(r >> 36); }   // Offset_Value (repeat code 1..3, or distance + 3)

__device__ __forceinline__ uint64_t ld8u(const uint8_t* g)   // 8 bytes at any alignment (may read up to 15 bytes past g)
{
    const uintptr_t a = (uintptr_t)g & ~(uintptr_t)7;
    const uint32_t sh = (uint32_t)((uintptr_t)g & 7) * 8;
    const uint64_t w0 = *(const uint64_t*)a;
    if (sh == 0) return w0;
    const uint64_t w1 = *(const uint64_t*)(a + 8);
    return (w0 >> sh) | (w1 << (64 - sh));
}

// ------------------------------------------------------------------ LZ77 matching
template <int MODE>
__global__ void __launch_bounds__(kEncMatchWarps * 32) k_enc_match(EncChunk* chunks, uint32_t n_chunks, uint32_t* ticket, uint8_t* gtab)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* table = (uint16_t*)smem + warp * (1u << kEncHashLog);
    using EM = EncMode<MODE>;
    constexpr bool LONG = EM::has_long;
    uint8_t* ltable = LONG ? gtab + (size_t)(blockIdx.x * kEncMatchWarps + warp) * EM::long_bytes : nullptr;
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(ticket, 1);
        ci = __shfl_sync(0xFFFFFFFFu, ci, 0);
        if (ci >= n_chunks) return;
        EncChunk& ch = chunks[ci];
        const uint8_t* __restrict__ src = ch.src;
        const uint32_t size = ch.size, foff = ch.foff;
        const uint8_t* __restrict__ fsrc = src - foff;                     // the frame: table entries and candidates are frame positions
        uint8_t* lit = ch.scratch + kScrLit;
        uint64_t* seq = (uint64_t*)(ch.scratch + kScrSeq);
        for (uint32_t i = lane; i < (1u << kEncHashLog) / 8; i += 32) ((uint4*)table)[i] = make_uint4(0, 0, 0, 0);
        if constexpr (LONG) for (uint32_t i = lane; i < EM::long_bytes / 16; i += 32) ((uint4*)ltable)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        // ---- the window of the previous blocks: the tables are seeded with the positions of the 64 KiB before this block
        // (inserts only; when several lanes of a step hash alike any one of them stays, and every candidate is verified)
        for (uint32_t q0 = foff > EM::window ? foff - EM::window : 0; q0 < foff; q0 += 32) {
            const uint32_t q = q0 + lane;
            if (q < foff) {
                const uint64_t v = ld8u(fsrc + q);
                if constexpr (LONG) {
                    if (foff - q <= kEncMaxOff) table[(uint32_t)(((v << 24) * 0x9E3779B185EBCA87ull) >> (64 - kEncHashLog))] = (uint16_t)q;
                    EM::store(ltable, (uint32_t)((v * 0x9E3779B185EBCA87ull) >> (64 - EM::long_log)), q);
                } else table[((uint32_t)v * 2654435761u) >> (32 - kEncHashLog)] = (uint16_t)q;
            }
        }
        __syncwarp();
        uint32_t anchor = 0, cur = 0, nseq = 0, nlit = 0;
        // repeat-offset history: {1, 4, 8} at the start of a frame (RFC 8878 3.1.1.5); a later block learns it from its own first
        // three offsets (`known`) and spends no repeat code before that
        uint32_t rep0 = 1, rep1 = 4, rep2 = 8, known = foff == 0 ? 3u : 0u;
        // positions whose 8-byte probe would run past the chunk are left to the trailing literals
        const uint32_t limit = size >= 16 ? size - 12 : 0;
        // The step's 8 input bytes and (LONG) its long-table entries are loaded one step EARLY, so that neither load sits on
        // the step's dependent chain (table -> candidate bytes -> extension).  An entry read early misses the positions the
        // previous step inserts after the read: the candidate is then an older occurrence, still verified byte by byte.
        uint64_t v_nx = 0; uint32_t hl_nx = 0, el_nx = 0;
        if (lane < limit) {
            v_nx = ld8u(src + lane);
            if constexpr (LONG) { hl_nx = (uint32_t)((v_nx * 0x9E3779B185EBCA87ull) >> (64 - EM::long_log)); el_nx = EM::load(ltable, hl_nx); }
        }
        for (uint32_t base = 0; base < limit; base += 32) {
            const uint32_t p = base + lane;
            const bool in = p < limit;
            const uint64_t v = in ? v_nx : 0; const uint32_t hl = hl_nx, el = el_nx;
            uint32_t h = 0;
            if (p + 32 < limit) {
                v_nx = ld8u(src + p + 32);
                if constexpr (LONG) { hl_nx = (uint32_t)((v_nx * 0x9E3779B185EBCA87ull) >> (64 - EM::long_log)); el_nx = EM::load(ltable, hl_nx); }
            }
#if FZ_ENC_PREFETCH
            if constexpr (LONG) {                                           // this step's long candidate: its bytes are needed ~100 instructions from now
                if (in) {
                    const int32_t c = EM::get(el, foff + p);
                    if (c >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(fsrc + c));
                }
            }
#endif
            if (in) {
                if constexpr (LONG) h = (uint32_t)(((v << 24) * 0x9E3779B185EBCA87ull) >> (64 - kEncHashLog));
                else h = ((uint32_t)v * 2654435761u) >> (32 - kEncHashLog);
            }
            // candidates: the nearest earlier lane of this step with the same hash, else the table
            const uint32_t same = __match_any_sync(0xFFFFFFFFu, in ? h : (0x80000000u | lane));
            const uint32_t below = same & ((1u << lane) - 1);
            const uint32_t P = foff + p;                                    // frame position of this lane
            int32_t cand = -1;                                              // candidates are frame positions (< P), possibly in an earlier block
            if (in) {
                if (below) cand = (int32_t)(foff + base + (31 - __clz(below)));
                else {
                    const uint32_t e = table[h];
                    int32_t c = (int32_t)((P & ~0xFFFFu) | e);
                    if (c >= (int32_t)P) c -= 65536;
                    cand = c;
                }
            }
            uint32_t samel = 0; int32_t candl = -1;
            if constexpr (LONG) {                                           // the last position with the same 8 bytes (hashed)
                samel = __match_any_sync(0xFFFFFFFFu, in ? hl : (0x80000000u | lane));
                const uint32_t belowl = samel & ((1u << lane) - 1);
                if (in && p >= cur) {
                    if (belowl) candl = (int32_t)(foff + base + (31 - __clz(belowl)));
                    else candl = EM::get(el, P);
                }
            }
            __syncwarp();
            if (in && (same >> lane) == 1) table[h] = (uint16_t)P;          // the highest lane of a group records it
            if constexpr (LONG) {
#ifndef FZ_ENC_LONG_STRIDE
#define FZ_ENC_LONG_STRIDE 1
#endif
                if (in && (samel >> lane) == 1 && (FZ_ENC_LONG_STRIDE == 1 || (P & (FZ_ENC_LONG_STRIDE - 1)) == 0)) EM::store(ltable, hl, P);
                // keep the candidate that shares the longer prefix of the first 8 bytes (the nearer one on a tie)
                if (in && p >= cur) {
                    const bool okl = candl >= 0 && P - (uint32_t)candl <= EM::long_window, oks = cand >= 0 && P - (uint32_t)cand <= kEncMaxOff;
                    const uint64_t xl = okl ? ld8u(fsrc + candl) ^ v : 1ull, xs = oks ? ld8u(fsrc + cand) ^ v : 1ull;
                    const uint32_t pl = xl ? (uint32_t)(__ffsll((long long)xl) - 1) >> 3 : 8u, ps = xs ? (uint32_t)(__ffsll((long long)xs) - 1) >> 3 : 8u;
                    if (!oks || (okl && (pl > ps || (pl == ps && candl > cand)))) cand = okl ? candl : -1;
                }
            }
            // verify + extend (8 bytes per probe)
            uint32_t len = 0;
            if (in && p >= cur && cand >= 0 && P - (uint32_t)cand <= EM::window) {
                const uint8_t* a = fsrc + cand; const uint8_t* b = src + p;
                const uint32_t maxlen = size - p;
                uint64_t x = ld8u(a) ^ v;
                if ((uint32_t)x == 0) {                                     // at least 4 bytes
                    // 8 bytes per probe while a whole probe fits in the chunk, then byte by byte: nothing past the last byte of
                    // the chunk is read (the source buffer may end exactly where its allocation ends)
                    for (;;) {
                        if (x) { len += (uint32_t)(__ffsll((long long)x) - 1) >> 3; break; }
                        len += 8;
                        if (len + 8 > maxlen) { while (len < maxlen && a[len] == b[len]) len++; break; }
                        x = ld8u(a + len) ^ ld8u(b + len);
                    }
                }
            }
            // greedy selection, left to right
            uint32_t avail = __ballot_sync(0xFFFFFFFFu, len >= kEncMinMatch);
            while (avail) {
                const uint32_t l = __ffs(avail) - 1;
                const uint32_t pl = base + l;
                const uint32_t ml = __shfl_sync(0xFFFFFFFFu, len, l);
#if FZ_ENC_LAZY
                // lazy step: every lane already knows its own match, so looking one position ahead is free -- a longer
                // match starting at the next byte wins and this byte becomes a literal
                if (pl >= cur && l < 31 && ((avail >> (l + 1)) & 1u) && __shfl_sync(0xFFFFFFFFu, len, l + 1) > ml) { avail &= ~(1u << l); continue; }
#endif
                const uint32_t off = foff + pl - (uint32_t)__shfl_sync(0xFFFFFFFFu, cand, l);
                if (pl >= cur) {
                    const uint32_t ll = pl - anchor;
                    for (uint32_t i = lane; i < ll; i += 32) lit[nlit + i] = src[anchor + i];
                    // Offset_Value: 1..3 name an entry of the history (shifted by one when the literal run is empty, where 3
                    // means rep0 - 1), anything else is the distance + 3.  Structured text repeats its distances, and a
                    // repeat code costs 2-3 bits against ~16 for a distance.
                    uint32_t ov = off + 3;
                    const uint32_t r0 = rep0, r1 = rep1, r2 = rep2;
                    if (known < 3) { known++; rep0 = off; rep1 = r0; rep2 = r1; }     // history still unknown: a plain offset
                    else if (ll) {
                        if (off == r0) ov = 1;
                        else if (off == r1) { ov = 2; rep0 = r1; rep1 = r0; }
                        else { if (off == r2) ov = 3; rep0 = off; rep1 = r0; rep2 = r1; }
                    } else {
                        if (off == r1) { ov = 1; rep0 = r1; rep1 = r0; }
                        else { if (off == r2) ov = 2; else if (off == r0 - 1 && off) ov = 3; rep0 = off; rep1 = r0; rep2 = r1; }
                    }
                    if (lane == 0) seq[nseq] = eseq_pack(ll, ml, ov);
                    nseq++; nlit += ll;
                    anchor = cur = pl + ml;
                }
                const uint32_t covered = cur - base;                        // lanes below `cur` are inside the match
                avail = covered >= 32 ? 0 : avail & ~((1u << covered) - 1) & ~((2u << l) - 1);
            }
        }
        const uint32_t rest = size - anchor;                                // trailing literals
        for (uint32_t i = lane; i < rest; i += 32) lit[nlit + i] = src[anchor + i];
        nlit += rest;
        if (lane == 0) { ch.nseq = nseq; ch.nlit = nlit; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ literals section
struct BitOut {             // forward bitstream (LSB first), flushed by bytes
    uint8_t* p; uint64_t acc; uint32_t n;
    __device__ __forceinline__ void init(uint8_t* dst) { p = dst; acc = 0; n = 0; }
    __device__ __forceinline__ void add(uint32_t v, uint32_t nb) { acc |= (uint64_t)v << n; n += nb; }   // caller keeps n <= 56
    __device__ __forceinline__ void flush() { while (n >= 8) { *p++ = (uint8_t)acc; acc >>= 8; n -= 8; } }
    __device__ __forceinline__ uint8_t* close() { add(1, 1); flush(); if (n) { *p++ = (uint8_t)acc; } return p; }   // sentinel bit
};

__device__ __forceinline__ uint32_t lit_header(uint8_t* dst, uint32_t type, uint32_t regen, uint32_t comp, bool four)
{
    if (type < 2) {                                   // Raw / RLE: 1, 2 or 3 bytes
        if (regen < 32) { dst[0] = (uint8_t)(type | (regen << 3)); return 1; }
        if (regen < 4096) { dst[0] = (uint8_t)(type | (1 << 2) | (regen << 4)); dst[1] = (uint8_t)(regen >> 4); return 2; }
        dst[0] = (uint8_t)(type | (3 << 2) | (regen << 4)); dst[1] = (uint8_t)(regen >> 4); dst[2] = (uint8_t)(regen >> 12); return 3;
    }
    if (regen < 1024 && comp < 1024) {
        const uint32_t v = type | ((four ? 1u : 0u) << 2) | (regen << 4) | (comp << 14);
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); return 3;
    }
    if (regen < 16384 && comp < 16384) {
        const uint32_t v = type | (2u << 2) | (regen << 4) | (comp << 18);
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); return 4;
    }
    const uint64_t v = type | (3u << 2) | ((uint64_t)regen << 4) | ((uint64_t)comp << 22);
    for (int i = 0; i < 5; i++) dst[i] = (uint8_t)(v >> (8 * i));
    return 5;
}
__device__ __forceinline__ uint32_t lit_header_size(uint32_t regen, uint32_t comp)
{
    return (regen < 1024 && comp < 1024) ? 3 : ((regen < 16384 && comp < 16384) ? 4 : 5);
}

constexpr int kLitWarps = 4;
__global__ void __launch_bounds__(kLitWarps * 32) k_enc_lit(EncChunk* chunks, uint32_t n_chunks, uint32_t* ticket)
{
    __shared__ uint32_t s_hist[kLitWarps][256];
    __shared__ uint16_t s_code[kLitWarps][256];
    __shared__ uint8_t s_len[kLitWarps][256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* hist = s_hist[warp]; uint16_t* code = s_code[warp]; uint8_t* len = s_len[warp];
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(ticket, 1);
        ci = __shfl_sync(0xFFFFFFFFu, ci, 0);
        if (ci >= n_chunks) return;
        EncChunk& ch = chunks[ci];
        const uint8_t* __restrict__ lit = ch.scratch + kScrLit;
        uint8_t* out = ch.scratch + kScrLitSec;
        const uint32_t nlit = ch.nlit;
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < nlit; i += 32) atomicAdd(&hist[lit[i]], 1u);
        __syncwarp();
        // per lane: symbols 8*lane .. 8*lane+7
        uint32_t cnt[8], nsym = 0, maxsym = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { cnt[j] = hist[8 * lane + j]; if (cnt[j]) { nsym++; maxsym = 8 * lane + j; } }
        nsym = __reduce_add_sync(0xFFFFFFFFu, nsym);
        maxsym = __reduce_max_sync(0xFFFFFFFFu, maxsym);
        uint32_t sec = 0;
        bool huf = nlit >= 256 && nsym >= 2 && maxsym <= 128;
        if (huf) {
            // ---- code lengths: ceil(log2(total / count)) capped at 11, then make the Kraft sum exactly 2^11
            uint32_t L[8]; uint32_t kraft = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                L[j] = 0;
                if (cnt[j]) { uint32_t l = 1; while (l < kHufMaxLen && ((uint64_t)cnt[j] << l) < nlit) l++; L[j] = l; kraft += 1u << (kHufMaxLen - l); }
            }
            kraft = __reduce_add_sync(0xFFFFFFFFu, kraft);
            while (kraft > (1u << kHufMaxLen)) {       // too many rare symbols at the cap: lengthen the rarest code that is below the cap
                uint32_t best = 0xFFFFFFFFu;           // key = count : 18 | length : 4 | symbol : 8
#pragma unroll
                for (int j = 0; j < 8; j++) if (cnt[j] && L[j] < kHufMaxLen) best = min(best, (cnt[j] << 12) | (L[j] << 8) | (8 * lane + j));
                best = __reduce_min_sync(0xFFFFFFFFu, best);
                const uint32_t s = best & 255, lold = (best >> 8) & 15;
#pragma unroll
                for (int j = 0; j < 8; j++) if (8 * lane + j == (int)s) L[j]++;
                kraft -= 1u << (kHufMaxLen - lold - 1);
            }
            while (kraft < (1u << kHufMaxLen)) {       // slack: shorten a longest code (its unit always fits), the most frequent one
                uint32_t lmax = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) lmax = max(lmax, L[j]);
                lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
                uint32_t best = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) if (L[j] == lmax) best = max(best, (cnt[j] << 8) | (8 * lane + j));
                best = __reduce_max_sync(0xFFFFFFFFu, best);
                const uint32_t s = best & 255;
#pragma unroll
                for (int j = 0; j < 8; j++) if (8 * lane + j == (int)s) L[j]--;
                kraft += 1u << (kHufMaxLen - lmax);
            }
            uint32_t lmax = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) lmax = max(lmax, L[j]);
            lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
#pragma unroll
            for (int j = 0; j < 8; j++) len[8 * lane + j] = (uint8_t)L[j];
            __syncwarp();
            // ---- canonical codes in the decoder's order: weight ascending (longest codes first), symbol ascending
            if (lane == 0) {
                uint32_t rank[kHufMaxLen + 2], next[kHufMaxLen + 2];
                for (uint32_t w = 0; w <= kHufMaxLen + 1; w++) rank[w] = 0;
                for (uint32_t s = 0; s <= maxsym; s++) if (len[s]) rank[lmax + 1 - len[s]]++;
                uint32_t cells = 0;
                for (uint32_t w = 1; w <= lmax; w++) { next[w] = cells >> (w - 1); cells += rank[w] << (w - 1); }
                for (uint32_t s = 0; s <= maxsym; s++) if (len[s]) { const uint32_t w = lmax + 1 - len[s]; code[s] = (uint16_t)next[w]++; }
            }
            __syncwarp();
            // ---- stream sizes first, so that the four streams can be written in place
            const uint32_t seg = (nlit + 3) / 4;
            uint32_t bits[4] = { 0, 0, 0, 0 };
            for (uint32_t i = lane; i < nlit; i += 32) {
                const uint32_t l = len[lit[i]];
                const uint32_t q = i / seg;
                bits[0] += q == 0 ? l : 0; bits[1] += q == 1 ? l : 0; bits[2] += q == 2 ? l : 0; bits[3] += q == 3 ? l : 0;
            }
            uint32_t bytes[4], total = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { bytes[q] = (__reduce_add_sync(0xFFFFFFFFu, bits[q]) + 1 + 7) / 8; total += bytes[q]; }
            const uint32_t tree = 1 + (maxsym + 1) / 2;          // header byte + maxsym direct weights (last one implicit)
            const uint32_t comp = tree + 6 + total;
            if (seg * 3 < nlit && comp + 8 < nlit && bytes[0] < 65536 && bytes[1] < 65536 && bytes[2] < 65536) {
                const uint32_t hs = lit_header_size(nlit, comp);
                if (lane == 0) {
                    lit_header(out, 2, nlit, comp, true);
                    uint8_t* t = out + hs;
                    t[0] = (uint8_t)(127 + maxsym);
                    for (uint32_t i = 0; i < maxsym; i += 2) {
                        const uint32_t w0 = len[i] ? lmax + 1 - len[i] : 0, w1 = (i + 1 < maxsym && len[i + 1]) ? lmax + 1 - len[i + 1] : 0;
                        t[1 + i / 2] = (uint8_t)((w0 << 4) | w1);
                    }
                    uint8_t* j = t + tree;
                    j[0] = (uint8_t)bytes[0]; j[1] = (uint8_t)(bytes[0] >> 8); j[2] = (uint8_t)bytes[1]; j[3] = (uint8_t)(bytes[1] >> 8);
                    j[4] = (uint8_t)bytes[2]; j[5] = (uint8_t)(bytes[2] >> 8);
                }
                if (lane < 4) {                                    // one Huffman stream per lane, symbols last to first
                    const uint32_t lo = lane * seg, hi = min(nlit, lo + seg);
                    uint32_t start = 0;
                    for (uint32_t q = 0; q < lane; q++) start += bytes[q];
                    BitOut bo; bo.init(out + hs + tree + 6 + start);
                    for (uint32_t i = hi; i > lo; i--) {
                        const uint32_t s = lit[i - 1];
                        bo.add(code[s], len[s]);
                        if (bo.n >= 40) bo.flush();
                    }
                    bo.close();
                }
                sec = hs + comp;
            } else huf = false;
        }
        if (!huf) {
            if (nsym == 1 && nlit > 1) {                           // RLE literals
                if (lane == 0) { const uint32_t hs = lit_header(out, 1, nlit, 0, false); out[hs] = lit[0]; sec = hs + 1; }
                sec = __shfl_sync(0xFFFFFFFFu, sec, 0);
            } else {                                               // Raw literals
                uint32_t hs = 0;
                if (lane == 0) hs = lit_header(out, 0, nlit, 0, false);
                hs = __shfl_sync(0xFFFFFFFFu, hs, 0);
                for (uint32_t i = lane; i < nlit; i += 32) out[hs + i] = lit[i];
                sec = hs + nlit;
            }
        }
        if (lane == 0) ch.lit_sec = sec;
        __syncwarp();
    }
}

// ------------------------------------------------------------------ sequences section (FSE)
// One warp per chunk.  All lanes histogram the LL / OF / ML codes; lanes 0..2 then each prepare one table:
// Predefined (few sequences), RLE (one symbol) or FSE_Compressed with counts normalised from the histogram
// (enc_normalize), its description (enc_write_ncount) and the encoding table (enc_build_ctable), all in shared
// memory; lane 0 finally writes the section: header, the three descriptions, the backward-readable bitstream
// (sequences last to first, RFC 8878 3.1.1.3.2.1.1 field order).
struct EncTables {
    uint32_t hist[3][56];
    int16_t norm[3][56];
    uint16_t state[3][512];
    uint32_t dnb[3][56];
    int32_t dfs[3][56];
    uint8_t tmp[3][512];
    uint16_t cumul[3][58];
    uint8_t desc[3][96];
    int desc_len[3], log[3], mode[3];
    uint32_t sbuf[144];      // bitstream staging: one batch of 32 sequences is at most 32 * 11 bytes
};

__device__ __forceinline__ uint32_t ll_code(uint32_t ll)
{
    if (ll < 16) return ll;
    if (ll < 32) return 16 + ((ll - 16) >> 1 < 4 ? (ll - 16) >> 1 : 4 + ((ll - 24) >> 2));   // 16,16,17,17,18,18,19,19,20 x4,21 x4
    if (ll < 64) return ll < 40 ? 22 : (ll < 48 ? 23 : 24);
    return (uint32_t)highbit(ll) + 19;
}
__device__ __forceinline__ uint32_t ml_code(uint32_t mlb)      // mlb = match length - 3
{
    if (mlb < 32) return mlb;
    if (mlb < 128) {
        if (mlb < 40) return 32 + ((mlb - 32) >> 1);                    // 32,32,33,33,34,34,35,35
        if (mlb < 48) return 36 + ((mlb - 40) >> 2);                    // 36 x4, 37 x4
        if (mlb < 64) return 38 + ((mlb - 48) >> 3);                    // 38 x8, 39 x8
        if (mlb < 96) return 40 + ((mlb - 64) >> 4);                    // 40 x16, 41 x16
        return 42;                                                      // 96 .. 127
    }
    return (uint32_t)highbit(mlb) + 36;
}

constexpr int kSeqEncWarps = 4;
__global__ void __launch_bounds__(kSeqEncWarps * 32) k_enc_seq(EncChunk* chunks, uint32_t n_chunks, uint32_t* ticket)
{
    __shared__ EncTables s_t[kSeqEncWarps];
    __shared__ SeqConsts K;
    for (uint32_t i = threadIdx.x; i < sizeof(SeqConsts) / 4; i += blockDim.x) ((uint32_t*)&K)[i] = ((const uint32_t*)&c_seq_consts)[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    EncTables& T = s_t[threadIdx.x >> 5];
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(ticket, 1);
        ci = __shfl_sync(0xFFFFFFFFu, ci, 0);
        if (ci >= n_chunks) return;
        EncChunk& ch = chunks[ci];
        const uint64_t* __restrict__ seq = (const uint64_t*)(ch.scratch + kScrSeq);
        uint8_t* out = ch.scratch + kScrSeqSec;
        const uint32_t nseq = ch.nseq;
        if (nseq == 0) { if (lane == 0) { out[0] = 0; ch.seq_sec = 1; } __syncwarp(); continue; }
        // ---- histograms of the three codes
        for (uint32_t i = lane; i < 3 * 56; i += 32) (&T.hist[0][0])[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < nseq; i += 32) {
            const uint64_t r = seq[i];
            atomicAdd(&T.hist[0][ll_code(eseq_ll(r))], 1u);
            atomicAdd(&T.hist[1][highbit(eseq_off(r))], 1u);
            atomicAdd(&T.hist[2][ml_code(eseq_ml(r) - 3)], 1u);
        }
        __syncwarp();
        // ---- one table per lane (0 LL, 1 OF, 2 ML)
        if (lane < 3) {
            const int t = (int)lane;
            const int n_sym = t == 0 ? 36 : (t == 1 ? 32 : 53), max_log = t == 1 ? 8 : 9;
            int present = 0, only = 0;
            for (int sy = 0; sy < n_sym; sy++) if (T.hist[t][sy]) { present++; only = sy; }
            int mode = nseq < 64 ? 0 : (present == 1 ? 1 : 2), log = 0, dlen = 0;
            if (mode == 2) {
                log = enc_table_log(nseq, present, max_log);
                if (enc_normalize(T.hist[t], n_sym, nseq, log, T.norm[t]) != 0) mode = 0;
                else { dlen = enc_write_ncount(T.desc[t], T.norm[t], n_sym, log); if (dlen < 0) mode = 0; }
            }
            if (mode == 1) { T.desc[t][0] = (uint8_t)only; dlen = 1; log = 0; }
            if (mode == 0) {
                const int16_t* def = t == 0 ? K.ll_def : (t == 1 ? K.of_def : K.ml_def);
                const int nd = t == 0 ? 36 : (t == 1 ? 29 : 53);
                for (int sy = 0; sy < 56; sy++) T.norm[t][sy] = sy < nd ? def[sy] : 0;
                log = t == 1 ? 5 : 6; dlen = 0;
            }
            if (mode != 1) enc_build_ctable(T.state[t], T.dnb[t], T.dfs[t], T.norm[t], mode == 0 ? (t == 0 ? 36 : (t == 1 ? 29 : 53)) : n_sym, log, T.tmp[t], T.cumul[t]);
            T.mode[t] = mode; T.log[t] = log; T.desc_len[t] = dlen;
        }
        __syncwarp();
        // ---- the section.  The bitstream is a serial chain (three FSE states, one bit container), but everything that
        // feeds it is not: the lanes load and pre-digest 32 sequences at a time (codes, extra bits), the chain then runs
        // on all lanes redundantly from shuffled values, the container is drained into shared memory and each batch's
        // bytes go to HBM with one cooperative copy.
        uint32_t hs = 0;
        if (lane == 0) {
            if (nseq < 128) { out[0] = (uint8_t)nseq; hs = 1; }
            else if (nseq < 0x7F00) { out[0] = (uint8_t)((nseq >> 8) + 128); out[1] = (uint8_t)nseq; hs = 2; }
            else { out[0] = 255; out[1] = (uint8_t)(nseq - 0x7F00); out[2] = (uint8_t)((nseq - 0x7F00) >> 8); hs = 3; }
            out[hs++] = (uint8_t)((T.mode[0] << 6) | (T.mode[1] << 4) | (T.mode[2] << 2));       // Symbol_Compression_Modes
            for (int t = 0; t < 3; t++) for (int i = 0; i < T.desc_len[t]; i++) out[hs++] = T.desc[t][i];   // LL, OF, ML
        }
        hs = __shfl_sync(0xFFFFFFFFu, hs, 0);
        uint8_t* const bs = out + hs;                                              // bitstream start
        const uint32_t limit = kEncChunkMax + kEncChunkMax / 2;                    // a section this large means a Raw block anyway
        uint32_t* const sw = T.sbuf;                                               // per-warp staging, 32-bit words
        uint64_t acc = 0; uint32_t nacc = 0, sfill = 0, gpos = 0;                  // container (< 32 bits between steps), words staged, bytes in HBM
        auto add = [&](uint32_t v, uint32_t nb) { acc |= (uint64_t)v << nacc; nacc += nb; };
        auto drain = [&]() { if (nacc >= 32) { if (lane == 0) sw[sfill] = (uint32_t)acc; sfill++; acc >>= 32; nacc -= 32; } };
        const bool rLL = T.mode[0] == 1, rOF = T.mode[1] == 1, rML = T.mode[2] == 1;
        auto enc = [&](int t, bool rle, uint32_t& st, uint32_t sy) {                // FSE_encodeSymbol
            if (rle) return;
            const uint32_t nb = (st + T.dnb[t][sy]) >> 16;
            add(st & ((1u << nb) - 1), nb);
            st = T.state[t][(st >> nb) + T.dfs[t][sy]];
        };
        auto init_state = [&](int t, bool rle, uint32_t sy) -> uint32_t {           // FSE_initCState2
            if (rle) return 0;
            const uint32_t nb = (T.dnb[t][sy] + (1u << 15)) >> 16;
            return T.state[t][((((nb << 16) - T.dnb[t][sy])) >> nb) + T.dfs[t][sy]];
        };
        uint32_t sLL = 0, sOF = 0, sML = 0;
        bool overflow = false;
        for (uint32_t hi = nseq; hi > 0 && !overflow;) {                            // sequences last to first, 32 per batch
            const uint32_t cnt = min(32u, hi);
            uint32_t codes = 0, lens = 0, ofx = 0;
            if (lane < cnt) {
                const uint64_t r = seq[hi - 1 - lane];
                const uint32_t ll = eseq_ll(r), mlb = eseq_ml(r) - 3, ofv = eseq_off(r);
                const uint32_t lc = ll_code(ll), mc = ml_code(mlb), oc = (uint32_t)highbit(ofv);
                codes = lc | (mc << 8) | (oc << 16) | ((uint32_t)K.ll_bits[lc] << 22) | ((uint32_t)K.ml_bits[mc] << 27);
                lens = (ll - K.ll_base[lc]) | ((mlb + 3 - K.ml_base[mc]) << 16);
                ofx = ofv - (1u << oc);
            }
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, codes, k), l = __shfl_sync(0xFFFFFFFFu, lens, k), o = __shfl_sync(0xFFFFFFFFu, ofx, k);
                const uint32_t lc = c & 255, mc = (c >> 8) & 255, oc = (c >> 16) & 63;
                if (hi == nseq && k == 0) { sML = init_state(2, rML, mc); sOF = init_state(1, rOF, oc); sLL = init_state(0, rLL, lc); }
                else { enc(1, rOF, sOF, oc); enc(2, rML, sML, mc); enc(0, rLL, sLL, lc); drain(); }   // <= 8 + 9 + 9 bits on top of < 32
                add(l & 0xFFFF, (c >> 22) & 31); add(l >> 16, c >> 27); drain();                       // <= 16 + 16
                add(o, oc); drain();                                                                    // <= 24 here (offsets inside a frame of <= 8 MiB)
            }
            hi -= cnt;
            // this batch's whole words -> HBM
            __syncwarp();
            const uint8_t* sb = (const uint8_t*)sw;
            for (uint32_t i = lane; i < sfill * 4; i += 32) bs[gpos + i] = sb[i];
            gpos += sfill * 4; sfill = 0;
            __syncwarp();
            overflow = hs + gpos > limit;
        }
        // FSE_flushCState: the final states, ML, OF, LL (0 bits for an RLE table), then the end mark
        if (!rML) add(sML - (1u << T.log[2]), (uint32_t)T.log[2]);
        if (!rOF) add(sOF - (1u << T.log[1]), (uint32_t)T.log[1]);
        drain();
        if (!rLL) add(sLL - (1u << T.log[0]), (uint32_t)T.log[0]);
        add(1, 1); drain();
        const uint32_t tail = (nacc + 7) / 8;                                      // < 4 bytes left in the container
        if (lane == 0) { sw[sfill] = (uint32_t)acc; }
        __syncwarp();
        {
            const uint8_t* sb = (const uint8_t*)sw;
            for (uint32_t i = lane; i < sfill * 4 + tail; i += 32) bs[gpos + i] = sb[i];
            gpos += sfill * 4 + tail;
        }
        if (lane == 0) ch.seq_sec = overflow ? kEncChunkMax : hs + gpos;
        __syncwarp();
    }
}

// ------------------------------------------------------------------ placement + final write
__device__ __forceinline__ uint32_t frame_header_size(uint32_t size) { return 4 + 1 + (size < 256 ? 1 : (size < 65536 + 256 ? 2 : 4)); }

// Seek table (FZG_SEEK_TABLE): the zstd "seekable format" -- a skippable frame after the last data frame holding
// (Compressed_Size, Decompressed_Size) of every frame and a footer (Number_Of_Frames, descriptor 0, 0x8F92EAB1).  Stock
// decoders skip it; fzg_decode_range uses it to decode only the frames a read touches.
__device__ __forceinline__ void st32le(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

__global__ void k_enc_place(const Item* items, EncChunk* chunks, const uint32_t* first_chunk, ItemOut* outs, uint32_t n_items, int seek)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    uint64_t pos = 0; int status = 0;
    const uint32_t c0 = first_chunk[i], c1 = first_chunk[i + 1];
    uint32_t nf = 0;
    for (uint32_t c = c0; c < c1; c++) {
        EncChunk& ch = chunks[c];
        const uint32_t body = ch.lit_sec + ch.seq_sec;
        const bool compressed = ch.size >= 32 && body + 8 < ch.size && body < kEncChunkMax;
        ch.raw = compressed ? 0 : 1;
        const bool first = ch.foff == 0, last = ch.foff + ch.size == ch.fsize;
        // the bytes this block puts into the file: the frame header before a frame's first block, the checksum after its last
        ch.frame_size = (first ? frame_header_size(ch.fsize) : 0) + 3 + (compressed ? body : ch.size) + (last ? 4 : 0);
        ch.out_off = pos; pos += ch.frame_size;
        nf += first;
    }
    const uint64_t table_at = pos;
    if (seek) pos += 8 + 8ull * nf + 9;
    if (pos > items[i].dst_cap) status = FZG_E_DSTSIZE;
    if (seek && !status) {
        uint8_t* t = items[i].dst + table_at;
        st32le(t, 0x184D2A5Eu); st32le(t + 4, 8 * nf + 9);
        uint32_t f = 0;
        for (uint32_t c = c0; c < c1;) {                              // one entry per frame: (compressed size, content size)
            uint32_t csz = 0; const uint32_t nb = chunks[c].nblk;
            for (uint32_t k = 0; k < nb; k++) csz += chunks[c + k].frame_size;
            st32le(t + 8 + 8 * f, csz); st32le(t + 12 + 8 * f, chunks[c].fsize);
            c += nb; f++;
        }
        uint8_t* e = t + 8 + 8ull * nf;
        st32le(e, nf); e[4] = 0; st32le(e + 5, 0x8F92EAB1u);
    }
    outs[i].dst_len = status ? 0 : pos; outs[i].status = status; outs[i].fail = status != 0;
}

__global__ void __launch_bounds__(128) k_enc_write(const Item* items, const EncChunk* chunks, const ItemOut* outs, uint32_t n_chunks)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ci >= n_chunks) return;
    const EncChunk& ch = chunks[ci];
    if (outs[ch.item].fail) return;
    uint8_t* o = items[ch.item].dst + ch.out_off;
    const uint32_t size = ch.size, fsize = ch.fsize;
    const bool compressed = ch.raw == 0, first = ch.foff == 0, last = ch.foff + size == fsize;
    const uint32_t body = compressed ? ch.lit_sec + ch.seq_sec : size;
    const uint32_t fh = first ? frame_header_size(fsize) : 0;
    if (lane == 0) {
        if (first) {
            o[0] = 0x28; o[1] = 0xB5; o[2] = 0x2F; o[3] = 0xFD;
            // Frame_Header_Descriptor: FCS flag | Single_Segment | Content_Checksum (what the reference's writer sets)
            if (fsize < 256) { o[4] = 0x24; o[5] = (uint8_t)fsize; }
            else if (fsize < 65536 + 256) { o[4] = 0x64; o[5] = (uint8_t)(fsize - 256); o[6] = (uint8_t)((fsize - 256) >> 8); }
            else { o[4] = 0xA4; o[5] = (uint8_t)fsize; o[6] = (uint8_t)(fsize >> 8); o[7] = (uint8_t)(fsize >> 16); o[8] = (uint8_t)(fsize >> 24); }
        }
        const uint32_t bh = (last ? 1u : 0u) | ((compressed ? 2u : 0u) << 1) | (body << 3);          // Last_Block | type | size
        o[fh] = (uint8_t)bh; o[fh + 1] = (uint8_t)(bh >> 8); o[fh + 2] = (uint8_t)(bh >> 16);
    }
    uint8_t* b = o + fh + 3;
    if (compressed) {
        const uint8_t* ls = ch.scratch + kScrLitSec; const uint8_t* ss = ch.scratch + kScrSeqSec;
        for (uint32_t i = lane; i < ch.lit_sec; i += 32) b[i] = ls[i];
        for (uint32_t i = lane; i < ch.seq_sec; i += 32) b[ch.lit_sec + i] = ss[i];
    } else {
        for (uint32_t i = lane; i < size; i += 32) b[i] = ch.src[i];
    }
}

// Content_Checksum of every frame: XXH64 is one serial chain over the frame's plain bytes, so a warp per FRAME (the warp of
// the frame's first chunk; lanes 0..3 own one accumulator each) walks the input once more and stores the low 32 bits
// after the frame's last block.
__global__ void __launch_bounds__(128) k_enc_hash(const Item* items, const EncChunk* chunks, const ItemOut* outs, uint32_t n_chunks)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ci >= n_chunks) return;
    const EncChunk& ch = chunks[ci];
    if (ch.foff != 0 || outs[ch.item].fail) return;
    const EncChunk& lastc = chunks[ci + ch.nblk - 1];
    const uint64_t acc = lane < 4 ? xx_lane(ch.src, ch.fsize, lane) : 0;
    const uint64_t v1 = __shfl_sync(0xFFFFFFFFu, acc, 0), v2 = __shfl_sync(0xFFFFFFFFu, acc, 1), v3 = __shfl_sync(0xFFFFFFFFu, acc, 2),
                   v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
    if (lane == 0) {
        const uint32_t x = (uint32_t)xx_combine(v1, v2, v3, v4, ch.src, ch.fsize);
        uint8_t* t = items[ch.item].dst + lastc.out_off + lastc.frame_size - 4;
        t[0] = (uint8_t)x; t[1] = (uint8_t)(x >> 8); t[2] = (uint8_t)(x >> 16); t[3] = (uint8_t)(x >> 24);
    }
}

}  // namespace fz

// ====================================================================== host side
using namespace fz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "fzgpu: %s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return -5 /*-EIO*/; } } while (0)

static const char* kEncStageNames[] = { "enc_match", "enc_lit", "enc_seq", "enc_place", "enc_write" };
const char* fzh_encode_stage_name(int s) { return s >= 0 && s < 5 ? kEncStageNames[s] : ""; }

static size_t enc_chunk_size(size_t chunk) { return chunk == 0 || chunk > kEncChunkMax ? kEncChunkMax : std::max<size_t>(chunk, 1024); }   // bytes per block
// bytes per independent frame (`chunk_size` of fzg_encode_batch): default 1 MiB = eight blocks that share a window
constexpr size_t kEncFrameDefault = 1u << 20, kEncFrameMax = 8u << 20;
static size_t enc_frame_size(size_t chunk) { return chunk == 0 ? kEncFrameDefault : std::min<size_t>(std::max<size_t>(chunk, 1024), kEncFrameMax); }

size_t fzh_encode_bound(size_t src_len, size_t chunk)
{
    const size_t cs = enc_chunk_size(chunk);
    const size_t frames = src_len ? (src_len + cs - 1) / cs : 1;
    return src_len + frames * 16 + 16 + (8 + 8 * frames + 9);          // + room for the seek table
}

int fzh_encode_setup(void)
{
    CK(cudaFuncSetAttribute(k_enc_match<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncMatchWarps * (2 << kEncHashLog)));
    CK(cudaFuncSetAttribute(k_enc_match<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncMatchWarps * (2 << kEncHashLog)));
    CK(cudaFuncSetAttribute(k_enc_match<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncMatchWarps * (2 << kEncHashLog)));
    return 0;
}

// Encodes items [first, first + n) of c->h_items (device pointers).  Results -> c->h_outs[first ..].  Blocking.
// The per-chunk scratch is large (5.5 bytes per input byte in the worst case), so the chunks are processed in waves.
int fzh_encode_run(FzCtx* c, uint32_t first, uint32_t n, int level, size_t chunk, int flags)
{
    // the reference passes 0..19, 0 = libzstd's default = 3 (src/main.rs:1237, 1283-1296): 1-2 select the single-table matcher
    const int mode = (level == 1 || level == 2) ? 0 : ((level == 0 || level == 3) ? 1 : 2);      // see EncMode
    cudaStream_t s = c->stream;
    const bool prof = flags & FZG_PROFILE;
    c->timing = fzg_timing_t{};
    if (n == 0) return 0;
    const Item* h_items = (const Item*)c->h_items.p + first;
    ItemOut* h_outs = (ItemOut*)c->h_outs.p + first;
    const size_t fs = enc_frame_size(chunk), cs = std::min<size_t>(fs, kEncChunkMax);      // frame / block sizes
    auto blocks_of = [&](uint64_t len) -> uint64_t {                  // blocks of a file: frames of fs bytes, blocks of cs bytes inside
        if (len == 0) return 1;
        const uint64_t full = len / fs, rem = len % fs;
        return full * ((fs + cs - 1) / cs) + (rem + cs - 1) / cs;
    };
    // ---- chunk (= block) descriptors, built on the host (sizes are known up front: src/main.rs:773 reads st_size)
    std::vector<uint32_t> first_chunk(n + 1);
    uint64_t n_chunks = 0;
    for (uint32_t i = 0; i < n; i++) { first_chunk[i] = (uint32_t)n_chunks; n_chunks += blocks_of(h_items[i].src_len); }
    first_chunk[n] = (uint32_t)n_chunks;
    if (n_chunks >= (1ull << 31)) return -22;
    int rc;
    if ((rc = c->e_chunks_h.reserve(n_chunks * sizeof(EncChunk)))) return rc;
    if ((rc = c->e_first_h.reserve((n + 1) * 4))) return rc;
    if ((rc = c->e_items.reserve(n_chunks * sizeof(EncChunk)))) return rc;
    if ((rc = c->d_outs.reserve(n * sizeof(ItemOut)))) return rc;
    if ((rc = c->d_totals.reserve(128))) return rc;
    // The per-chunk scratch is large, so the chunks are processed in waves of <= kWaveChunks (1 GiB of input).  Frame
    // offsets depend only on the chunks of the same item, so a wave that holds whole items is sized, placed and written
    // in one pass; an item larger than a wave is sized wave by wave first and then regenerated and written (the stages
    // are deterministic).
    constexpr uint64_t kWaveChunks = 8192;
    const char* wave_env = getenv("FZG_ENC_WAVE_CHUNKS");              // tests: a small wave puts ordinary files on the two-pass path
    const uint64_t wave_max = wave_env && atoi(wave_env) > 0 ? (uint64_t)atoi(wave_env) : kWaveChunks;
    const uint64_t wave = std::min<uint64_t>(n_chunks, wave_max);
    if ((rc = c->e_work.reserve(wave * (uint64_t)kScrBytes))) return rc;
    EncChunk* hc = (EncChunk*)c->e_chunks_h.p;
    memcpy(c->e_first_h.p, first_chunk.data(), (n + 1) * 4);
    struct Group { uint32_t ia, ib; };                 // items [ia, ib): one single-pass wave, or one oversized item
    std::vector<Group> groups;
    for (uint32_t ia = 0; ia < n;) {
        uint32_t ib = ia + 1;
        while (ib < n && first_chunk[ib + 1] - first_chunk[ia] <= wave) ib++;
        groups.push_back({ ia, ib }); ia = ib;
    }
    for (const Group& g : groups)
        for (uint32_t i = g.ia; i < g.ib; i++) {
            uint64_t off = 0; uint32_t frame_first = first_chunk[i];
            for (uint32_t k = first_chunk[i]; k < first_chunk[i + 1]; k++) {
                EncChunk& ch = hc[k];
                const uint64_t fstart = off - off % fs;                                    // the frame this block belongs to
                ch.fsize = (uint32_t)std::min<uint64_t>(fs, h_items[i].src_len - fstart); ch.foff = (uint32_t)(off - fstart);
                if (ch.foff == 0) frame_first = k;
                ch.src = h_items[i].src + off; ch.size = (uint32_t)std::min<uint64_t>(cs, fstart + ch.fsize - off); off += ch.size;
                ch.scratch = (uint8_t*)c->e_work.p + (uint64_t)((k - first_chunk[g.ia]) % wave) * kScrBytes;
                ch.item = i; ch.nseq = ch.nlit = ch.lit_sec = ch.seq_sec = ch.frame_size = ch.raw = 0; ch.out_off = 0; ch.nblk = 0; ch.pad = 0;
                hc[frame_first].nblk++;
            }
        }
    EncChunk* d_chunks = (EncChunk*)c->e_items.p;
    ItemOut* d_outs = (ItemOut*)c->d_outs.p;
    uint32_t* d_tickets = (uint32_t*)c->d_totals.p;
    CK(cudaMemcpyAsync(d_chunks, hc, n_chunks * sizeof(EncChunk), cudaMemcpyHostToDevice, s));
    const int seek = (flags & FZG_SEEK_TABLE) ? 1 : 0;
    int ev = 0;
    auto mark = [&]() { if (prof || ev == 0) cudaEventRecord(c->ev[ev], s); ev++; };
    mark();
    int launches = 0;
    uint32_t* d_first = (uint32_t*)c->e_first_h.p;                         // pinned host memory, read zero-copy
    static const bool dbg = getenv("FZG_DEBUG_SYNC") != nullptr;           // locate a faulting kernel: synchronise after every stage
    auto check = [&](const char* what) -> int {
        if (!dbg) return 0;
        const cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { fprintf(stderr, "fzgpu: %s: %s\n", what, cudaGetErrorString(e)); return -5; }
        return 0;
    };
    auto run_wave = [&](uint64_t lo, uint32_t cnt, bool marks) -> int {   // match -> literals -> sequences for chunks [lo, lo + cnt)
        CK(cudaMemsetAsync(d_tickets, 0, 16, s));
        static const uint32_t ctas_per_sm = [] { const char* e = getenv("FZG_ENC_CTAS_PER_SM"); const uint32_t full = (200u << 10) / (kEncMatchWarps * (2u << kEncHashLog));
                                                 return e && atoi(e) > 0 ? std::min<uint32_t>((uint32_t)atoi(e), full) : full; }();
        const uint32_t match_ctas = std::min<uint32_t>((cnt + kEncMatchWarps - 1) / kEncMatchWarps, 148 * ctas_per_sm);
        const uint32_t match_smem = kEncMatchWarps * (2 << kEncHashLog);
        if (mode == 2) {
            if (c->e_tab.reserve((size_t)match_ctas * kEncMatchWarps * EncMode<2>::long_bytes)) return -12;
            k_enc_match<2><<<match_ctas, kEncMatchWarps * 32, match_smem, s>>>(d_chunks + lo, cnt, d_tickets, (uint8_t*)c->e_tab.p);
        } else if (mode == 1) {
            if (c->e_tab.reserve((size_t)match_ctas * kEncMatchWarps * EncMode<1>::long_bytes)) return -12;
            k_enc_match<1><<<match_ctas, kEncMatchWarps * 32, match_smem, s>>>(d_chunks + lo, cnt, d_tickets, (uint8_t*)c->e_tab.p);
        } else k_enc_match<0><<<match_ctas, kEncMatchWarps * 32, match_smem, s>>>(d_chunks + lo, cnt, d_tickets, nullptr);
        if (check("k_enc_match")) return -5;
        if (marks) mark();
        k_enc_lit<<<std::min<uint32_t>((cnt + kLitWarps - 1) / kLitWarps, 148 * 8), kLitWarps * 32, 0, s>>>(d_chunks + lo, cnt, d_tickets + 1);
        if (check("k_enc_lit")) return -5;
        if (marks) mark();
        k_enc_seq<<<std::min<uint32_t>((cnt + kSeqEncWarps - 1) / kSeqEncWarps, 148 * 7), kSeqEncWarps * 32, 0, s>>>(d_chunks + lo, cnt, d_tickets + 2);
        if (check("k_enc_seq")) return -5;
        if (marks) mark();
        launches += 3;
        return 0;
    };
    const bool one = groups.size() == 1 && first_chunk[n] <= wave;         // per-stage events only make sense for a single wave
    for (const Group& g : groups) {
        const uint64_t lo = first_chunk[g.ia]; const uint64_t cnt = first_chunk[g.ib] - lo; const uint32_t ni = g.ib - g.ia;
        if (cnt <= wave) {
            if ((rc = run_wave(lo, (uint32_t)cnt, one))) return rc;
            k_enc_place<<<(ni + 127) / 128, 128, 0, s>>>(h_items + g.ia, d_chunks, d_first + g.ia, d_outs + g.ia, ni, seek); if (one) mark();
            if (check("k_enc_place")) return -5;
            k_enc_write<<<(uint32_t)((cnt * 32 + 127) / 128), 128, 0, s>>>(h_items, d_chunks + lo, d_outs, (uint32_t)cnt);
            k_enc_hash<<<(uint32_t)((cnt * 32 + 127) / 128), 128, 0, s>>>(h_items, d_chunks + lo, d_outs, (uint32_t)cnt); if (one) mark();
            if (check("k_enc_write")) return -5;
            launches += 3;
        } else {                                                           // one oversized item
            for (uint64_t w = lo; w < lo + cnt; w += wave) if ((rc = run_wave(w, (uint32_t)std::min<uint64_t>(wave, lo + cnt - w), false))) return rc;
            k_enc_place<<<1, 128, 0, s>>>(h_items + g.ia, d_chunks, d_first + g.ia, d_outs + g.ia, 1, seek);
            k_enc_hash<<<(uint32_t)((cnt * 32 + 127) / 128), 128, 0, s>>>(h_items, d_chunks + lo, d_outs, (uint32_t)cnt);       // needs the placement only
            launches += 2;
            for (uint64_t w = lo; w < lo + cnt; w += wave) {
                const uint32_t wc = (uint32_t)std::min<uint64_t>(wave, lo + cnt - w);
                if ((rc = run_wave(w, wc, false))) return rc;
                k_enc_write<<<(wc * 32 + 127) / 128, 128, 0, s>>>(h_items, d_chunks + w, d_outs, wc);
                launches++;
            }
        }
    }
    if (!one) { while (ev < 5) mark(); }
    if (!prof) ev = 5;
    cudaEventRecord(c->ev[ev], s);
    CK(cudaMemcpyAsync(h_outs, d_outs, n * sizeof(ItemOut), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    c->timing.launches = launches;
    cudaEventElapsedTime(&c->timing.total_ms, c->ev[0], c->ev[ev]);
    if (prof) for (int k = 0; k < 5; k++) cudaEventElapsedTime(&c->timing.kernel_ms[k], c->ev[k], c->ev[k + 1]);
    return 0;
}
