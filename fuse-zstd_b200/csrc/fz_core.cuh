/*
 * fz_core.cuh -- per-thread building blocks of the sm_100a zstd decoder.
 *
 * Everything here is written from the Zstandard format description (RFC 8878); it replaces the
 * libzstd arithmetic that fuse-zstd reaches through zstd::stream::copy_decode
 * (/root/reference/src/main.rs:463-467).  Functions are FZ_HD so the same source compiles
 *   - under nvcc into the kernels of fz_decode.cu (the product path), and
 *   - under g++ into tests/emul (TEST ONLY), which replays the launch sequence thread by thread
 *     so pipeline logic can be checked against the oracle without a GPU.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define FZ_HD __host__ __device__ __forceinline__
#define FZ_HD_RARE static __host__ __device__ __noinline__ // rare paths: kept out of line so that the hot loops stay small
#else
#define FZ_HD inline
#define FZ_HD_RARE inline
#endif

#include "../../include/fzgpu.h"

#ifdef __CUDA_ARCH__
#define FZ_SYNCWARP(mask) __syncwarp(mask)
#else
#define FZ_SYNCWARP(mask) ((void)(mask))
#endif

namespace fz {

constexpr uint32_t kMagic = 0xFD2FB528u;
constexpr uint32_t kBlockMax = 128u * 1024u;
constexpr uint64_t kWindowMax = 1ull << 27;   // zstd-rs streaming decoder default (windowLogMax = 27)
constexpr int kMaxLL = 35, kMaxOF = 31, kMaxML = 52;
constexpr int kLLLog = 9, kOFLog = 8, kMLLog = 9, kHufLogMax = 12;

// ------------------------------------------------------------------ descriptors (HBM)
struct Item {            // one .zst file of the batch
    const uint8_t* src; uint64_t src_len;
    uint8_t* dst; uint64_t dst_cap;
};

struct ItemInfo {        // written by the count pass, consumed by the scan
    uint32_t n_frames, n_blocks, n_seq_jobs, n_huf_jobs;
    uint64_t lit_bytes, n_seq;
    int32_t walk_status; uint32_t pad;
};

struct ItemBase {        // exclusive prefix sums over items
    uint32_t frame, block, seq_job, huf_job;
    uint64_t lit, seq;
};

struct Frame {
    uint32_t item, first_block, n_blocks, block_max;
    uint64_t fcs;            // Frame_Content_Size (valid when has_fcs)
    uint64_t out_off;        // offset of the frame's first byte in the item's dst
    uint64_t out_size;       // regenerated size (sum of block rsize)
    uint32_t checksum;       // stored XXH64 low 32 bits
    uint8_t has_fcs, has_checksum; uint16_t pad;
    int32_t status; uint32_t pad2;
};

enum : uint8_t { BT_RAW = 0, BT_RLE = 1, BT_COMPRESSED = 2 };
enum : uint8_t { LT_RAW = 0, LT_RLE = 1, LT_HUF = 2, LT_TREELESS = 3 };

struct Block {
    const uint8_t* src;      // block content (after the 3-byte header)
    const uint8_t* lit;      // regenerated literals: into src (Raw) or into literal scratch
    uint64_t seq_base;       // first record in the sequence scratch
    uint64_t out_off;        // offset in item dst (filled by the offsets pass)
    uint32_t csize;          // Block_Size field
    uint32_t rsize;          // regenerated size (Raw/RLE: known; Compressed: filled by the sequence pass)
    uint32_t lit_hdr, lit_regen, lit_csize;
    uint32_t nseq, seq_hdr;  // seq_hdr: offset of the byte after nbSeq(+modes) inside the block
    uint32_t frame;
    int32_t huf_src, ll_src, of_src, ml_src;   // global block index that carries the table description
    uint32_t rep_out[3];     // repeat-offset history after the block, possibly symbolic (sequence pass)
    uint32_t rep_in[3];      // resolved history at the start of the block (offsets pass)
    uint8_t type, last, lit_type, lit_streams, modes, pad[3];
    int32_t status;
};

// 8-byte sequence record in HBM, produced by the sequence pass and consumed by the execute pass:
//   E  [0:18)   output position (inside the block) just after this sequence's match
//   LE [18:36)  literals consumed (inside the block) after this sequence's literal run
//   off[36:64)  match distance: concrete (<= 2^27) or a symbolic reference to the repeat-offset
//               history the block started with (see off_sym)
// Sequence i therefore writes literals lit[LE(i-1) .. LE(i)) at E(i-1), then its match up to E(i):
// every copy's source and destination are known without a serial scan.
FZ_HD uint64_t rec_pack(uint32_t e, uint32_t le, uint32_t off) { return (uint64_t)e | ((uint64_t)le << 18) | ((uint64_t)off << 36); }
FZ_HD uint32_t rec_e(uint64_t r) { return (uint32_t)r & 0x3FFFFu; }
FZ_HD uint32_t rec_le(uint64_t r) { return (uint32_t)(r >> 18) & 0x3FFFFu; }
FZ_HD uint32_t rec_off(uint64_t r) { return (uint32_t)(r >> 36); }

// Repeat offsets across blocks (RFC 8878 3.1.1.5).  Blocks are entropy-decoded in parallel, so a
// block does not know the three-entry history it starts with.  The sequence pass therefore starts
// from three symbols in_0..in_2; a history slot is either a concrete distance (<= kOffMax) or
// "max(in_k - d, 1)" (the `rep0 - 1` rule applied d times), encoded kOffMax | k << 24 | (d + 1).
// The offsets pass resolves each block's starting history serially (a few values per block) and the
// execute pass resolves the few records that still carry a symbol, in parallel.
constexpr uint32_t kOffMax = 1u << 27;         // largest window the reference's streaming decoder accepts
FZ_HD uint32_t off_sym(uint32_t k) { return kOffMax | (k << 24) | 1u; }
FZ_HD uint32_t off_dec(uint32_t v) { return v > kOffMax ? v + 1 : (v > 1 ? v - 1 : 1u); }   // rep0 - 1, 0 -> 1
FZ_HD uint32_t off_resolve(uint32_t v, uint32_t in0, uint32_t in1, uint32_t in2)
{
    if (v <= kOffMax) return v;
    const uint32_t k = (v >> 24) & 3u, d = (v & 0xFFFFFFu) - 1u;
    const uint32_t x = k == 0 ? in0 : (k == 1 ? in1 : in2);
    return x > d ? x - d : 1u;
}

// ------------------------------------------------------------------ small helpers
FZ_HD int highbit(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
FZ_HD uint32_t shl_c(uint32_t v, uint32_t n)   // clamped shifts: a count >= 32 gives 0 (PTX shl/shr semantics)
{
#ifdef __CUDA_ARCH__
    uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r;
#else
    return n >= 32 ? 0u : v << n;
#endif
}
FZ_HD uint32_t shr_c(uint32_t v, uint32_t n)
{
#ifdef __CUDA_ARCH__
    uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r;
#else
    return n >= 32 ? 0u : v >> n;
#endif
}
FZ_HD uint32_t fsl_c(uint32_t lo, uint32_t hi, uint32_t n)                        // high word of (hi:lo << n), n clamped to 32
{
#ifdef __CUDA_ARCH__
    return __funnelshift_lc(lo, hi, n);
#else
    return n >= 32 ? lo : (n == 0 ? hi : (hi << n) | (lo >> (32 - n)));
#endif
}
FZ_HD uint32_t rd24(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
FZ_HD uint32_t rd32u(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// ------------------------------------------------------------------ frame header
struct FrameHdr { uint64_t fcs, window; uint32_t hsize; uint8_t has_fcs, checksum; };

// RFC 8878 3.1.1.1.  p points at the magic; n = bytes available.  Returns FZG_* status.
FZ_HD int parse_frame_header(const uint8_t* p, uint64_t n, FrameHdr& h)
{
    if (n < 5) return FZG_E_TRUNCATED;
    uint32_t fhd = p[4];
    uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    if (fhd & 8) return FZG_E_UNSUPPORTED;
    h.checksum = (fhd >> 2) & 1;
    uint32_t pos = 5;
    h.window = 0;
    if (!single) {
        if (n < pos + 1) return FZG_E_TRUNCATED;
        uint32_t b = p[pos++];
        uint64_t base = 1ull << (10 + (b >> 3));
        h.window = base + (base >> 3) * (b & 7);
    }
    uint32_t db = did_flag == 3 ? 4 : did_flag;
    if (n < pos + db) return FZG_E_TRUNCATED;
    uint32_t did = 0;
    for (uint32_t i = 0; i < db; i++) did |= (uint32_t)p[pos + i] << (8 * i);
    pos += db;
    if (did != 0) return FZG_E_UNSUPPORTED;
    uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (n < pos + fb) return FZG_E_TRUNCATED;
    h.has_fcs = fb != 0; h.fcs = 0;
    for (uint32_t i = 0; i < fb; i++) h.fcs |= (uint64_t)p[pos + i] << (8 * i);
    if (fb == 2) h.fcs += 256;
    pos += fb;
    if (single) h.window = h.fcs;
    if (h.window > kWindowMax) return FZG_E_UNSUPPORTED;
    h.hsize = pos;
    return FZG_OK;
}

// ------------------------------------------------------------------ literals / sequences section headers
struct LitHdr { uint32_t type, streams, hsize, regen, csize; };

// RFC 8878 3.1.1.3.1.1.  Returns 0 or FZG_E_CORRUPT.
FZ_HD int parse_lit_header(const uint8_t* p, uint32_t n, uint32_t block_max, LitHdr& h)
{
    if (n < 1) return FZG_E_CORRUPT;
    uint32_t b0 = p[0];
    h.type = b0 & 3; uint32_t sf = (b0 >> 2) & 3;
    h.streams = 1; h.csize = 0;
    if (h.type < 2) {
        if (sf == 0 || sf == 2) { h.hsize = 1; h.regen = b0 >> 3; }
        else if (sf == 1) { if (n < 2) return FZG_E_CORRUPT; h.hsize = 2; h.regen = (b0 >> 4) | ((uint32_t)p[1] << 4); }
        else { if (n < 3) return FZG_E_CORRUPT; h.hsize = 3; h.regen = (b0 >> 4) | ((uint32_t)p[1] << 4) | ((uint32_t)p[2] << 12); }
        if (h.regen > block_max) return FZG_E_CORRUPT;
        h.csize = h.type == LT_RAW ? h.regen : 1;
        if (h.hsize + h.csize > n) return FZG_E_CORRUPT;
        return 0;
    }
    if (sf < 2) {
        if (n < 3) return FZG_E_CORRUPT;
        uint32_t v = rd24(p);
        h.hsize = 3; h.regen = (v >> 4) & 0x3FF; h.csize = (v >> 14) & 0x3FF; h.streams = sf ? 4 : 1;
    } else if (sf == 2) {
        if (n < 4) return FZG_E_CORRUPT;
        uint32_t v = rd32u(p);
        h.hsize = 4; h.regen = (v >> 4) & 0x3FFF; h.csize = (v >> 18) & 0x3FFF; h.streams = 4;
    } else {
        if (n < 5) return FZG_E_CORRUPT;
        uint64_t v = (uint64_t)rd32u(p) | ((uint64_t)p[4] << 32);
        h.hsize = 5; h.regen = (uint32_t)((v >> 4) & 0x3FFFF); h.csize = (uint32_t)((v >> 22) & 0x3FFFF); h.streams = 4;
    }
    if (h.regen > block_max || h.regen == 0) return FZG_E_CORRUPT;
    if (h.streams == 4 && h.regen < 6) return FZG_E_CORRUPT;
    if (h.hsize + h.csize > n) return FZG_E_CORRUPT;
    return 0;
}

// Sequences_Section_Header: nbSeq and the modes byte.  p = start of the section, n = bytes left
// in the block.  hsize = bytes up to and including the modes byte (1 when nseq == 0).
FZ_HD int parse_seq_header(const uint8_t* p, uint32_t n, uint32_t& nseq, uint32_t& hsize, uint32_t& modes)
{
    if (n < 1) return FZG_E_CORRUPT;
    nseq = p[0]; hsize = 1; modes = 0;
    if (nseq >= 128) {
        if (nseq == 255) { if (n < 3) return FZG_E_CORRUPT; nseq = p[1] + ((uint32_t)p[2] << 8) + 0x7F00; hsize = 3; }
        else { if (n < 2) return FZG_E_CORRUPT; nseq = ((nseq - 128) << 8) + p[1]; hsize = 2; }
    }
    if (nseq == 0) return hsize == n ? 0 : FZG_E_CORRUPT;
    if (n < hsize + 1) return FZG_E_CORRUPT;
    modes = p[hsize]; hsize += 1;
    if (modes & 3) return FZG_E_CORRUPT;
    return 0;
}

// ------------------------------------------------------------------ item walk (count pass and fill pass)
// One thread walks one item: frames, skippable frames, block headers, literals / sequences section
// headers.  FILL=false only counts; FILL=true writes Frame / Block descriptors and job lists at
// the bases the scan produced.  Both passes take identical decisions.
template <bool FILL>
FZ_HD void walk_item(uint32_t item_idx, const Item& it, ItemInfo& info, const ItemBase* base,
                     Frame* frames, Block* blocks, uint32_t* seq_jobs, uint32_t* huf_jobs,
                     uint8_t* lit_scratch)
{
    const uint8_t* src = it.src; const uint64_t n = it.src_len;
    uint64_t ip = 0;
    uint32_t nf = 0, nb = 0, nsj = 0, nhj = 0; uint64_t lit_bytes = 0, n_seq = 0;
    int status = FZG_OK;
    while (ip < n) {
        if (n - ip < 4) { status = FZG_E_TRUNCATED; break; }
        uint32_t magic = rd32u(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - ip < 8) { status = FZG_E_TRUNCATED; break; }
            uint64_t sz = rd32u(src + ip + 4);
            if (n - ip - 8 < sz) { status = FZG_E_TRUNCATED; break; }
            ip += 8 + sz; continue;
        }
        if (magic != kMagic) { status = FZG_E_MAGIC; break; }
        FrameHdr fh;
        status = parse_frame_header(src + ip, n - ip, fh);
        if (status) break;
        ip += fh.hsize;
        uint32_t block_max = fh.window < kBlockMax ? (uint32_t)fh.window : kBlockMax;
        uint32_t frame_gidx = 0, first_block = 0;
        if (FILL) { frame_gidx = base->frame + nf; first_block = base->block + nb; }
        int32_t huf_src = -1, ll_src = -1, of_src = -1, ml_src = -1;
        uint32_t frame_blocks = 0;
        for (;;) {
            if (n - ip < 3) { status = FZG_E_TRUNCATED; break; }
            uint32_t bh = rd24(src + ip); ip += 3;
            uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) { status = FZG_E_CORRUPT; break; }
            if (bsize > block_max) { status = FZG_E_CORRUPT; break; }
            uint32_t csize = type == BT_RLE ? 1 : bsize;
            if (n - ip < csize) { status = FZG_E_TRUNCATED; break; }
            Block b;
            b.src = src + ip; b.lit = nullptr; b.seq_base = 0; b.out_off = 0;
            b.csize = bsize; b.rsize = type == BT_COMPRESSED ? 0 : bsize;
            b.lit_hdr = b.lit_regen = b.lit_csize = b.nseq = b.seq_hdr = 0;
            b.frame = frame_gidx; b.huf_src = b.ll_src = b.of_src = b.ml_src = -1;
            b.type = (uint8_t)type; b.last = (uint8_t)last; b.lit_type = 0; b.lit_streams = 0; b.modes = 0;
            b.pad[0] = b.pad[1] = b.pad[2] = 0; b.status = 0;
            for (int r = 0; r < 3; r++) { b.rep_out[r] = off_sym((uint32_t)r); b.rep_in[r] = 0; }
            uint32_t gb = FILL ? base->block + nb : 0;
            if (type == BT_COMPRESSED) {
                if (bsize < 2) { status = FZG_E_CORRUPT; break; }
                LitHdr lh;
                status = parse_lit_header(src + ip, bsize, block_max, lh);
                if (status) break;
                uint32_t litsec = lh.hsize + lh.csize;
                uint32_t nseq, shs, modes;
                status = parse_seq_header(src + ip + litsec, bsize - litsec, nseq, shs, modes);
                if (status) break;
                b.lit_hdr = lh.hsize; b.lit_regen = lh.regen; b.lit_csize = lh.csize;
                b.lit_type = (uint8_t)lh.type; b.lit_streams = (uint8_t)lh.streams;
                b.nseq = nseq; b.seq_hdr = litsec + shs; b.modes = (uint8_t)modes;
                if (nseq == 0) b.rsize = lh.regen;         // no sequences: the block regenerates exactly its literals
                // table provenance is tracked with item-local block numbers (identical in both passes)
                if (lh.type == LT_HUF) huf_src = (int32_t)nb;
                else if (lh.type == LT_TREELESS && huf_src < 0) { status = FZG_E_CORRUPT; break; }
                if (nseq) {
                    uint32_t mll = (modes >> 6) & 3, mof = (modes >> 4) & 3, mml = (modes >> 2) & 3;
                    if (mll != 3) ll_src = (int32_t)nb; else if (ll_src < 0) { status = FZG_E_CORRUPT; break; }
                    if (mof != 3) of_src = (int32_t)nb; else if (of_src < 0) { status = FZG_E_CORRUPT; break; }
                    if (mml != 3) ml_src = (int32_t)nb; else if (ml_src < 0) { status = FZG_E_CORRUPT; break; }
                }
                if (FILL) {
                    const int32_t g0 = (int32_t)base->block;
                    b.huf_src = lh.type >= LT_HUF ? g0 + huf_src : -1;
                    if (nseq) { b.ll_src = g0 + ll_src; b.of_src = g0 + of_src; b.ml_src = g0 + ml_src; }
                }
                if (lh.type == LT_RAW) b.lit = src + ip + lh.hsize;
                else {
                    if (FILL) b.lit = lit_scratch + base->lit + lit_bytes;
                    lit_bytes += (lh.regen + 15u) & ~15u;
                }
                if (nseq) {
                    if (FILL) { b.seq_base = base->seq + n_seq; seq_jobs[base->seq_job + nsj] = gb; }
                    n_seq += (nseq + 1u + 3u) & ~3u; nsj++;     // + the tail record (k_records); records of a block start on a 32-byte boundary (256-bit stores)
                }
                if (lh.type != LT_RAW) { if (FILL) huf_jobs[base->huf_job + nhj] = gb; nhj++; }
            }
            if (FILL) blocks[gb] = b;
            ip += csize; nb++; frame_blocks++;
            if (last) break;
        }
        if (status) break;
        uint32_t stored = 0;
        if (fh.checksum) {
            if (n - ip < 4) { status = FZG_E_TRUNCATED; break; }
            stored = rd32u(src + ip); ip += 4;
        }
        if (FILL) {
            Frame f;
            f.item = item_idx; f.first_block = first_block; f.n_blocks = frame_blocks; f.block_max = block_max;
            f.fcs = fh.fcs; f.out_off = 0; f.out_size = 0; f.checksum = stored;
            f.has_fcs = fh.has_fcs; f.has_checksum = fh.checksum; f.pad = 0; f.status = 0; f.pad2 = 0;
            frames[frame_gidx] = f;
        }
        nf++;
    }
    info.n_frames = nf; info.n_blocks = nb; info.n_seq_jobs = nsj; info.n_huf_jobs = nhj;
    info.lit_bytes = lit_bytes; info.n_seq = n_seq; info.walk_status = status; info.pad = 0;
}

// ------------------------------------------------------------------ forward bit reader (FSE table descriptions)
struct FwdBits {
    const uint8_t* p; uint32_t n; uint32_t bitpos;
    FZ_HD uint32_t peek(uint32_t nb) const   // nb <= 24; zero padded past the end
    {
        uint32_t byte = bitpos >> 3, sh = bitpos & 7; uint32_t w = 0;
        for (uint32_t i = 0; i < 4; i++) if (byte + i < n) w |= (uint32_t)p[byte + i] << (8 * i);
        return (w >> sh) & ((1u << nb) - 1);
    }
};

// RFC 8878 4.1.1.  norm[] receives probabilities (-1 = "less than one").  Returns bytes used or -1.
// Single exit (errors are carried in a flag): lanes of a warp that parse different descriptions reconverge after every
// loop instead of running the rest of the caller one after the other.
template <class NORM>
FZ_HD int read_ncount(const uint8_t* p, uint32_t n, int max_sym, int max_log, NORM norm, int& n_sym, int& log)
{
    FwdBits br{ p, n, 0 };
    bool bad = n == 0;
    log = bad ? 5 : (int)br.peek(4) + 5; br.bitpos = 4;
    if (log > max_log) bad = true;
    int remaining = bad ? 0 : 1 << log, sym = 0;
    while (remaining > 0 && sym <= max_sym && !bad) {
        int bits = highbit((uint32_t)remaining + 1) + 1;
        uint32_t v = br.peek((uint32_t)bits);
        uint32_t lower = (1u << (bits - 1)) - 1;
        uint32_t thr = (1u << bits) - 1 - ((uint32_t)remaining + 1);
        if ((v & lower) < thr) { v &= lower; br.bitpos += (uint32_t)bits - 1; }
        else { if (v > lower) v -= thr; br.bitpos += (uint32_t)bits; }
        int prob = (int)v - 1;
        remaining -= prob < 0 ? 1 : prob;
        norm[sym++] = (int16_t)prob;
        if (prob == 0) {
            uint32_t rep = 3;
            while (rep == 3 && !bad) {
                rep = br.peek(2); br.bitpos += 2;
                for (uint32_t i = 0; i < rep && !bad; i++) { if (sym > max_sym) bad = true; else norm[sym++] = 0; }
            }
        }
        if (br.bitpos > n * 8) bad = true;
    }
    if (remaining != 0 || br.bitpos > n * 8) bad = true;
    n_sym = sym;
    return bad ? -1 : (int)((br.bitpos + 7) >> 3);
}

// FSE decode cell, 32 bits: baseline[0:16) | nbBits[16:20) | nbExtra[20:25) | symbol[25:32)
FZ_HD uint32_t cell_pack(uint32_t base, uint32_t nb, uint32_t extra, uint32_t sym) { return base | (nb << 16) | (extra << 20) | (sym << 25); }
FZ_HD uint32_t cell_base(uint32_t c) { return c & 0xFFFFu; }
FZ_HD uint32_t cell_nb(uint32_t c) { return (c >> 16) & 15u; }
FZ_HD uint32_t cell_sym(uint32_t c) { return c >> 25; }

// Builds the decode table (RFC 8878 4.1.1).  extra_bits[sym] is folded into each cell so the
// sequence loop needs one lookup per state.  `cnt` is per-thread scratch of >= 64 uint16.
// Returns 0 or -1.  table has 1<<log cells.
FZ_HD int build_fse_table(uint32_t* table, const int16_t* norm, int n_sym, int log, const uint8_t* extra_bits, uint16_t* cnt)
{
    const int size = 1 << log; int high = size - 1;
    for (int s = 0; s < n_sym; s++) {
        if (norm[s] == -1) { table[high--] = (uint32_t)s; cnt[s] = 1; }
        else cnt[s] = (uint16_t)norm[s];
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1; int pos = 0;
    for (int s = 0; s < n_sym; s++) {
        for (int i = 0; i < norm[s]; i++) {
            table[pos] = (uint32_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    if (pos != 0) return -1;
    for (int u = 0; u < size; u++) {
        uint32_t s = table[u];
        uint32_t nx = cnt[s]++;
        uint32_t nb = (uint32_t)(log - highbit(nx));
        table[u] = cell_pack((nx << nb) - (uint32_t)size, nb, extra_bits ? extra_bits[s] : s, s);
    }
    return 0;
}

// ------------------------------------------------------------------ backward bit reader (Huffman + sequences)
// Top-aligned 64-bit container in two 32-bit registers, refilled 32 bits at a time from
// 4-byte-aligned words (coalescing-friendly and legal for any stream alignment).  Words wholly
// below the stream start are never loaded (zeros are supplied); `left` tracks the unread bits
// of the real stream, so a stream that over-reads ends with left < 0 and is reported corrupt.
struct BackBits {
    uint32_t hi, lo; int avail; int left;
    const uint32_t* wp; const uint32_t* wmin;

    FZ_HD int init(const uint8_t* p, uint32_t n)
    {
        if (n == 0) return -1;
        const uint8_t* lastp = p + n - 1;
        uint32_t lastb = *lastp;
        if (lastb == 0) return -1;
        uintptr_t a = (uintptr_t)lastp & ~(uintptr_t)3;
        uint32_t keep = (uint32_t)((uintptr_t)lastp & 3) + 1;
        uint32_t w = *(const uint32_t*)a;
        if (keep < 4) w &= (1u << (8 * keep)) - 1;
        int hb = highbit(w);                        // sentinel position inside the word
        hi = shl_c(w, 32 - hb); lo = 0; avail = hb;
        left = (int)(n - 1) * 8 + highbit(lastb);
        wp = (const uint32_t*)a - 1;
        wmin = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
        return 0;
    }
    FZ_HD void refill()                             // precondition: avail <= 32
    {
        uint32_t w = wp >= wmin ? *wp : 0u;
        wp--;
        hi |= shr_c(w, (uint32_t)avail);
        lo = shl_c(w, 32 - (uint32_t)avail);
        avail += 32;
    }
    FZ_HD uint32_t peek(uint32_t nb) const { return shr_c(hi, 32 - nb); }   // nb <= 32, nb <= avail
    FZ_HD void skip(uint32_t nb)
    {
        hi = fsl_c(lo, hi, nb); lo = shl_c(lo, nb);
        avail -= (int)nb; left -= (int)nb;
    }
    FZ_HD uint32_t read(uint32_t nb) { uint32_t v = peek(nb); skip(nb); return v; }
};

// ------------------------------------------------------------------ Huffman literals
struct HufInfo { int log; uint32_t used; };   // used = bytes of the tree description

// RFC 8878 4.2.1.  Decodes the tree description at p into weights w[0..n_w) (incl. the implicit
// last weight).  Returns 0 or -1.  `ft` is scratch for the 64-cell FSE table, cnt >= 64 uint16.
FZ_HD int huf_read_weights(const uint8_t* p, uint32_t n, uint8_t* w, int& n_w, HufInfo& info, uint32_t* ft, uint16_t* cnt)
{
    if (n < 1) return -1;
    uint32_t h = p[0]; int nw = 0;
    if (h >= 128) {
        nw = (int)h - 127;
        uint32_t bytes = (uint32_t)(nw + 1) / 2;
        if (1 + bytes > n) return -1;
        for (int i = 0; i < nw; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        info.used = 1 + bytes;
    } else {
        if (h == 0 || h + 1 > n) return -1;
        int16_t norm[16]; int ns, log;
        int hb = read_ncount(p + 1, h, 12, 6, norm, ns, log);
        if (hb < 0 || (uint32_t)hb >= h) return -1;
        if (build_fse_table(ft, norm, ns, log, nullptr, cnt) != 0) return -1;
        BackBits br;
        if (br.init(p + 1 + hb, h - (uint32_t)hb) != 0) return -1;
        br.refill(); if (br.avail <= 32) br.refill();
        uint32_t s1 = br.read((uint32_t)log), s2 = br.read((uint32_t)log);
        if (br.left < 0) return -1;
        for (;;) {
            if (br.avail <= 32) br.refill();
            if (nw > 253) return -1;
            uint32_t c1 = ft[s1];
            w[nw++] = (uint8_t)cell_sym(c1);
            s1 = cell_base(c1) + br.read(cell_nb(c1));
            if (br.left < 0) { w[nw++] = (uint8_t)cell_sym(ft[s2]); break; }
            if (nw > 253) return -1;
            uint32_t c2 = ft[s2];
            w[nw++] = (uint8_t)cell_sym(c2);
            s2 = cell_base(c2) + br.read(cell_nb(c2));
            if (br.left < 0) { w[nw++] = (uint8_t)cell_sym(ft[s1]); break; }
        }
        info.used = 1 + h;
    }
    uint32_t sum = 0; int rank1 = 0;
    for (int i = 0; i < nw; i++) {
        if (w[i] > kHufLogMax) return -1;
        if (w[i]) sum += 1u << (w[i] - 1);
        rank1 += w[i] == 1;
    }
    if (sum == 0) return -1;
    int log = highbit(sum) + 1;
    if (log > kHufLogMax) return -1;
    uint32_t left = (1u << log) - sum;
    if (left & (left - 1)) return -1;
    int last = highbit(left) + 1;
    w[nw++] = (uint8_t)last; rank1 += last == 1;
    if (rank1 < 2 || (rank1 & 1)) return -1;
    n_w = nw; info.log = log;
    return 0;
}

// Canonical table fill: cells are (symbol | nbBits << 8), 1 << log of them.  Weight 1 (longest
// codes) first, ascending symbol inside a weight.
FZ_HD void huf_fill_table(uint16_t* table, const uint8_t* w, int n_w, int log, uint32_t* start /*[kHufLogMax + 2]*/, uint32_t* rank /*[kHufLogMax + 2]*/)
{
    for (int k = 0; k <= kHufLogMax + 1; k++) rank[k] = 0;
    for (int s = 0; s < n_w; s++) rank[w[s]]++;
    uint32_t cur = 0;
    for (int k = 1; k <= log; k++) { start[k] = cur; cur += rank[k] << (k - 1); }
    for (int s = 0; s < n_w; s++) {
        uint32_t wt = w[s]; if (!wt) continue;
        uint32_t len = 1u << (wt - 1), at = start[wt];
        uint16_t cell = (uint16_t)((uint32_t)s | ((uint32_t)(log + 1 - (int)wt) << 8));
        for (uint32_t i = 0; i < len; i++) table[at + i] = cell;
        start[wt] = at + len;
    }
}

// ------------------------------------------------------------------ sequence tables
struct SeqConsts {       // small read-only tables, staged in shared memory by the kernels
    uint32_t ll_base[36]; uint32_t ml_base[53];
    uint8_t ll_bits[36]; uint8_t ml_bits[53];
    int16_t ll_def[36]; int16_t of_def[29]; int16_t ml_def[53];
};

#define FZ_LL_BASE { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 }
#define FZ_ML_BASE { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 }
#define FZ_LL_BITS { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 }
#define FZ_ML_BITS { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 }
#define FZ_LL_DEF { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 }
#define FZ_OF_DEF { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 }
#define FZ_ML_DEF { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 }

// Position and mode of table `which` (0 LL, 1 OF, 2 ML) inside block b's sequences section.
// Returns the mode (0 predefined, 1 rle, 2 fse); sets p/n to the description bytes.  -1 on error.
FZ_HD int locate_table(const Block& b, int which, const uint8_t*& p, uint32_t& n)
{
    uint32_t pos = b.seq_hdr; int res = -1; bool bad = false;
    for (int t = 0; t <= which; t++) {                       // single exit: see read_ncount
        const int mode = (b.modes >> (6 - 2 * t)) & 3;
        if (pos > b.csize) bad = true;
        if (bad) continue;
        if (t == which) { p = b.src + pos; n = b.csize - pos; res = mode; }
        else if (mode == 1) pos += 1;
        else if (mode == 2) {
            int16_t norm[64]; int ns, log;
            const int used = read_ncount(b.src + pos, b.csize - pos, t == 0 ? kMaxLL : (t == 1 ? kMaxOF : kMaxML),
                                         t == 0 ? kLLLog : (t == 1 ? kOFLog : kMLLog), norm, ns, log);
            if (used < 0) bad = true; else pos += (uint32_t)used;
        }
    }
    return bad ? -1 : res;
}

// ------------------------------------------------------------------ sequence bitstream reader
// The sequence pass is one serial dependency chain per block and the lanes of a warp run different
// blocks in lockstep, so the reader must neither wait for HBM nor branch.  The backward bitstream is
// mirrored into a 256-byte shared-memory ring (ring byte = global address & 255) by cp.async, 16 bytes
// at a time and ~240 bytes ahead of use; the consumer addresses the ring with a bit cursor (SeqCursor).
FZ_HD void ring_commit()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N> FZ_HD void ring_wait()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---- shared-memory access of the chain loop: 32-bit shared-window addresses on the device (one LDS with an
// immediate offset, no generic-address arithmetic), plain pointers in the host emulation
#ifdef __CUDA_ARCH__
typedef uint32_t sm_t;
FZ_HD sm_t sm_of(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FZ_HD uint32_t sm_ld16(sm_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
FZ_HD uint32_t sm_ld32(sm_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
FZ_HD sm_t ring_slot(sm_t ring, uint32_t b) { return ring | (b & 0xFCu); }            // the ring is 256-byte aligned
FZ_HD void ring_fetch_sm(sm_t slot, const uint8_t* gsrc) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gsrc) : "memory"); }
FZ_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
FZ_HD uint32_t fsl_w(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_l(lo, hi, n); }   // high word of (hi:lo << (n & 31))
FZ_HD uint32_t andn32(uint32_t a, uint32_t b) { uint32_t r; asm("lop3.b32 %0, %1, %2, 0, 0x30;" : "=r"(r) : "r"(a), "r"(b)); return r; }   // a & ~b
FZ_HD uint32_t log2p(uint32_t v) { uint32_t r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }                                   // v = 2^k -> k
FZ_HD sm_t opaque(sm_t v) { asm("" : "+r"(v)); return v; }          // keeps an address sum out of the reassociation of the final add
FZ_HD void st_rec4(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d)      // four records, one 256-bit store (p 32-byte aligned)
{
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
#else
typedef uintptr_t sm_t;
FZ_HD sm_t sm_of(const void* p) { return (uintptr_t)p; }
FZ_HD uint32_t sm_ld16(sm_t a) { return *(const uint16_t*)a; }
FZ_HD uint32_t sm_ld32(sm_t a) { return *(const uint32_t*)a; }
FZ_HD sm_t ring_slot(sm_t ring, uint32_t b) { return ring + (b & 0xFCu); }
FZ_HD void ring_fetch_sm(sm_t slot, const uint8_t* gsrc) { for (int i = 0; i < 16; i++) ((uint8_t*)slot)[i] = gsrc[i]; }
FZ_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
FZ_HD uint32_t fsl_w(uint32_t lo, uint32_t hi, uint32_t n) { n &= 31; return n ? (hi << n) | (lo >> (32 - n)) : hi; }
FZ_HD uint32_t andn32(uint32_t a, uint32_t b) { return a & ~b; }
FZ_HD uint32_t log2p(uint32_t v) { return (uint32_t)highbit(v); }
FZ_HD sm_t opaque(sm_t v) { return v; }
FZ_HD void st_rec4(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
#endif

// Backward bitstream of a sequences section, addressed by a BIT CURSOR instead of a bit container: `cur` is the bit
// address of the highest unread bit, counted from gbase = (stream pointer & ~255), so that (cur >> 3) & 255 is at the
// same time the byte's slot in the 256-byte cp.async ring.  A sequence needs no refill logic and no branches: it
// loads the three ring words under the cursor, funnels them into a 64-bit window and moves the cursor down.
// Bits below the first byte of the stream read as garbage; a stream that uses them ends below `lo` and is corrupt.
struct SeqCursor {
    int32_t cur, lo;         // lo = bit address of the stream's first bit; everything is read <=> cur == lo - 1
    const uint8_t* gbase;
    sm_t ring;
    int32_t fa, fmin;        // byte offset (from gbase) of the next 16-byte chunk to fetch, going down / of the lowest chunk

    FZ_HD void fill()        // every free ring slot: a slot is free once the cursor's word is below the chunk it held
    {
        const int32_t bw = (cur >> 3) & ~3;
        while (fa >= fmin && fa + 256 > bw) { ring_fetch_sm(ring + ((uint32_t)fa & 255u), gbase + fa); fa -= 16; }
    }
    FZ_HD int init(const uint8_t* p, uint32_t n, uint8_t* ring_)
    {
        if (n == 0) return -1;
        const uint8_t* lastp = p + n - 1;
        const uint32_t lastb = *lastp;
        if (lastb == 0) return -1;
        gbase = (const uint8_t*)((uintptr_t)p & ~(uintptr_t)255);
        ring = sm_of(ring_);
        lo = (int32_t)(p - gbase) * 8;
        cur = (int32_t)(lastp - gbase) * 8 + highbit(lastb) - 1;     // just below the sentinel bit
        fa = (int32_t)(lastp - gbase) & ~15; fmin = (int32_t)(p - gbase) & ~15;
        for (int i = 0; i < 16 && fa >= fmin; i++) { ring_fetch_sm(ring + ((uint32_t)fa & 255u), gbase + fa); fa -= 16; }
        ring_commit(); ring_wait<0>();
        return 0;
    }
    FZ_HD uint32_t peek() const                        // the 32 bits under the cursor, highest first
    {
        const uint32_t b = (uint32_t)(cur >> 3);
        return fsl_w(sm_ld32(ring_slot(ring, b - 4)), sm_ld32(ring_slot(ring, b)), ~(uint32_t)cur & 31u);
    }
    FZ_HD uint32_t read(uint32_t nb) { const uint32_t v = shr_c(peek(), 32 - nb); cur -= (int32_t)nb; return v; }   // nb <= 32
};

// One Huffman stream (RFC 8878 4.2.2): regenerates n_out bytes at `out` from the n bytes at p, through the
// ring-fed cursor reader.  Lanes of a warp decode different streams in lockstep: `bound` is the warp-uniform number
// of 4-symbol iterations (>= n_out / 4 of every lane in `mask`), `ok` false masks the lane out.  Symbols are
// stored four at a time once the output is 4-byte aligned.  Returns 0 or -1.
// Four symbols take at most 4 * 12 bits: one 64-bit window (three ring words) per iteration, shifted in registers
// after each symbol; the chain per symbol is index -> LDS -> length -> shift.  Bits below the first byte of the stream
// read as garbage: a valid stream never depends on them (they only fill the don't-care part of a table index), a
// stream that consumes them ends below `lo` and is corrupt.
FZ_HD int huf_decode_stream(const uint16_t* table, int log, const uint8_t* p, uint32_t n, uint8_t* out, uint32_t n_out,
                            uint8_t* ring, uint32_t bound, uint32_t mask, bool ok)
{
    SeqCursor cs; cs.cur = 0; cs.lo = 0; cs.gbase = nullptr; cs.ring = sm_of(ring); cs.fa = -1; cs.fmin = 0;
    if (ok && cs.init(p, n, ring) != 0) ok = false;
    const uint32_t sh = 32u - (uint32_t)log;                      // log >= 1 for any valid tree
    const sm_t tab = sm_of(table);
    uint32_t i = 0;
    if (ok) {
        uint32_t head = (uint32_t)((4 - ((uintptr_t)out & 3)) & 3);                        // <= 3 symbols up to alignment
        if (head > n_out) head = n_out;
        for (; i < head; i++) { const uint32_t c = sm_ld16(tab + 2 * shr_c(cs.peek(), sh)); cs.cur -= (int32_t)(c >> 8); out[i] = (uint8_t)c; }
    }
    const uint32_t quads = ok ? (n_out - i) / 4 : 0;
    FZ_SYNCWARP(mask);
    for (uint32_t q = 0; q < bound; q++) {
        if ((q & 3) == 0) {          // <= 24 bytes are consumed between two visits: a chunk fetched at visit v is not read before visit v + 8
            if (q < quads) cs.fill();
            ring_commit(); ring_wait<6>();
        }
        if (q < quads) {
            const uint32_t bb = (uint32_t)(cs.cur >> 3), t = ~(uint32_t)cs.cur & 31u;
            const uint32_t wa = sm_ld32(ring_slot(cs.ring, bb)), wb = sm_ld32(ring_slot(cs.ring, bb - 4)), wc = sm_ld32(ring_slot(cs.ring, bb - 8));
            uint32_t x = fsl_w(wb, wa, t), x1 = fsl_w(wc, wb, t);
            const uint32_t c0 = sm_ld16(tab + 2 * shr_c(x, sh)), n0 = c0 >> 8; x = fsl_c(x1, x, n0); x1 = shl_c(x1, n0);
            const uint32_t c1 = sm_ld16(tab + 2 * shr_c(x, sh)), n1 = c1 >> 8; x = fsl_c(x1, x, n1); x1 = shl_c(x1, n1);
            const uint32_t c2 = sm_ld16(tab + 2 * shr_c(x, sh)), n2 = c2 >> 8; x = fsl_c(x1, x, n2);
            const uint32_t c3 = sm_ld16(tab + 2 * shr_c(x, sh)), n3 = c3 >> 8;
            cs.cur -= (int32_t)(n0 + n1 + n2 + n3);
            *(uint32_t*)(out + i) = (c0 & 255u) | ((c1 & 255u) << 8) | ((c2 & 255u) << 16) | (c3 << 24);
            i += 4;
        }
    }
    ring_wait<0>();
    if (!ok) return -1;
    for (; i < n_out; i++) {
        cs.fill(); ring_commit(); ring_wait<0>();
        const uint32_t c = sm_ld16(tab + 2 * shr_c(cs.peek(), sh)); cs.cur -= (int32_t)(c >> 8); out[i] = (uint8_t)c;
    }
    return cs.cur == cs.lo - 1 ? 0 : -1;
}

// ------------------------------------------------------------------ sequence pass, stage A: the FSE chain
// The three-state FSE chain is the only inherently serial part of a block, so stage A does nothing
// else: one thread per block walks the backward bitstream and leaves one 8-byte RAW record per
// sequence; everything that can be done for many sequences at once (states -> symbols, extra bits ->
// values, running positions, repeat offsets) is stage B (warp per block, fz_decode.cu).
//
// Shared memory per stream (2816 bytes, 82 streams per SM):
//   uint16 cLL[512] | uint16 cML[512] | uint16 cOF[256] | ring[256]
// A chain cell is 16 bits: J[0:10) | extra[10:15), where J encodes (baseline, nbBits) jointly as
// ((baseline >> nb) << 1 | 1) << nb -- nb = ctz(J), baseline = (J & (J - 1)) >> 1 -- and `extra` is the
// number of extra bits of the cell's symbol (for offsets that IS the symbol).  The LL / ML symbols are
// not kept here: the RAW record carries the STATES, and stage B maps state -> symbol with its own
// copy of the (cheap) symbol spread.
//
// RAW record, fast form (the extra bits of the sequence are the top bits of the 32-bit window x):
//   x[0:32) | stateLL[32:41) | stateML[41:50) | offset code[50:55) | 0
// RAW record, slow form (long lengths / offsets, decoded field by field):
//   ll[0:18) | (ml - 3)[18:35) | offset_value[35:63) | 1[63]
constexpr uint32_t kChainCellsLL = 512, kChainCellsML = 512, kChainCellsOF = 256;
constexpr uint32_t kChainCellBytes = (kChainCellsLL + kChainCellsML + kChainCellsOF) * 2;
constexpr uint32_t kChainBytes = kChainCellBytes + 256;

FZ_HD uint32_t chain_pack(uint32_t base, uint32_t nb, uint32_t extra) { return ((((base >> nb) << 1) | 1u) << nb) | (extra << 10); }
// Where table `which` (0 LL, 1 OF, 2 ML) of block b is described (its own sequences section, or -- Repeat mode --
// the section of the block recorded in b.*_src) and how.  p / n on entry: the current position in b's section.
struct TableSrc { int mode; const uint8_t* p; uint32_t n; bool own; };
FZ_HD int resolve_table(const Block* blocks, const Block& b, int which, const uint8_t* p, uint32_t n, TableSrc& t)
{
    t.mode = (b.modes >> (6 - 2 * which)) & 3; t.p = p; t.n = n; t.own = t.mode != 3;
    int rc = 0;
    if (!t.own) {
        const int32_t src = which == 0 ? b.ll_src : (which == 1 ? b.of_src : b.ml_src);
        if (src < 0) rc = -1;
        else {
            t.mode = locate_table(blocks[src], which, t.p, t.n);
            if (t.mode < 0 || t.mode == 3) rc = -1;
        }
    }
    return rc;
}
// Per-thread work arrays of the table stage, element i of thread t at base[i * stride + t]: in shared memory the 32 lanes
// of a warp then touch consecutive bytes (no bank conflicts) although each builds a different table.
template <class T> struct Lanewise {
    T* p; uint32_t stride;
    FZ_HD T& operator[](uint32_t i) const { return p[i * stride]; }
};
struct TabWork { Lanewise<uint8_t> sym; Lanewise<int16_t> norm; Lanewise<uint16_t> cnt; };   // 512 symbols, 64 counts, 64 counters

FZ_HD void st_cells8(uint16_t* p, const uint32_t* w)          // eight 16-bit cells, one 16-byte store (p 16-byte aligned)
{
#ifdef __CUDA_ARCH__
    *(uint4*)p = make_uint4(w[0], w[1], w[2], w[3]);
#else
    for (int i = 0; i < 4; i++) { p[2 * i] = (uint16_t)w[i]; p[2 * i + 1] = (uint16_t)(w[i] >> 16); }
#endif
}

// Chain cells (and, for LL / ML, the state -> symbol map) of table `which` of block b, straight to HBM.
// The lanes of a warp build different tables, so every loop here has a trip count that depends only on the table size:
// the symbol spread walks all 1 << log visits of the (pos + step) permutation instead of looping per symbol, and the
// cells leave eight at a time.
FZ_HD int build_chain_seq_table(const Block* blocks, const Block& b, int which, const uint8_t* p, uint32_t n,
                                const SeqConsts& K, uint16_t* cell, int& log, uint32_t& used, const TabWork& w, uint8_t* ymap)
{
    const int max_sym = which == 0 ? kMaxLL : (which == 1 ? kMaxOF : kMaxML);
    const int max_log = which == 0 ? kLLLog : (which == 1 ? kOFLog : kMLLog);
    const uint8_t* extra = which == 0 ? K.ll_bits : (which == 2 ? K.ml_bits : nullptr);
    TableSrc t; used = 0; log = 0;
    bool bad = resolve_table(blocks, b, which, p, n, t) != 0;
    int ns = 0;
    if (!bad && t.mode == 1) {
        if (t.n < 1 || t.p[0] > max_sym) bad = true;
        else {
            const uint32_t sy = t.p[0];
            cell[0] = (uint16_t)chain_pack(0, 0, extra ? extra[sy] : sy);
            if (ymap) ymap[0] = (uint8_t)sy;
            if (t.own) used = 1;
        }
    } else if (!bad && t.mode == 0) {
        const int16_t* def = which == 0 ? K.ll_def : (which == 1 ? K.of_def : K.ml_def);
        ns = which == 0 ? 36 : (which == 1 ? 29 : 53); log = which == 1 ? 5 : 6;
        for (int i = 0; i < ns; i++) w.norm[i] = def[i];
    } else if (!bad) {
        const int u = read_ncount(t.p, t.n, max_sym, max_log, w.norm, ns, log);
        if (u < 0) { bad = true; ns = 0; log = 0; }
        else if (t.own) used = (uint32_t)u;
    }
    const bool fse = !bad && t.mode != 1;               // a table to spread (Predefined or FSE_Compressed)
    const int size = fse ? 1 << log : 0; int high = size - 1;
    for (int sy = 0; sy < ns; sy++) {
        const int pr = w.norm[sy];
        if (pr == -1) { w.sym[high--] = (uint8_t)sy; w.cnt[sy] = 1; } else w.cnt[sy] = (uint16_t)pr;
    }
    {   // spread (RFC 8878 4.1.1): visit k lands on (k * step) & mask; visits above `high` are skipped, the others take the symbols in order
        const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
        int sy = 0, rem = 0, pos = 0, placed = 0;
        for (int k = 0; k < size; k++) {
            if (pos <= high) {
                while (rem <= 0 && sy < ns) { rem = w.norm[sy]; sy++; }          // next symbol with a probability >= 1 (sy - 1 is current)
                if (rem <= 0) bad = true;                                         // more free cells than probabilities: malformed
                w.sym[pos] = (uint8_t)(sy > 0 ? sy - 1 : 0); rem--; placed++;
            }
            pos = (pos + step) & mask;
        }
        while (rem <= 0 && sy < ns) { rem = w.norm[sy]; sy++; }
        if (fse && (rem > 0 || placed != high + 1)) bad = true;                   // probabilities left over
    }
    for (int c0 = 0; c0 < size; c0 += 8) {                                        // cells, eight at a time (size >= 32)
        uint32_t cw[4] = { 0, 0, 0, 0 }, yw[2] = { 0, 0 };
        for (int j = 0; j < 8; j++) {
            const uint32_t sy = w.sym[c0 + j];
            const uint32_t nx = w.cnt[sy] | (bad ? 1u : 0u); w.cnt[sy] = (uint16_t)(nx + 1);   // (a malformed table may hold zero counts)
            const uint32_t nb = (uint32_t)(log - highbit(nx)) & 15u;
            cw[j >> 1] |= (chain_pack((nx << nb) - (uint32_t)size, nb, extra ? extra[sy] : sy) & 0xFFFFu) << (16 * (j & 1));
            yw[j >> 2] |= sy << (8 * (j & 3));
        }
        st_cells8(cell + c0, cw);
        if (ymap) { ((uint32_t*)(ymap + c0))[0] = yw[0]; ((uint32_t*)(ymap + c0))[1] = yw[1]; }
    }
    return bad ? -1 : 0;
}

// ------------------------------------------------------------------ sequence pass, table stage (one thread per block)
// Everything stage A and stage B need to know about a block's three FSE tables, built once, by one thread per block with
// all blocks of the batch in flight (the build is serial, data-dependent code: inside stage A it kept 31 lanes of a warp
// waiting for the slowest), into HBM: the chain cells that stage A copies into shared memory, the state -> symbol maps
// of the LL and ML tables that stage B needs, and a small header.
//   per job: uint16 cLL[512] | cML[512] | cOF[256] | uint8 yLL[512] | yML[512]   (kJobTableBytes = 3584)
struct SeqJobHdr { uint32_t bits_off; uint8_t logLL, logOF, logML, bad; };       // bits_off: first byte of the bitstream inside the block
constexpr uint32_t kJobTableBytes = kChainCellBytes + 1024;
FZ_HD void seq_tables_thread(const Block* blocks, const Block& b, const SeqConsts& K, uint8_t* tab, SeqJobHdr& h, const TabWork& w)
{
    uint16_t* cLL = (uint16_t*)tab; uint16_t* cML = cLL + kChainCellsLL; uint16_t* cOF = cML + kChainCellsML;
    uint8_t* yLL = tab + kChainCellBytes; uint8_t* yML = yLL + 512;
    const uint8_t* p = b.src + b.seq_hdr; uint32_t n = b.csize - b.seq_hdr; uint32_t used = 0;
    int logLL = 0, logOF = 0, logML = 0; bool bad = false;
    if (build_chain_seq_table(blocks, b, 0, p, n, K, cLL, logLL, used, w, yLL) != 0 || used > n) bad = true;
    if (!bad) { p += used; n -= used; if (build_chain_seq_table(blocks, b, 1, p, n, K, cOF, logOF, used, w, nullptr) != 0 || used > n) bad = true; }
    if (!bad) { p += used; n -= used; if (build_chain_seq_table(blocks, b, 2, p, n, K, cML, logML, used, w, yML) != 0 || used > n) bad = true; }
    if (!bad) { p += used; n -= used; }
    h.bits_off = (uint32_t)(p - b.src); h.logLL = (uint8_t)logLL; h.logOF = (uint8_t)logOF; h.logML = (uint8_t)logML; h.bad = bad;
}

// RAW record of stage A: x[0:32) | y[32:64), y = LL cell byte offset [0:12) | ML cell byte offset [12:24) | offset code [24:29) |
// far [31], the offsets counted from the start of the stream's shared memory (LL cells at 0, ML cells at 1024): exactly
// what the chain loop holds in registers.  x is the 32-bit window of the bitstream that starts with the sequence's
// extra bits (offset, match length, literal length); when those exceed 32 bits (far = 1: long lengths with a
// far offset, rare) x is the sequence's BIT CURSOR instead and stage B reads the extra bits from the stream itself.
FZ_HD uint64_t raw_pack(uint32_t x, uint32_t y) { return (uint64_t)x | ((uint64_t)y << 32); }
FZ_HD uint32_t stream_bits(const uint8_t* gbase, int32_t cur, uint32_t nb)     // nb <= 32 bits of the stream, the highest at bit address cur
{
    uint64_t v = 0;
    for (int32_t k = 0; k < (int32_t)nb; k++) { const int32_t a = cur - k; v = (v << 1) | (a >= 0 ? (gbase[a >> 3] >> (a & 7)) & 1u : 0u); }
    return (uint32_t)v;
}
// RAW -> (literal length, match length, offset value), with the block's state -> symbol maps; `bits` = first byte of the
// block's sequence bitstream.  Returns false when the offset code is beyond any legal window.
FZ_HD bool raw_unpack(uint64_t r, const SeqConsts& K, const uint8_t* yLL, const uint8_t* yML, const uint8_t* bits, uint32_t& ll, uint32_t& ml, uint32_t& ofv)
{
    const uint32_t x = (uint32_t)r, y = (uint32_t)(r >> 32);
    const uint32_t yll = yLL[(y & 0x3FFu) >> 1], yml = yML[((y >> 12) & 0x3FFu) >> 1], yof = (y >> 24) & 31;
    const uint32_t ofb = yof, mlb = K.ml_bits[yml], llb = K.ll_bits[yll];
    if (y >> 31) {
        const uint8_t* gbase = (const uint8_t*)((uintptr_t)bits & ~(uintptr_t)255);
        const int32_t cur = (int32_t)x;
        ofv = (1u << yof) + stream_bits(gbase, cur, ofb);
        ml = K.ml_base[yml] + stream_bits(gbase, cur - (int32_t)ofb, mlb);
        ll = K.ll_base[yll] + stream_bits(gbase, cur - (int32_t)(ofb + mlb), llb);
        return yof <= 27;
    }
    ofv = (1u << yof) + shr_c(x, 32 - ofb);
    ml = K.ml_base[yml] + shr_c(shl_c(x, ofb), 32 - mlb);
    ll = K.ll_base[yll] + shr_c(shl_c(x, ofb + mlb), 32 - llb);
    return yof <= 27;
}
// RFC 8878 3.1.1.5 on a (possibly symbolic) history; returns the match distance of this sequence.
FZ_HD uint32_t rep_update(uint32_t ofv, bool ll0, uint32_t& rep0, uint32_t& rep1, uint32_t& rep2)
{
    uint32_t off;
    if (ofv > 3) { off = ofv - 3; rep2 = rep1; rep1 = rep0; rep0 = off; }
    else {
        const uint32_t idx = ofv - 1 + (ll0 ? 1u : 0u);
        if (idx == 0) off = rep0;
        else {
            off = idx == 3 ? off_dec(rep0) : (idx == 1 ? rep1 : rep2);
            if (idx != 1) rep2 = rep1;
            rep1 = rep0; rep0 = off;
        }
    }
    return off;
}

// The registers of one chain: addresses of the current LL / ML / OF cells, the table bases, and the constant that turns
// the LL / ML addresses into the record's offsets.
struct ChainRegs { sm_t aLL, aML, aOF, tLL, tML, tOF; uint32_t ypack; };

// One sequence of one chain: emits its RAW record and moves the three states and the cursor on.  Straight-line code.
// vote_mask != 0: all lanes of vote_mask are here together (main loop) and the rare sequence with more than 32 extra
// bits is handled under a warp-uniform branch; vote_mask == 0: the caller is already divergent (ragged end).
template <bool VOTED>
FZ_HD uint64_t chain_step(ChainRegs& r, SeqCursor& cs, uint32_t vote_mask)
{
    const uint32_t cl = sm_ld16(r.aLL), co = sm_ld16(r.aOF), cm = sm_ld16(r.aML);
    const int32_t cur0 = cs.cur;
    const uint32_t bb = (uint32_t)(cur0 >> 3), t = ~(uint32_t)cur0 & 31u;
    const uint32_t wa = sm_ld32(ring_slot(cs.ring, bb)), wb = sm_ld32(ring_slot(cs.ring, bb - 4)), wc = sm_ld32(ring_slot(cs.ring, bb - 8));
    const uint32_t x = fsl_w(wb, wa, t), x1 = fsl_w(wc, wb, t);                               // the 64 bits under the cursor
    const uint32_t llb = cl >> 10, ofb = co >> 10, mlb = cm >> 10, a3 = llb + ofb + mlb;      // extra bits of this sequence (<= 63)
    const uint32_t l1 = cl - 1, o1 = co - 1, m1 = cm - 1;
    const uint32_t pLL = andn32(cl, l1), pOF = andn32(co, o1), pML = andn32(cm, m1);          // 2^nb of the three updates
    const sm_t bLL = opaque(r.tLL + (cl & l1 & 1023u)), bOF = opaque(r.tOF + (co & o1 & 1023u)), bML = opaque(r.tML + (cm & m1 & 1023u));   // cells of the baselines
    const uint32_t nsum = log2p(pLL * pML * pOF);
    uint32_t y, xrec = x, yrec = (uint32_t)r.aLL + ((uint32_t)r.aML << 12) + (ofb << 24) - r.ypack;
    const bool far = a3 > 32;
#if defined(__CUDA_ARCH__) && !defined(FZ_CHAIN_VOTED)
    if (VOTED) {                                          // no control flow at all: a fourth ring word covers a3 <= 63
        const uint32_t wd = sm_ld32(ring_slot(cs.ring, bb - 12)), x2 = fsl_w(wd, wc, t);
        y = fsl_c(fsl_c(x2, x1, a3), fsl_c(x1, x, a3), far ? a3 - 32 : 0u);
        xrec = far ? (uint32_t)cur0 : x; yrec += far ? 1u << 31 : 0u;
    } else
#endif
    {
        y = fsl_c(x1, x, a3);                             // the 32 bits after the extra bits
#ifdef __CUDA_ARCH__
        if (__builtin_expect(VOTED ? __any_sync(vote_mask, far) : far, 0))
#else
        if (far)
#endif
        {
            if (far) {
                SeqCursor c2 = cs; c2.cur = cur0 - (int32_t)a3;
                y = c2.peek(); xrec = (uint32_t)cur0; yrec |= 1u << 31;
            }
        }
    }
    const uint32_t y1 = y * pLL, y2 = y1 * pML;
    r.aLL = bLL + 2 * mulhi32(y, pLL);
    r.aML = bML + 2 * mulhi32(y1, pML);
    r.aOF = bOF + 2 * mulhi32(y2, pOF);
    cs.cur = cur0 - (int32_t)(a3 + nsum);
    return raw_pack(xrec, yrec);
}

// Stage A.  `mem` = this stream's kChainBytes of shared memory (256-byte aligned on the device).  Writes nseq RAW
// records at `out`.
// SIMT shape: the lanes of a warp run different blocks, so the loop must stay in lockstep or the warp
// degenerates into serial threads: single exit, the table build (data-dependent control flow) is
// fenced off with a warp barrier, and the loop runs a warp-uniform number of iterations (`bound` =
// the largest nseq - 1 among the lanes in `mask`), each lane masking itself out when its block is
// done.  The last sequence of a block updates no state; it is decoded after the loop, by all lanes at once.
// gtab / h: the block's tables and header from the table stage (seq_tables_thread).
//
// One iteration (chain_step) is straight-line code, ~55 instructions, and its critical path is
//   state -> LDS cell -> sum of the extra-bit counts -> funnel -> multiply-high -> next state:
// * the states are kept as the ADDRESSES of their cells (table base + 2 * state);
// * with J = ((baseline >> nb) << 1 | 1) << nb, 2 * baseline is J & (J - 1) and 2^nb is J & ~(J - 1), so the nb state
//   bits at the top of a window y are umulhi(y, 2^nb) and the window moves on by y * 2^nb: no bit counts, no variable
//   shifts; the bits consumed by the three updates together are log2 of the product of the three powers;
// * the window (x:xlo = the 64 bits under the cursor) does not depend on the cells, so its loads overlap theirs.
// A sequence whose extra bits exceed 32 (long lengths with a far offset: rare) hands stage B its bit cursor.
// The loop is latency-bound: an SM holds 82 chains whatever their arrangement in warps, so what counts is the length
// of one iteration of one chain, i.e. the instruction count and the dependent chain of chain_step.
FZ_HD int decode_sequences_chain(const Block& b, const uint8_t* gtab, const SeqJobHdr& h, uint8_t* mem,
                                 uint64_t* out, uint32_t bound, uint32_t mask)
{
    uint16_t* cLL = (uint16_t*)mem; uint16_t* cML = cLL + kChainCellsLL; uint16_t* cOF = cML + kChainCellsML;
    uint8_t* ring = mem + kChainCellBytes;
    int st = h.bad ? FZG_E_CORRUPT : 0;
    SeqCursor cs; cs.cur = 0; cs.lo = 0; cs.gbase = nullptr; cs.ring = sm_of(ring); cs.fa = -1; cs.fmin = 0;
    const sm_t tLL = sm_of(cLL), tML = sm_of(cML), tOF = sm_of(cOF);
    sm_t aLL = tLL, aML = tML, aOF = tOF;                             // addresses of the current cells
    // the block's chain cells: HBM (table stage) -> shared memory, the same 160 chunks for every lane
    for (uint32_t c = 0; c < kChainCellBytes; c += 16) ring_fetch_sm(tLL + c, gtab + c);
    ring_commit(); ring_wait<0>();
    if (!st) {
        if (h.bits_off > b.csize || cs.init(b.src + h.bits_off, b.csize - h.bits_off, ring) != 0) st = FZG_E_CORRUPT;
        if (!st) {
            aLL = tLL + 2 * cs.read(h.logLL);
            aOF = tOF + 2 * cs.read(h.logOF);
            aML = tML + 2 * cs.read(h.logML);
            if (cs.cur < cs.lo - 1) st = FZG_E_CORRUPT;
        }
    }
    const uint32_t nseq = b.nseq;
    const uint32_t live = st ? 0 : nseq - 1;            // this lane's trip count (nseq >= 1 for a sequence job)
#ifdef __CUDA_ARCH__
    const uint32_t common = __reduce_min_sync(mask, live) & ~3u;     // iterations every lane of the warp runs
#else
    const uint32_t common = live & ~3u;
#endif
    ChainRegs r{ aLL, aML, aOF, tLL, tML, tOF, (uint32_t)tLL + ((uint32_t)tLL << 12) };
    FZ_SYNCWARP(mask);
    // main loop: every lane is live, nothing is predicated; the ring is topped up every fourth sequence: <= 45 bytes are
    // consumed between two visits, so a chunk fetched at visit v (>= 240 bytes below the cursor of visit v - 1) is not
    // read before visit v + 3
    uint32_t i = 0;
    for (; i < common; i += 4, out += 4) {
        cs.fill();
        ring_commit(); ring_wait<3>();
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        uint64_t raw[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (uint32_t k = 0; k < 4; k++) raw[k] = chain_step<true>(r, cs, mask);
        st_rec4(out, raw[0], raw[1], raw[2], raw[3]);     // a scattered store costs the LSU one cycle per lane whatever its width
    }
    // the ragged end: lanes drop out one by one
    for (; i < bound; i += 4, out += 4) {
        if (i < live) cs.fill();
        ring_commit(); ring_wait<3>();
        for (uint32_t k = 0; k < 4; k++)
            if (i + k < live) out[k] = chain_step<false>(r, cs, 0);
    }
    out -= (bound + 3) & ~3u;
    ring_wait<0>();
    if (!st) {                                    // the last sequence: its three fields, no state update
        cs.fill(); ring_commit(); ring_wait<0>();
        const uint32_t co = sm_ld16(r.aOF), cl = sm_ld16(r.aLL), cm = sm_ld16(r.aML);
        const uint32_t ofb = co >> 10, a3 = ofb + (cl >> 10) + (cm >> 10);
        const uint32_t far = a3 > 32 ? 1u : 0u;
        out[nseq - 1] = raw_pack(far ? (uint32_t)cs.cur : cs.peek(), (uint32_t)r.aLL + ((uint32_t)r.aML << 12) + (ofb << 24) + (far << 31) - r.ypack);
        if (cs.cur - (int32_t)a3 != cs.lo - 1) st = FZG_E_CORRUPT;
    }
    return st;
}

// ------------------------------------------------------------------ XXH64 (RFC 8878 3.1.1: Content_Checksum)
constexpr uint64_t XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull,
                   XP4 = 0x85EBCA77C2B2AE63ull, XP5 = 0x27D4EB2F165667C5ull;
FZ_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
FZ_HD uint64_t xx_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
FZ_HD uint64_t xx_merge(uint64_t h, uint64_t v) { return (h ^ xx_round(0, v)) * XP1 + XP4; }
FZ_HD uint64_t rd64u(const uint8_t* p)
{
    if (((uintptr_t)p & 7) == 0) return *(const uint64_t*)p;
    uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v;
}
// tail: everything after the 32-byte stripes; h already merged (or seed + P5 for short inputs)
FZ_HD uint64_t xx_finish(uint64_t h, const uint8_t* p, uint64_t rem, uint64_t total)
{
    h += total;
    while (rem >= 8) { h ^= xx_round(0, rd64u(p)); h = rotl64(h, 27) * XP1 + XP4; p += 8; rem -= 8; }
    if (rem >= 4) { h ^= (uint64_t)rd32u(p) * XP1; h = rotl64(h, 23) * XP2 + XP3; p += 4; rem -= 4; }
    while (rem) { h ^= (uint64_t)(*p) * XP5; h = rotl64(h, 11) * XP1; p++; rem--; }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

}  // namespace fz
