/*
 * zstd_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product path (libfzgpu.so / the fuse-zstd_b200 package).
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg.
 *
 * What this restates
 * ------------------
 * fuse-zstd's decode hot path is one call,
 *     zstd::stream::copy_decode(source_file, target_file)   /root/reference/src/main.rs:463-467
 * whose arithmetic lives in a third-party dependency that is NOT vendored under
 * /root/reference: crate `zstd` 0.13.2 -> `zstd-safe` 7.2.1 -> `zstd-sys`
 * 2.0.13+zstd.1.5.6 (/root/reference/Cargo.lock:2371-2390), i.e. libzstd 1.5.6's
 * ZSTD_decompressStream.  This file restates the PUBLISHED algorithm (RFC 8878,
 * "Zstandard Compression and the application/zstd Media Type") as a plain,
 * byte-at-a-time C decoder, written from the format description -- it shares no
 * code with libzstd and none with the CUDA path.
 *
 * Parity pin
 * ----------
 * tests/test_oracle.py checks this decoder against
 *   (1) the golden vectors the reference's own tests hold for this path
 *       (tests/cmdline.rs:34-43, tests/convert.rs:16-43,54-98, tests/cmdline.rs:160-178;
 *       all Raw-block frames -- the reference never tests a Compressed block), and
 *   (2) frames produced by the system libzstd (1.5.5, the same C library family the
 *       reference links) for every block/literal/sequence mode, committed under
 *       tests/golden/ with the generating script, and
 *   (3) live differential runs against libzstd.so.1 when it is present at test time.
 * Status: Raw-block flow pinned by the reference's tests; Compressed-block
 * arithmetic pinned by libzstd outputs (the reference's tests leave it unpinned).
 *
 * Behavioural choices that mirror libzstd's streaming decoder (what copy_decode runs):
 *   - concatenated frames and skippable frames are consumed until EOF; empty input is OK;
 *   - window > 2^27 (incl. Single_Segment FCS > 2^27), reserved FHD bit, non-zero
 *     Dictionary_ID -> error;  block type 3 -> error;  FCS mismatch -> error;
 *   - XXH64 trailer is verified when Content_Checksum is set;
 *   - repeat-offset "rep0-1 == 0" is forced to 1 (libzstd does `temp += !temp`).
 */
#include "zstd_oracle.h"

#include <stdlib.h>
#include <string.h>

#define ZMAGIC 0xFD2FB528u
#define BLOCK_MAX (128u * 1024u)
#define WINDOW_MAX (1ull << 27)

#define MAX_LL 35
#define MAX_OF 31
#define MAX_ML 52
#define LL_LOG_MAX 9
#define OF_LOG_MAX 8
#define ML_LOG_MAX 9
#define HUF_LOG_MAX 12

static inline int highbit32(uint32_t v) { return 31 - __builtin_clz(v); }

/* ------------------------------------------------------------------ XXH64 */
#define P1 0x9E3779B185EBCA87ull
#define P2 0xC2B2AE3D27D4EB4Full
#define P3 0x165667B19E3779F9ull
#define P4 0x85EBCA77C2B2AE63ull
#define P5 0x27D4EB2F165667C5ull
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xxround(uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; }
static inline uint64_t xxmerge(uint64_t h, uint64_t v) { return (h ^ xxround(0, v)) * P1 + P4; }

uint64_t fzo_xxh64(const void* data, size_t len, uint64_t seed)
{
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = xxround(v1, rd64(p));      v2 = xxround(v2, rd64(p + 8));
            v3 = xxround(v3, rd64(p + 16)); v4 = xxround(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxmerge(h, v1); h = xxmerge(h, v2); h = xxmerge(h, v3); h = xxmerge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= xxround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end)    { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end)      { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------- forward bit reader (LSB first) */
typedef struct { const uint8_t* p; size_t n; size_t bitpos; } fbr_t;
static inline uint32_t fbr_peek(const fbr_t* b, int nb)  /* nb <= 25; zero padded past the end */
{
    size_t byte = b->bitpos >> 3; int sh = (int)(b->bitpos & 7);
    uint64_t w = 0;
    for (int i = 0; i < 5; i++) if (byte + i < b->n) w |= (uint64_t)b->p[byte + i] << (8 * i);
    return (uint32_t)((w >> sh) & ((1ull << nb) - 1));
}

/* ---------------------------------------------- backward bit reader (RFC 8878 4.1) */
typedef struct { const uint8_t* p; size_t n; int64_t left; } bbr_t;  /* left = unread bits */
static int bbr_init(bbr_t* b, const uint8_t* p, size_t n)
{
    if (n == 0 || p[n - 1] == 0) return -1;
    b->p = p; b->n = n;
    b->left = (int64_t)(n - 1) * 8 + highbit32(p[n - 1]);
    return 0;
}
/* next nb bits (nb <= 32) as a number, first-read bit most significant; bits below
 * the start of the stream read as zero */
static uint64_t bbr_peek(const bbr_t* b, int nb)
{
    if (nb == 0) return 0;
    int64_t lo = b->left - nb;
    int shift_up = 0;
    if (lo < 0) { shift_up = (int)(-lo); lo = 0; if (shift_up >= nb) return 0; }
    int take = nb - shift_up;
    size_t byte = (size_t)(lo >> 3); int sh = (int)(lo & 7);
    uint64_t w = 0;
    for (int i = 0; i < 8; i++) if (byte + i < b->n) w |= (uint64_t)b->p[byte + i] << (8 * i);
    uint64_t v = (w >> sh) & ((take >= 64) ? ~0ull : ((1ull << take) - 1));
    return v << shift_up;
}
static inline uint64_t bbr_read(bbr_t* b, int nb) { uint64_t v = bbr_peek(b, nb); b->left -= nb; return v; }

/* ------------------------------------------------------------- FSE tables */
typedef struct { uint16_t base; uint8_t nb; uint8_t sym; } fse_cell_t;
typedef struct { int log; fse_cell_t cell[512]; } fse_table_t;

/* RFC 8878 4.1.1: read a normalised distribution.  Returns bytes consumed or -1. */
static int fse_read_ncount(const uint8_t* p, size_t n, int max_sym, int max_log,
                           int16_t* norm, int* n_sym, int* log_out)
{
    if (n == 0) return -1;
    fbr_t br = { p, n, 0 };
    int log = (int)fbr_peek(&br, 4) + 5; br.bitpos += 4;
    if (log > max_log) return -1;
    int remaining = 1 << log;
    int sym = 0;
    while (remaining > 0 && sym <= max_sym) {
        int bits = highbit32((uint32_t)remaining + 1) + 1;
        uint32_t v = fbr_peek(&br, bits);
        uint32_t lower = (1u << (bits - 1)) - 1;
        uint32_t thr = (1u << bits) - 1 - ((uint32_t)remaining + 1);
        if ((v & lower) < thr) { v &= lower; br.bitpos += bits - 1; }
        else { if (v > lower) v -= thr; br.bitpos += bits; }
        int prob = (int)v - 1;
        remaining -= (prob < 0) ? 1 : prob;
        norm[sym++] = (int16_t)prob;
        if (prob == 0) {
            for (;;) {
                uint32_t rep = fbr_peek(&br, 2); br.bitpos += 2;
                for (uint32_t i = 0; i < rep; i++) { if (sym > max_sym) return -1; norm[sym++] = 0; }
                if (rep != 3) break;
            }
        }
        if (br.bitpos > n * 8) return -1;
    }
    if (remaining != 0) return -1;
    if (br.bitpos > n * 8) return -1;
    *n_sym = sym; *log_out = log;
    return (int)((br.bitpos + 7) >> 3);
}

/* RFC 8878 4.1.1 "from normalized distribution to decoding tables" */
static int fse_build(fse_table_t* t, const int16_t* norm, int n_sym, int log)
{
    int size = 1 << log, high = size - 1;
    uint16_t next[64];
    for (int s = 0; s < n_sym; s++) {
        if (norm[s] == -1) { t->cell[high--].sym = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s < n_sym; s++) {
        for (int i = 0; i < norm[s]; i++) {
            t->cell[pos].sym = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    if (pos != 0) return -1;
    for (int u = 0; u < size; u++) {
        int s = t->cell[u].sym;
        uint32_t nx = next[s]++;
        int nb = log - highbit32(nx);
        t->cell[u].nb = (uint8_t)nb;
        t->cell[u].base = (uint16_t)((nx << nb) - size);
    }
    t->log = log;
    return 0;
}
static void fse_build_rle(fse_table_t* t, int sym)
{
    t->log = 0; t->cell[0].sym = (uint8_t)sym; t->cell[0].nb = 0; t->cell[0].base = 0;
}

static const int16_t LL_DEFAULT[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
static const int16_t OF_DEFAULT[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };
static const int16_t ML_DEFAULT[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                                        1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
static const uint32_t LL_BASE[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,
                                      1024,2048,4096,8192,16384,32768,65536 };
static const uint8_t LL_BITS[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
static const uint32_t ML_BASE[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,
                                      32,33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,
                                      32771,65539 };
static const uint8_t ML_BITS[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,
                                     4,4,5,7,8,9,10,11,12,13,14,15,16 };

/* ------------------------------------------------------------- Huffman */
typedef struct { int log; uint8_t sym[1 << HUF_LOG_MAX]; uint8_t nb[1 << HUF_LOG_MAX]; } huf_table_t;

/* RFC 8878 4.2.1: tree description -> weights -> decode table.  Returns bytes consumed or -1. */
static int huf_read_table(huf_table_t* ht, const uint8_t* p, size_t n)
{
    uint8_t w[256];
    int nw = 0;
    size_t used;
    if (n < 1) return -1;
    int h = p[0];
    if (h >= 128) {                     /* direct 4-bit weights */
        nw = h - 127;
        size_t bytes = (size_t)(nw + 1) / 2;
        if (1 + bytes > n) return -1;
        for (int i = 0; i < nw; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        used = 1 + bytes;
    } else {                            /* FSE-compressed weights, 2 interleaved states */
        if (h == 0 || (size_t)h + 1 > n) return -1;
        int16_t norm[16]; int ns, log;
        int hb = fse_read_ncount(p + 1, (size_t)h, 12, 6, norm, &ns, &log);
        if (hb < 0 || hb >= h) return -1;
        fse_table_t* ft = (fse_table_t*)malloc(sizeof(fse_table_t));
        if (fse_build(ft, norm, ns, log) != 0) { free(ft); return -1; }
        bbr_t br;
        if (bbr_init(&br, p + 1 + hb, (size_t)h - hb) != 0) { free(ft); return -1; }
        uint32_t s1 = (uint32_t)bbr_read(&br, log), s2 = (uint32_t)bbr_read(&br, log);
        if (br.left < 0) { free(ft); return -1; }
        for (;;) {
            if (nw > 253) { free(ft); return -1; }
            w[nw++] = ft->cell[s1].sym;
            s1 = ft->cell[s1].base + (uint32_t)bbr_read(&br, ft->cell[s1].nb);
            if (br.left < 0) { w[nw++] = ft->cell[s2].sym; break; }
            if (nw > 253) { free(ft); return -1; }
            w[nw++] = ft->cell[s2].sym;
            s2 = ft->cell[s2].base + (uint32_t)bbr_read(&br, ft->cell[s2].nb);
            if (br.left < 0) { w[nw++] = ft->cell[s1].sym; break; }
        }
        free(ft);
        used = 1 + (size_t)h;
    }
    /* implicit last weight */
    uint32_t sum = 0; int rank[HUF_LOG_MAX + 2]; memset(rank, 0, sizeof rank);
    for (int i = 0; i < nw; i++) {
        if (w[i] > HUF_LOG_MAX) return -1;
        if (w[i]) sum += 1u << (w[i] - 1);
        rank[w[i]]++;
    }
    if (sum == 0) return -1;
    int log = highbit32(sum) + 1;
    if (log > HUF_LOG_MAX) return -1;
    uint32_t left = (1u << log) - sum;
    if (left & (left - 1)) return -1;               /* must be a power of two */
    int last = highbit32(left) + 1;
    w[nw++] = (uint8_t)last; rank[last]++;
    if (rank[1] < 2 || (rank[1] & 1)) return -1;     /* libzstd: by construction */
    /* canonical fill: weight 1 (longest codes) first, ascending symbol within a weight */
    uint32_t start[HUF_LOG_MAX + 2]; uint32_t cur = 0;
    for (int k = 1; k <= log; k++) { start[k] = cur; cur += (uint32_t)rank[k] << (k - 1); }
    for (int s = 0; s < nw; s++) {
        int wt = w[s]; if (!wt) continue;
        uint32_t len = 1u << (wt - 1);
        for (uint32_t i = 0; i < len; i++) { ht->sym[start[wt] + i] = (uint8_t)s; ht->nb[start[wt] + i] = (uint8_t)(log + 1 - wt); }
        start[wt] += len;
    }
    ht->log = log;
    return (int)used;
}

static int huf_decode_stream(const huf_table_t* ht, const uint8_t* p, size_t n, uint8_t* out, size_t n_out)
{
    bbr_t br;
    if (bbr_init(&br, p, n) != 0) return -1;
    for (size_t i = 0; i < n_out; i++) {
        uint32_t idx = (uint32_t)bbr_peek(&br, ht->log);
        out[i] = ht->sym[idx];
        br.left -= ht->nb[idx];
    }
    return br.left == 0 ? 0 : -1;
}

/* ------------------------------------------------------------- frame state */
typedef struct {
    huf_table_t huf; int huf_valid;
    fse_table_t ll, of, ml; int ll_valid, of_valid, ml_valid;
    uint32_t rep[3];
    uint8_t lit[BLOCK_MAX];
} fctx_t;

static void trace_block(fzo_trace_t* tr, int type, uint32_t nseq, uint32_t litsize, int littype, int modes)
{
    if (!tr) return;
    if (tr->blocks_len < tr->blocks_cap) {
        if (tr->block_littype) tr->block_littype[tr->blocks_len] = (uint8_t)littype;
        if (tr->block_modes) tr->block_modes[tr->blocks_len] = (uint8_t)modes;
        if (tr->block_nseq) tr->block_nseq[tr->blocks_len] = nseq;
        if (tr->block_litsize) tr->block_litsize[tr->blocks_len] = litsize;
        if (tr->block_type) tr->block_type[tr->blocks_len] = (uint8_t)type;
    }
    tr->blocks_len++;
}

/* RFC 8878 3.1.1.3.1 literals section; returns bytes consumed or -1 */
static int64_t decode_literals(fctx_t* c, const uint8_t* p, size_t n, size_t* lit_size, uint32_t block_max)
{
    if (n < 1) return -1;
    int type = p[0] & 3, sf = (p[0] >> 2) & 3;
    size_t hs, regen, comp = 0; int streams = 1;
    if (type < 2) {                                     /* Raw / RLE */
        if (sf == 0 || sf == 2) { hs = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return -1; hs = 2; regen = (p[0] >> 4) | ((size_t)p[1] << 4); }
        else { if (n < 3) return -1; hs = 3; regen = (p[0] >> 4) | ((size_t)p[1] << 4) | ((size_t)p[2] << 12); }
        if (regen > block_max) return -1;
        if (type == 0) {
            if (hs + regen > n) return -1;
            memcpy(c->lit, p + hs, regen); *lit_size = regen; return (int64_t)(hs + regen);
        }
        if (hs + 1 > n) return -1;
        memset(c->lit, p[hs], regen); *lit_size = regen; return (int64_t)(hs + 1);
    }
    if (sf == 0 || sf == 1) {
        if (n < 3) return -1;
        uint32_t v = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        hs = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; streams = sf ? 4 : 1;
    } else if (sf == 2) {
        if (n < 4) return -1;
        uint32_t v = rd32(p);
        hs = 4; regen = (v >> 4) & 0x3FFF; comp = (v >> 18) & 0x3FFF; streams = 4;
    } else {
        if (n < 5) return -1;
        uint64_t v = (uint64_t)rd32(p) | ((uint64_t)p[4] << 32);
        hs = 5; regen = (size_t)((v >> 4) & 0x3FFFF); comp = (size_t)((v >> 22) & 0x3FFFF); streams = 4;
    }
    if (regen > block_max || regen == 0) return -1;
    if (streams == 4 && regen < 6) return -1;
    if (hs + comp > n) return -1;
    const uint8_t* q = p + hs; size_t qn = comp;
    if (type == 2) {
        int used = huf_read_table(&c->huf, q, qn);
        if (used < 0) return -1;
        c->huf_valid = 1; q += used; qn -= (size_t)used;
    } else if (!c->huf_valid) return -1;
    if (streams == 1) {
        if (huf_decode_stream(&c->huf, q, qn, c->lit, regen) != 0) return -1;
    } else {
        if (qn < 10) return -1;
        size_t l1 = q[0] | ((size_t)q[1] << 8), l2 = q[2] | ((size_t)q[3] << 8), l3 = q[4] | ((size_t)q[5] << 8);
        if (6 + l1 + l2 + l3 > qn) return -1;
        size_t l4 = qn - 6 - l1 - l2 - l3;
        size_t seg = (regen + 3) / 4;
        if (seg * 3 > regen) return -1;
        const uint8_t* s = q + 6;
        if (huf_decode_stream(&c->huf, s, l1, c->lit, seg) != 0) return -1;
        if (huf_decode_stream(&c->huf, s + l1, l2, c->lit + seg, seg) != 0) return -1;
        if (huf_decode_stream(&c->huf, s + l1 + l2, l3, c->lit + 2 * seg, seg) != 0) return -1;
        if (huf_decode_stream(&c->huf, s + l1 + l2 + l3, l4, c->lit + 3 * seg, regen - 3 * seg) != 0) return -1;
    }
    *lit_size = regen;
    return (int64_t)(hs + comp);
}

/* one of LL/OF/ML: build the decode table according to its mode; returns bytes consumed or -1 */
static int seq_table(fse_table_t* t, int* valid, int mode, const uint8_t* p, size_t n,
                     const int16_t* def, int def_n, int def_log, int max_sym, int max_log)
{
    int16_t norm[64]; int ns, log, used;
    switch (mode) {
    case 0: if (fse_build(t, def, def_n, def_log) != 0) return -1; *valid = 1; return 0;
    case 1: if (n < 1 || p[0] > max_sym) return -1; fse_build_rle(t, p[0]); *valid = 1; return 1;
    case 2:
        used = fse_read_ncount(p, n, max_sym, max_log, norm, &ns, &log);
        if (used < 0) return -1;
        if (fse_build(t, norm, ns, log) != 0) return -1;
        *valid = 1; return used;
    default: return *valid ? 0 : -1;
    }
}

/* Compressed block: literals + sequences + execution.  out_base = frame start. */
static int decode_compressed_block(fctx_t* c, const uint8_t* p, size_t n, uint8_t* out_base, size_t out_pos,
                                   size_t out_cap, uint32_t block_max, size_t* produced, fzo_trace_t* tr)
{
    if (n < 2) return FZO_E_CORRUPT;
    size_t lit_size = 0;
    int64_t lc = decode_literals(c, p, n, &lit_size, block_max);
    if (lc < 0) return FZO_E_CORRUPT;
    if (tr && tr->literals) {
        size_t room = tr->literals_cap > tr->literals_len ? tr->literals_cap - tr->literals_len : 0;
        memcpy(tr->literals + tr->literals_len, c->lit, lit_size < room ? lit_size : room);
    }
    if (tr) tr->literals_len += lit_size;
    const uint8_t* q = p + lc; size_t qn = n - (size_t)lc;
    if (qn < 1) return FZO_E_CORRUPT;
    uint32_t nseq = q[0]; size_t hs = 1;
    if (nseq >= 128) {
        if (nseq == 255) { if (qn < 3) return FZO_E_CORRUPT; nseq = q[1] + ((uint32_t)q[2] << 8) + 0x7F00; hs = 3; }
        else { if (qn < 2) return FZO_E_CORRUPT; nseq = ((nseq - 128) << 8) + q[1]; hs = 2; }
    }
    {
        int lt = p[0] & 3, sf = (p[0] >> 2) & 3, info = lt;
        if (lt >= 2 && sf != 0) info |= 4;
        if (lt == 2) { size_t lh = sf < 2 ? 3 : (sf == 2 ? 4 : 5); if (p[lh] < 128) info |= 8; }
        int md = 0;
        if (nseq && qn >= hs + 1) md = q[hs];
        trace_block(tr, 2, nseq, (uint32_t)lit_size, info, md);
    }
    uint8_t* op = out_base + out_pos;
    size_t room = out_cap - out_pos;
    if (nseq == 0) {
        if (hs != qn) return FZO_E_CORRUPT;
        if (lit_size > block_max) return FZO_E_CORRUPT;
        if (lit_size > room) return FZO_E_DSTSIZE;
        memcpy(op, c->lit, lit_size); *produced = lit_size; return FZO_OK;
    }
    if (qn < hs + 1) return FZO_E_CORRUPT;
    int modes = q[hs];
    if (modes & 3) return FZO_E_CORRUPT;
    q += hs + 1; qn -= hs + 1;
    int u;
    u = seq_table(&c->ll, &c->ll_valid, (modes >> 6) & 3, q, qn, LL_DEFAULT, 36, 6, MAX_LL, LL_LOG_MAX);
    if (u < 0) return FZO_E_CORRUPT;
    q += u; qn -= (size_t)u;
    u = seq_table(&c->of, &c->of_valid, (modes >> 4) & 3, q, qn, OF_DEFAULT, 29, 5, MAX_OF, OF_LOG_MAX);
    if (u < 0) return FZO_E_CORRUPT;
    q += u; qn -= (size_t)u;
    u = seq_table(&c->ml, &c->ml_valid, (modes >> 2) & 3, q, qn, ML_DEFAULT, 53, 6, MAX_ML, ML_LOG_MAX);
    if (u < 0) return FZO_E_CORRUPT;
    q += u; qn -= (size_t)u;

    bbr_t br;
    if (bbr_init(&br, q, qn) != 0) return FZO_E_CORRUPT;
    uint32_t sll = (uint32_t)bbr_read(&br, c->ll.log);
    uint32_t sof = (uint32_t)bbr_read(&br, c->of.log);
    uint32_t sml = (uint32_t)bbr_read(&br, c->ml.log);
    if (br.left < 0) return FZO_E_CORRUPT;

    size_t lit_pos = 0, done = 0;
    for (uint32_t i = 0; i < nseq; i++) {
        int ofc = c->of.cell[sof].sym, mlc = c->ml.cell[sml].sym, llc = c->ll.cell[sll].sym;
        if (ofc > MAX_OF || mlc > MAX_ML || llc > MAX_LL) return FZO_E_CORRUPT;
        uint32_t ofv = (1u << ofc) + (uint32_t)bbr_read(&br, ofc);
        uint32_t ml = ML_BASE[mlc] + (uint32_t)bbr_read(&br, ML_BITS[mlc]);
        uint32_t ll = LL_BASE[llc] + (uint32_t)bbr_read(&br, LL_BITS[llc]);
        if (i + 1 < nseq) {
            sll = c->ll.cell[sll].base + (uint32_t)bbr_read(&br, c->ll.cell[sll].nb);
            sml = c->ml.cell[sml].base + (uint32_t)bbr_read(&br, c->ml.cell[sml].nb);
            sof = c->of.cell[sof].base + (uint32_t)bbr_read(&br, c->of.cell[sof].nb);
        }
        if (br.left < 0) return FZO_E_CORRUPT;
        /* repeat offsets, RFC 8878 3.1.1.5 */
        uint32_t off;
        if (ofv > 3) { off = ofv - 3; c->rep[2] = c->rep[1]; c->rep[1] = c->rep[0]; c->rep[0] = off; }
        else {
            uint32_t idx = ofv - 1 + (ll == 0);
            if (idx == 0) off = c->rep[0];
            else {
                off = (idx == 3) ? c->rep[0] - 1 : c->rep[idx];
                if (off == 0) off = 1;
                if (idx != 1) c->rep[2] = c->rep[1];
                c->rep[1] = c->rep[0]; c->rep[0] = off;
            }
        }
        if (tr) {
            if (tr->seqs && tr->seqs_len < tr->seqs_cap) {
                fzo_seq_t s = { ll, ml, off, ofv }; tr->seqs[tr->seqs_len] = s;
            }
            tr->seqs_len++;
        }
        /* execute */
        if (ll > lit_size - lit_pos) return FZO_E_CORRUPT;
        if (done + ll + ml > block_max) return FZO_E_CORRUPT;
        if (done + ll + ml > room) return FZO_E_DSTSIZE;
        memcpy(op + done, c->lit + lit_pos, ll); lit_pos += ll; done += ll;
        if (off > out_pos + done) return FZO_E_CORRUPT;
        { uint8_t* d = op + done; const uint8_t* s = d - off; for (uint32_t k = 0; k < ml; k++) d[k] = s[k]; }
        done += ml;
    }
    if (br.left != 0) return FZO_E_CORRUPT;
    size_t rest = lit_size - lit_pos;
    if (done + rest > block_max) return FZO_E_CORRUPT;
    if (done + rest > room) return FZO_E_DSTSIZE;
    memcpy(op + done, c->lit + lit_pos, rest); done += rest;
    *produced = done;
    return FZO_OK;
}

/* RFC 8878 3.1.1.1 frame header.  Returns header size or negative FZO code. */
typedef struct { uint64_t fcs; int has_fcs; uint64_t window; int checksum; size_t hsize; } fhdr_t;
static int parse_frame_header(const uint8_t* p, size_t n, fhdr_t* h)
{
    if (n < 5) return FZO_E_TRUNCATED;
    int fhd = p[4];
    int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    if (fhd & 8) return FZO_E_UNSUPPORTED;
    h->checksum = (fhd >> 2) & 1;
    size_t pos = 5;
    h->window = 0;
    if (!single) {
        if (n < pos + 1) return FZO_E_TRUNCATED;
        int b = p[pos++]; int exp = b >> 3, mant = b & 7;
        uint64_t base = 1ull << (10 + exp);
        h->window = base + (base >> 3) * (uint64_t)mant;
    }
    static const int did_bytes[4] = { 0, 1, 2, 4 };
    int db = did_bytes[did_flag];
    if (n < pos + (size_t)db) return FZO_E_TRUNCATED;
    uint32_t did = 0;
    for (int i = 0; i < db; i++) did |= (uint32_t)p[pos + i] << (8 * i);
    pos += (size_t)db;
    if (did != 0) return FZO_E_UNSUPPORTED;
    int fb = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (n < pos + (size_t)fb) return FZO_E_TRUNCATED;
    h->has_fcs = fb != 0; h->fcs = 0;
    for (int i = 0; i < fb; i++) h->fcs |= (uint64_t)p[pos + i] << (8 * i);
    if (fb == 2) h->fcs += 256;
    pos += (size_t)fb;
    if (single) h->window = h->fcs;
    if (h->window > WINDOW_MAX) return FZO_E_UNSUPPORTED;
    h->hsize = pos;
    return FZO_OK;
}

int fzo_decode_trace(const void* src_, size_t src_len, void* dst_, size_t dst_cap, size_t* out_len, fzo_trace_t* tr)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    size_t ip = 0, op = 0;
    int rc = FZO_OK;
    fctx_t* c = (fctx_t*)malloc(sizeof(fctx_t));
    if (tr) { tr->literals_len = 0; tr->seqs_len = 0; tr->blocks_len = 0; }
    while (ip < src_len) {
        if (src_len - ip < 4) { rc = FZO_E_TRUNCATED; break; }
        uint32_t magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {             /* skippable frame */
            if (src_len - ip < 8) { rc = FZO_E_TRUNCATED; break; }
            uint64_t sz = rd32(src + ip + 4);
            if (src_len - ip - 8 < sz) { rc = FZO_E_TRUNCATED; break; }
            ip += 8 + (size_t)sz; continue;
        }
        if (magic != ZMAGIC) { rc = FZO_E_MAGIC; break; }
        fhdr_t h;
        rc = parse_frame_header(src + ip, src_len - ip, &h);
        if (rc) break;
        ip += h.hsize;
        uint32_t block_max = h.window < BLOCK_MAX ? (uint32_t)h.window : BLOCK_MAX;
        size_t frame_start = op;
        c->huf_valid = c->ll_valid = c->of_valid = c->ml_valid = 0;
        c->rep[0] = 1; c->rep[1] = 4; c->rep[2] = 8;
        for (;;) {
            if (src_len - ip < 3) { rc = FZO_E_TRUNCATED; break; }
            uint32_t bh = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            ip += 3;
            int last = bh & 1, type = (bh >> 1) & 3; uint32_t bsize = bh >> 3;
            if (type == 3) { rc = FZO_E_CORRUPT; break; }
            if (bsize > block_max) { rc = FZO_E_CORRUPT; break; }
            if (type == 0) {
                if (src_len - ip < bsize) { rc = FZO_E_TRUNCATED; break; }
                if (dst_cap - op < bsize) { rc = FZO_E_DSTSIZE; break; }
                memcpy(dst + op, src + ip, bsize); ip += bsize; op += bsize;
                trace_block(tr, 0, 0, 0, 0, 0);
            } else if (type == 1) {
                if (src_len - ip < 1) { rc = FZO_E_TRUNCATED; break; }
                if (dst_cap - op < bsize) { rc = FZO_E_DSTSIZE; break; }
                memset(dst + op, src[ip], bsize); ip += 1; op += bsize;
                trace_block(tr, 1, 0, 0, 0, 0);
            } else {
                if (src_len - ip < bsize) { rc = FZO_E_TRUNCATED; break; }
                size_t produced = 0;
                rc = decode_compressed_block(c, src + ip, bsize, dst + frame_start, op - frame_start,
                                             dst_cap - frame_start, block_max, &produced, tr);
                if (rc) break;
                ip += bsize; op += produced;
            }
            if (last) break;
        }
        if (rc) break;
        if (h.has_fcs && (uint64_t)(op - frame_start) != h.fcs) { rc = FZO_E_FCS; break; }
        if (h.checksum) {
            if (src_len - ip < 4) { rc = FZO_E_TRUNCATED; break; }
            uint32_t want = rd32(src + ip); ip += 4;
            if ((uint32_t)fzo_xxh64(dst + frame_start, op - frame_start, 0) != want) { rc = FZO_E_CHECKSUM; break; }
        }
    }
    free(c);
    if (out_len) *out_len = op;
    return rc;
}

int fzo_decode(const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* out_len)
{
    return fzo_decode_trace(src, src_len, dst, dst_cap, out_len, NULL);
}

int fzo_frame_info(const void* src_, size_t src_len, uint64_t* content_size, uint64_t* n_frames)
{
    const uint8_t* src = (const uint8_t*)src_;
    size_t ip = 0; uint64_t total = 0, frames = 0; int unknown = 0;
    while (ip < src_len) {
        if (src_len - ip < 4) return FZO_E_TRUNCATED;
        uint32_t magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (src_len - ip < 8) return FZO_E_TRUNCATED;
            uint64_t sz = rd32(src + ip + 4);
            if (src_len - ip - 8 < sz) return FZO_E_TRUNCATED;
            ip += 8 + (size_t)sz; continue;
        }
        if (magic != ZMAGIC) return FZO_E_MAGIC;
        fhdr_t h; int rc = parse_frame_header(src + ip, src_len - ip, &h);
        if (rc) return rc;
        ip += h.hsize;
        if (h.has_fcs) total += h.fcs; else unknown = 1;
        for (;;) {
            if (src_len - ip < 3) return FZO_E_TRUNCATED;
            uint32_t bh = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            ip += 3;
            int type = (bh >> 1) & 3; uint32_t bsize = bh >> 3;
            if (type == 3) return FZO_E_CORRUPT;
            size_t adv = type == 1 ? 1 : bsize;
            if (src_len - ip < adv) return FZO_E_TRUNCATED;
            ip += adv;
            if (bh & 1) break;
        }
        if (h.checksum) { if (src_len - ip < 4) return FZO_E_TRUNCATED; ip += 4; }
        frames++;
    }
    if (content_size) *content_size = unknown ? UINT64_MAX : total;
    if (n_frames) *n_frames = frames;
    return FZO_OK;
}
