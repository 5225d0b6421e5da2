// Analysis tool (not product, not oracle): CPU model of k_enc_match's parse -- 32 positions per step, candidates from the table as of the
// step's start + the nearest earlier position of the step with the same hash, one-step lazy selection, repeat-offset codes -- with an
// entropy cost estimate of the resulting block, for trying matcher variants without a GPU (profiles/r01_notes.md, session 4).
// build: gcc -O2 -o /tmp/enc_model tools/enc_matcher_model.c fuse-zstd_b200/csrc/corpus_gen.c -lm -lpthread
// run:   /tmp/enc_model hashlog 13 hbytes 5 long 1 longlog 15 maxoff 65535     (the shipped level-3 configuration: predicts 2.690, measured 2.629)
//        /tmp/enc_model hashlog 13 hbytes 4 maxoff 65535                          (levels 1-2: predicts 2.449, measured 2.396)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
extern void fzc_generate(uint64_t idx, void* out, size_t n);
static int LOOK = 1; static int HASHLOG = 13, MINMATCH = 4, LAZY = 1, HBYTES = 4, REPCHK = 0, LONGTAB = 0, LONGLOG = 12, MAXOFF = 65535;
static uint64_t rd8(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t hash_of(uint64_t v, int bytes, int log) {
    if (bytes == 4) return ((uint32_t)v * 2654435761u) >> (32 - log);
    uint64_t m = bytes >= 8 ? v : (v << (64 - 8 * bytes));
    return (uint32_t)((m * 0x9E3779B185EBCA87ull) >> (64 - log));
}
static uint32_t mlen(const uint8_t* a, const uint8_t* b, uint32_t max) { uint32_t l = 0; while (l < max && a[l] == b[l]) l++; return l; }
typedef struct { uint32_t ll, ml, off, ov; } Seq;
static double ent(const uint32_t* h, int n) { double t = 0, s = 0; for (int i = 0; i < n; i++) t += h[i]; if (t == 0) return 0; for (int i = 0; i < n; i++) if (h[i]) s -= h[i] * log2(h[i] / t); return s; }
static int hb(uint32_t v) { return 31 - __builtin_clz(v); }
static int ll_code(uint32_t ll) { static const uint8_t t[64] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,16,17,17,18,18,19,19,20,20,20,20,21,21,21,21,22,22,22,22,22,22,22,22,23,23,23,23,23,23,23,23,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24,24}; return ll > 63 ? hb(ll) + 19 : t[ll]; }
static int ml_code(uint32_t mlb) { static const uint8_t t[128] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,32,33,33,34,34,35,35,36,36,36,36,37,37,37,37,38,38,38,38,38,38,38,38,39,39,39,39,39,39,39,39,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42}; return mlb > 127 ? hb(mlb) + 36 : t[mlb]; }
static const uint8_t LLB[36] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16};
static const uint8_t MLB[53] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16};
// returns estimated compressed bytes for the chunk
static double parse_chunk(const uint8_t* src, uint32_t size, uint64_t* tot_seq, uint64_t* tot_lit, uint64_t* tot_rep)
{
    static uint32_t table[1 << 17]; static uint32_t ltab[1 << 17];
    memset(table, 0, sizeof table); memset(ltab, 0, sizeof ltab);
    Seq* seqs = malloc(sizeof(Seq) * (size / 3 + 8)); uint32_t nseq = 0;
    uint8_t* lit = malloc(size + 8); uint32_t nlit = 0;
    uint32_t anchor = 0, cur = 0, rep0 = 1, rep1 = 4, rep2 = 8;
    const uint32_t limit = size >= 16 ? size - 12 : 0;
    for (uint32_t base = 0; base < limit; base += 32) {
        uint32_t len[32], off[32]; int isrep[32];
        uint32_t hs[32], hl[32];
        for (int l = 0; l < 32; l++) { uint32_t p = base + l; if (p < limit) { uint64_t v = rd8(src + p); hs[l] = hash_of(v, HBYTES, HASHLOG); hl[l] = hash_of(v, 8, LONGLOG); } }
        for (int l = 0; l < 32; l++) {
            uint32_t p = base + l; len[l] = 0; off[l] = 0; isrep[l] = 0;
            if (p >= limit) continue;
            int32_t cand = -1;
            for (int k = l - 1; k >= 0; k--) if (hs[k] == hs[l]) { cand = base + k; break; }
            if (cand < 0) { int32_t c = (int32_t)table[hs[l]] - 1; cand = c; }
            uint32_t best = 0, boff = 0;
            if (cand >= 0 && p - cand <= (uint32_t)MAXOFF) { uint32_t m = mlen(src + cand, src + p, size - p); if (m >= 4) { best = m; boff = p - cand; } }
            if (LONGTAB) {
                int32_t c2 = -1;
                for (int k = l - 1; k >= 0; k--) if (hl[k] == hl[l]) { c2 = base + k; break; }
                if (c2 < 0) { int32_t c = (int32_t)ltab[hl[l]] - 1; c2 = c; }
                if (c2 >= 0 && p - c2 <= (uint32_t)MAXOFF) { uint32_t m = mlen(src + c2, src + p, size - p); if (m > best) { best = m; boff = p - c2; } }
            }
            if (REPCHK && p >= rep0) { uint32_t m = mlen(src + p - rep0, src + p, size - p); if (m >= 4 && m + REPCHK > best) { best = m; boff = rep0; isrep[l] = 1; } }
            len[l] = best; off[l] = boff;
        }
        for (int l = 0; l < 32; l++) { uint32_t p = base + l; if (p < limit) { table[hs[l]] = p + 1; ltab[hl[l]] = p + 1; } }
        for (int l = 0; l < 32; l++) {
            uint32_t pl = base + l;
            if (len[l] < (uint32_t)MINMATCH || pl < cur) continue;
            if (LAZY) { int skip = 0; for (int k = 1; k <= LOOK && l + k < 32; k++) if (len[l + k] >= (uint32_t)MINMATCH && len[l + k] > len[l] + (k - 1)) { skip = 1; break; } if (skip) continue; }
            uint32_t ll = pl - anchor, o = off[l], ov = o + 3;
            memcpy(lit + nlit, src + anchor, ll); nlit += ll;
            uint32_t r0 = rep0, r1 = rep1, r2 = rep2;
            if (ll) { if (o == r0) ov = 1; else if (o == r1) { ov = 2; rep0 = r1; rep1 = r0; } else { if (o == r2) ov = 3; rep0 = o; rep1 = r0; rep2 = r1; } }
            else { if (o == r1) { ov = 1; rep0 = r1; rep1 = r0; } else { if (o == r2) ov = 2; else if (o == r0 - 1 && o) ov = 3; rep0 = o; rep1 = r0; rep2 = r1; } }
            seqs[nseq++] = (Seq){ ll, len[l], o, ov };
            anchor = cur = pl + len[l];
        }
    }
    memcpy(lit + nlit, src + anchor, size - anchor); nlit += size - anchor;
    // cost estimate
    uint32_t hlit[256] = {0}, hll[36] = {0}, hml[53] = {0}, hof[32] = {0}; double extra = 0; uint64_t reps = 0;
    for (uint32_t i = 0; i < nlit; i++) hlit[lit[i]]++;
    for (uint32_t i = 0; i < nseq; i++) { int lc = ll_code(seqs[i].ll), mc = ml_code(seqs[i].ml - 3), oc = hb(seqs[i].ov); hll[lc]++; hml[mc]++; hof[oc]++; extra += LLB[lc] + MLB[mc] + oc; if (seqs[i].ov <= 3) reps++; }
    double bits = ent(hlit, 256) + ent(hll, 36) + ent(hml, 53) + ent(hof, 32) + extra;
    *tot_seq += nseq; *tot_lit += nlit; *tot_rep += reps;
    free(seqs); free(lit);
    return bits / 8 + 60;   // headers / table descriptions
}
int main(int argc, char** argv)
{
    for (int i = 1; i + 1 < argc; i += 2) {
        int v = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "hashlog")) HASHLOG = v; else if (!strcmp(argv[i], "minmatch")) MINMATCH = v; else if (!strcmp(argv[i], "lazy")) LAZY = v;
        else if (!strcmp(argv[i], "hbytes")) HBYTES = v; else if (!strcmp(argv[i], "rep")) REPCHK = v; else if (!strcmp(argv[i], "long")) LONGTAB = v; else if (!strcmp(argv[i], "longlog")) LONGLOG = v; else if (!strcmp(argv[i], "look")) LOOK = v; else if (!strcmp(argv[i], "maxoff")) MAXOFF = v;
    }
    const size_t fsz = 4u << 20; uint8_t* buf = malloc(fsz + 64); double total = 0; uint64_t ns = 0, nl = 0, nr = 0; int files = 4;
    for (int f = 0; f < files; f++) {
        fzc_generate(7000000 + f, buf, fsz);
        for (size_t o = 0; o < fsz; o += 131072) total += parse_chunk(buf + o, 131072, &ns, &nl, &nr);
    }
    printf("hashlog %d hbytes %d minmatch %d lazy %d rep %d long %d/%d: ratio %.3f  seqs/MiB %.0f  lit/MiB %.0f  rep %.1f%%  avg ml %.2f\n", HASHLOG, HBYTES, MINMATCH, LAZY, REPCHK, LONGTAB, LONGLOG,
           files * (double)fsz / total, ns / (files * 4.0), nl / (files * 4.0), 100.0 * nr / ns, (files * (double)fsz - nl) / ns);
    return 0;
}
