/*
 * corpus_gen.c -- deterministic synthetic JSON-lines corpus.
 *
 * Bench/test support, host only: produces the "synthetic JSON-like corpora of the named
 * sizes" BASELINE.json asks for.  File i is generated from seed = BASE_SEED + i
 * (BASE_SEED = 20261018); records are
 *   {"id":…, "title":…, "authors":[…], "year":…, "doi":…, "abstract":(20-120 words),
 *    "citations":…, "venue":…, "keywords":[…], "score":…}
 * with words drawn Zipf(s = 1.1) from a fixed, seeded 4096-word vocabulary; the file is
 * truncated to the exact requested size.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BASE_SEED 20261018ull
#define VOCAB 4096
#define MAXW 14

static char g_word[VOCAB][MAXW + 1];
static uint8_t g_wlen[VOCAB];
static uint32_t g_prob[VOCAB];   /* alias method: accept threshold (scaled to 2^32) */
static uint16_t g_alias[VOCAB];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t* r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }

static void init_tables(void)
{
    static const char* cons = "bcdfghjklmnprstvwz";
    static const char* vow = "aeiou";
    rng_t r = { BASE_SEED * 7919ull };
    for (int i = 0; i < VOCAB; i++) {
        /* frequent words are short: length grows slowly with rank */
        int len = 2 + (int)rng_below(&r, 3) + (i > 16) + (i > 128) * (int)rng_below(&r, 3) + (i > 1024) * (int)rng_below(&r, 4);
        if (len > MAXW) len = MAXW;
        for (int k = 0; k < len; k++) g_word[i][k] = (k & 1) ? vow[rng_below(&r, 5)] : cons[rng_below(&r, 18)];
        g_word[i][len] = 0; g_wlen[i] = (uint8_t)len;
    }
    /* Zipf weights -> Vose alias tables */
    static double p[VOCAB]; double sum = 0;
    for (int i = 0; i < VOCAB; i++) { p[i] = 1.0 / pow((double)(i + 1), 1.1); sum += p[i]; }
    static int small[VOCAB], large[VOCAB]; int ns = 0, nl = 0;
    for (int i = 0; i < VOCAB; i++) { p[i] = p[i] / sum * VOCAB; if (p[i] < 1.0) small[ns++] = i; else large[nl++] = i; }
    while (ns && nl) {
        int s = small[--ns], l = large[--nl];
        g_prob[s] = (uint32_t)(p[s] * 4294967295.0); g_alias[s] = (uint16_t)l;
        p[l] = p[l] + p[s] - 1.0;
        if (p[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    while (nl) { int l = large[--nl]; g_prob[l] = 0xFFFFFFFFu; g_alias[l] = (uint16_t)l; }
    while (ns) { int s = small[--ns]; g_prob[s] = 0xFFFFFFFFu; g_alias[s] = (uint16_t)s; }
}

static inline int zipf_word(rng_t* r)
{
    uint64_t x = rng_next(r);
    int i = (int)(x & (VOCAB - 1));
    return ((uint32_t)(x >> 32) <= g_prob[i]) ? i : g_alias[i];
}

typedef struct { char* p; char* end; } wr_t;   /* writer with slack; caller guarantees room */
static inline void put(wr_t* w, const char* s, size_t n) { memcpy(w->p, s, n); w->p += n; }
#define PUTS(w, lit) put(w, lit, sizeof(lit) - 1)
static inline void put_word(wr_t* w, int i) { put(w, g_word[i], g_wlen[i]); }
static inline void put_uint(wr_t* w, uint64_t v)
{
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *w->p++ = tmp[--n];
}
static void put_words(wr_t* w, rng_t* r, int n, int capitalise)
{
    for (int k = 0; k < n; k++) {
        if (k) *w->p++ = ' ';
        char* at = w->p;
        put_word(w, zipf_word(r));
        if (capitalise && k == 0) *at = (char)(*at - 32);
    }
}

/* one record is < 2 KiB; generate into a scratch line, then copy what fits */
static size_t gen_record(rng_t* r, uint64_t id, char* line)
{
    wr_t w = { line, line + 4096 };
    PUTS(&w, "{\"id\": "); put_uint(&w, id);
    PUTS(&w, ", \"title\": \""); put_words(&w, r, 4 + (int)rng_below(r, 9), 1);
    PUTS(&w, "\", \"authors\": [");
    int na = 1 + (int)rng_below(r, 5);
    for (int a = 0; a < na; a++) {
        if (a) PUTS(&w, ", ");
        *w.p++ = '"'; *w.p++ = (char)('A' + rng_below(r, 26)); PUTS(&w, ". ");
        char* at = w.p; put_word(&w, 64 + (int)rng_below(r, VOCAB - 64)); *at = (char)(*at - 32);
        *w.p++ = '"';
    }
    PUTS(&w, "], \"year\": "); put_uint(&w, 1950 + rng_below(r, 77));
    PUTS(&w, ", \"doi\": \"10."); put_uint(&w, 1000 + rng_below(r, 9000)); *w.p++ = '/';
    put_word(&w, (int)rng_below(r, VOCAB)); *w.p++ = '.'; put_uint(&w, rng_below(r, 1000000));
    PUTS(&w, "\", \"abstract\": \""); put_words(&w, r, 20 + (int)rng_below(r, 101), 1);
    PUTS(&w, ".\", \"citations\": "); put_uint(&w, rng_below(r, 1 + rng_below(r, 2000)));
    PUTS(&w, ", \"venue\": \""); put_words(&w, r, 2 + (int)rng_below(r, 4), 1);
    PUTS(&w, "\", \"keywords\": [");
    int nk = 2 + (int)rng_below(r, 6);
    for (int k = 0; k < nk; k++) { if (k) PUTS(&w, ", "); *w.p++ = '"'; put_word(&w, zipf_word(r)); *w.p++ = '"'; }
    PUTS(&w, "], \"score\": 0."); { uint32_t s = rng_below(r, 10000); char d[4]; for (int k = 3; k >= 0; k--) { d[k] = (char)('0' + s % 10); s /= 10; } put(&w, d, 4); }
    PUTS(&w, "}\n");
    return (size_t)(w.p - line);
}

void fzc_generate(uint64_t file_index, uint8_t* dst, size_t size)
{
    pthread_once(&g_once, init_tables);
    rng_t r = { (BASE_SEED + file_index) * 0x2545F4914F6CDD1Dull + 1 };
    char line[4096];
    size_t pos = 0; uint64_t id = file_index * 1000003ull % 900000ull + 100000ull;
    while (pos < size) {
        size_t n = gen_record(&r, id, line);
        id += 1 + rng_below(&r, 7);
        if (n > size - pos) n = size - pos;
        memcpy(dst + pos, line, n); pos += n;
    }
}

typedef struct { uint64_t first; size_t n; uint8_t* dst; size_t size; size_t stride; volatile size_t next; } job_t;
static void* worker(void* arg)
{
    job_t* j = (job_t*)arg;
    for (;;) {
        size_t i = __sync_fetch_and_add(&j->next, 1);
        if (i >= j->n) break;
        fzc_generate(j->first + i, j->dst + i * j->stride, j->size);
    }
    return NULL;
}

/* files first..first+n-1, each `size` bytes, written at dst + k*stride */
void fzc_generate_many(uint64_t first, size_t n, uint8_t* dst, size_t size, size_t stride, int threads)
{
    pthread_once(&g_once, init_tables);
    job_t j = { first, n, dst, size, stride, 0 };
    if (threads < 1) threads = 1;
    pthread_t th[256]; if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}
