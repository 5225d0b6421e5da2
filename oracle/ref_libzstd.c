/*
 * ref_libzstd.c -- TEST / BASELINE INFRASTRUCTURE ONLY (built into oracle/_ref/libfzref.so).
 *
 * The reference's codec arithmetic is libzstd (zstd-sys 2.0.13+zstd.1.5.6,
 * /root/reference/Cargo.lock:2389-2390), which is not vendored under /root/reference and
 * cannot be rebuilt here (no cargo/rustc).  The same C library family is installed in this
 * image as /usr/lib/x86_64-linux-gnu/libzstd.so.1 (1.5.5; no headers), so this harness
 * dlopen()s it at run time and restates, call for call, the two zstd-rs call sites:
 *
 *   fzr_copy_decode   = zstd::stream::copy_decode(src, dst)        /root/reference/src/main.rs:463-467
 *       zstd-rs: io::copy(Decoder(BufReader(cap = ZSTD_DStreamInSize() = 131075)), dst) with
 *       io::copy's 8 KiB buffer => ZSTD_decompressStream(out 8 KiB, in <=131075) in a loop,
 *       multi-frame until EOF, EOF inside a frame => error.
 *   fzr_writer_encode = zstd::stream::Encoder::new(w, level) + set_pledged_src_size +
 *       include_checksum(true) + io::copy + finish()               /root/reference/src/main.rs:781-791
 *       => ZSTD_compressStream2(e_continue) per 8 KiB read, out buffer ZSTD_CStreamOutSize(),
 *       then ZSTD_compressStream2(e_end) until 0.
 *   fzr_bulk_compress = zstd::bulk::compress(data, level)          /root/reference/tests/convert.rs:18
 *
 * If libzstd.so.1 is absent, fzr_available() returns 0 and callers fall back to the
 * plain-C restatement in zstd_oracle.c.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { const void* src; size_t size; size_t pos; } in_buf_t;
typedef struct { void* dst; size_t size; size_t pos; } out_buf_t;

static void* g_lib;
static void* (*p_createDCtx)(void);
static size_t (*p_freeDCtx)(void*);
static size_t (*p_decompressStream)(void*, out_buf_t*, in_buf_t*);
static size_t (*p_decompressDCtx)(void*, void*, size_t, const void*, size_t);
static void* (*p_createCCtx)(void);
static size_t (*p_freeCCtx)(void*);
static size_t (*p_CCtx_setParameter)(void*, int, int);
static size_t (*p_CCtx_setPledgedSrcSize)(void*, unsigned long long);
static size_t (*p_CCtx_reset)(void*, int);
static size_t (*p_compressStream2)(void*, out_buf_t*, in_buf_t*, int);
static size_t (*p_compress)(void*, size_t, const void*, size_t, int);
static size_t (*p_compressBound)(size_t);
static unsigned (*p_isError)(size_t);
static unsigned (*p_versionNumber)(void);
static const char* (*p_getErrorName)(size_t);
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void load_lib(void)
{
    const char* names[] = { "libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1", "libzstd.so", NULL };
    for (int i = 0; names[i] && !g_lib; i++) g_lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!g_lib) return;
#define SYM(v, n) *(void**)(&v) = dlsym(g_lib, n); if (!v) { g_lib = NULL; return; }
    SYM(p_createDCtx, "ZSTD_createDCtx") SYM(p_freeDCtx, "ZSTD_freeDCtx")
    SYM(p_decompressStream, "ZSTD_decompressStream") SYM(p_decompressDCtx, "ZSTD_decompressDCtx")
    SYM(p_createCCtx, "ZSTD_createCCtx") SYM(p_freeCCtx, "ZSTD_freeCCtx")
    SYM(p_CCtx_setParameter, "ZSTD_CCtx_setParameter") SYM(p_CCtx_setPledgedSrcSize, "ZSTD_CCtx_setPledgedSrcSize")
    SYM(p_CCtx_reset, "ZSTD_CCtx_reset") SYM(p_compressStream2, "ZSTD_compressStream2")
    SYM(p_compress, "ZSTD_compress") SYM(p_compressBound, "ZSTD_compressBound")
    SYM(p_isError, "ZSTD_isError") SYM(p_versionNumber, "ZSTD_versionNumber") SYM(p_getErrorName, "ZSTD_getErrorName")
#undef SYM
}

int fzr_available(void) { pthread_once(&g_once, load_lib); return g_lib != NULL; }
unsigned fzr_version(void) { return fzr_available() ? p_versionNumber() : 0; }
size_t fzr_compress_bound(size_t n) { return fzr_available() ? p_compressBound(n) : 0; }

#define IN_CAP 131075u   /* ZSTD_DStreamInSize(): BufReader capacity in zstd-rs */
#define COPY_CHUNK 8192u /* std::io::copy stack buffer */
#define OUT_CAP 131591u  /* ZSTD_CStreamOutSize(): zstd-rs Writer buffer */

/* returns 0 ok, 1 libzstd error, 2 EOF inside a frame, 5 dst too small, -1 library missing */
int fzr_copy_decode(const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* out_len)
{
    if (!fzr_available()) return -1;
    void* d = p_createDCtx();
    uint8_t* inbuf = (uint8_t*)malloc(IN_CAP);
    uint8_t chunk[COPY_CHUNK];
    size_t ip = 0, op = 0, hint = 0;
    int rc = 0, eof = 0;
    in_buf_t in = { inbuf, 0, 0 };
    for (;;) {
        if (in.pos == in.size && !eof) {                 /* BufReader::fill_buf */
            size_t n = src_len - ip < IN_CAP ? src_len - ip : IN_CAP;
            memcpy(inbuf, (const uint8_t*)src + ip, n); ip += n;
            in.size = n; in.pos = 0;
            if (n == 0) eof = 1;
        }
        if (eof && in.pos == in.size && hint == 0) break;   /* EOF on a frame boundary (or empty input) */
        out_buf_t out = { chunk, COPY_CHUNK, 0 };
        hint = p_decompressStream(d, &out, &in);
        if (p_isError(hint)) { rc = 1; break; }
        if (out.pos) {
            if (dst_cap - op < out.pos) { rc = 5; break; }
            memcpy((uint8_t*)dst + op, chunk, out.pos); op += out.pos;   /* write(2) into the tmpfile */
        } else if (eof && in.pos == in.size && hint != 0) { rc = 2; break; }  /* UnexpectedEof inside a frame */
    }
    free(inbuf); p_freeDCtx(d);
    if (out_len) *out_len = op;
    return rc;
}

/* the same loop on file descriptors, as fuse-zstd runs it: read(2) of up to 131075 bytes into the BufReader, write(2) of every
 * <= 8 KiB chunk into the tmpfile (src/main.rs:463-467).  Used by the CPU arm of the mount benchmark (oracle/fzfs_codec_ref.c). */
int fzr_copy_decode_fd(int src_fd, int dst_fd, uint64_t* out_len)
{
    if (!fzr_available()) return -1;
    void* d = p_createDCtx();
    uint8_t* inbuf = (uint8_t*)malloc(IN_CAP);
    uint8_t chunk[COPY_CHUNK];
    size_t hint = 0; uint64_t op = 0;
    int rc = 0, eof = 0;
    in_buf_t in = { inbuf, 0, 0 };
    for (;;) {
        if (in.pos == in.size && !eof) {
            ssize_t n;
            do n = read(src_fd, inbuf, IN_CAP); while (n < 0 && errno == EINTR);
            if (n < 0) { rc = 1; break; }
            in.size = (size_t)n; in.pos = 0;
            if (n == 0) eof = 1;
        }
        if (eof && in.pos == in.size && hint == 0) break;
        out_buf_t out = { chunk, COPY_CHUNK, 0 };
        hint = p_decompressStream(d, &out, &in);
        if (p_isError(hint)) { rc = 1; break; }
        if (out.pos) {
            size_t w = 0;
            while (w < out.pos) { ssize_t k = write(dst_fd, chunk + w, out.pos - w); if (k < 0) { if (errno == EINTR) continue; rc = 1; break; } w += (size_t)k; }
            if (rc) break;
            op += out.pos;
        } else if (eof && in.pos == in.size && hint != 0) { rc = 2; break; }
    }
    free(inbuf); p_freeDCtx(d);
    if (out_len) *out_len = op;
    return rc;
}

/* one-shot upper bound: ZSTD_decompressDCtx over concatenated frames */
int fzr_decode_oneshot(const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* out_len)
{
    if (!fzr_available()) return -1;
    void* d = p_createDCtx();
    size_t r = p_decompressDCtx(d, dst, dst_cap, src, src_len);
    p_freeDCtx(d);
    if (p_isError(r)) return 1;
    if (out_len) *out_len = r;
    return 0;
}

/* the reference writer (src/main.rs:781-791).  checksum/pledge are parameters so tests can
 * also produce "stock zstd CLI"-style frames; the reference uses pledge=1, checksum=1. */
int fzr_writer_encode(const void* src, size_t src_len, void* dst, size_t dst_cap, int level,
                      int pledge, int checksum, int window_log, size_t* out_len)
{
    if (!fzr_available()) return -1;
    void* c = p_createCCtx();
    int rc = 0;
    if (p_isError(p_CCtx_setParameter(c, 100 /*ZSTD_c_compressionLevel*/, level))) rc = 1;
    if (window_log && p_isError(p_CCtx_setParameter(c, 101 /*ZSTD_c_windowLog*/, window_log))) rc = 1;
    if (p_isError(p_CCtx_setParameter(c, 201 /*ZSTD_c_checksumFlag*/, checksum))) rc = 1;
    if (pledge && p_isError(p_CCtx_setPledgedSrcSize(c, src_len))) rc = 1;
    uint8_t* outbuf = (uint8_t*)malloc(OUT_CAP);
    size_t op = 0, ip = 0;
    while (!rc && ip < src_len) {                       /* io::copy: 8 KiB reads, write_all into the encoder */
        size_t n = src_len - ip < COPY_CHUNK ? src_len - ip : COPY_CHUNK;
        in_buf_t in = { (const uint8_t*)src + ip, n, 0 };
        while (in.pos < in.size) {
            out_buf_t out = { outbuf, OUT_CAP, 0 };
            size_t r = p_compressStream2(c, &out, &in, 0 /*ZSTD_e_continue*/);
            if (p_isError(r)) { rc = 1; break; }
            if (out.pos) { if (dst_cap - op < out.pos) { rc = 5; break; } memcpy((uint8_t*)dst + op, outbuf, out.pos); op += out.pos; }
        }
        ip += n;
    }
    while (!rc) {                                       /* finish() */
        in_buf_t in = { NULL, 0, 0 };
        out_buf_t out = { outbuf, OUT_CAP, 0 };
        size_t r = p_compressStream2(c, &out, &in, 2 /*ZSTD_e_end*/);
        if (p_isError(r)) { rc = 1; break; }
        if (out.pos) { if (dst_cap - op < out.pos) { rc = 5; break; } memcpy((uint8_t*)dst + op, outbuf, out.pos); op += out.pos; }
        if (r == 0) break;
    }
    free(outbuf); p_freeCCtx(c);
    if (out_len) *out_len = op;
    return rc;
}

int fzr_bulk_compress(const void* src, size_t src_len, void* dst, size_t dst_cap, int level, size_t* out_len)
{
    if (!fzr_available()) return -1;
    size_t r = p_compress(dst, dst_cap, src, src_len, level);
    if (p_isError(r)) return 1;
    if (out_len) *out_len = r;
    return 0;
}

/* ---------------------------------------------------------------- threaded batches */
typedef struct {
    int mode;  /* 0 copy_decode, 1 oneshot decode, 2 writer encode */
    size_t n; const void* const* src; const size_t* src_len; void* const* dst; const size_t* dst_cap;
    size_t* out_len; int* status; int level;
    volatile size_t next; int nthreads;
} batch_t;

static void* batch_worker(void* arg)
{
    batch_t* b = (batch_t*)arg;
    for (;;) {
        size_t i = __sync_fetch_and_add(&b->next, 1);
        if (i >= b->n) break;
        size_t o = 0; int rc;
        if (b->mode == 0) rc = fzr_copy_decode(b->src[i], b->src_len[i], b->dst[i], b->dst_cap[i], &o);
        else if (b->mode == 1) rc = fzr_decode_oneshot(b->src[i], b->src_len[i], b->dst[i], b->dst_cap[i], &o);
        else rc = fzr_writer_encode(b->src[i], b->src_len[i], b->dst[i], b->dst_cap[i], b->level & 0xFF, 1, 1, (b->level >> 8) & 0xFF, &o);   /* level | windowLog << 8 */
        b->out_len[i] = o; b->status[i] = rc;
    }
    return NULL;
}

/* runs the batch on `threads` host threads (one file per thread at a time); returns wall seconds */
double fzr_batch(int mode, int threads, size_t n, const void* const* src, const size_t* src_len,
                 void* const* dst, const size_t* dst_cap, size_t* out_len, int* status, int level)
{
    if (!fzr_available()) return -1.0;
    batch_t b = { mode, n, src, src_len, dst, dst_cap, out_len, status, level, 0, threads };
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, batch_worker, &b);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
