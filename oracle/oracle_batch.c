/*
 * oracle_batch.c -- TEST / BASELINE INFRASTRUCTURE ONLY.
 * Runs the plain-C restatement (zstd_oracle.c) over a batch of files on N host threads,
 * one file per thread at a time; used as bench.py's cpu_baseline "port" leg when the
 * system libzstd is not available on the timing box.
 */
#include "zstd_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

typedef struct {
    size_t n; const void* const* src; const size_t* src_len; void* const* dst; const size_t* dst_cap;
    size_t* out_len; int* status; volatile size_t next;
} ob_t;

static void* ob_worker(void* arg)
{
    ob_t* b = (ob_t*)arg;
    for (;;) {
        size_t i = __sync_fetch_and_add(&b->next, 1);
        if (i >= b->n) break;
        size_t o = 0;
        b->status[i] = fzo_decode(b->src[i], b->src_len[i], b->dst[i], b->dst_cap[i], &o);
        b->out_len[i] = o;
    }
    return NULL;
}

double fzo_decode_batch(int threads, size_t n, const void* const* src, const size_t* src_len,
                        void* const* dst, const size_t* dst_cap, size_t* out_len, int* status)
{
    ob_t b = { n, src, src_len, dst, dst_cap, out_len, status, 0 };
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, ob_worker, &b);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
