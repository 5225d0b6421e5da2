/*
 * fzgpu.h -- C ABI of libfzgpu.so, the B200-native zstd codec behind fuse-zstd's codec boundary.
 *
 * Plain pointers and sizes only; no torch / C++ types.  Every entry point names the reference
 * interface it replaces (paths relative to /root/reference).  INTEGRATION.md shows the ~30-line
 * Rust `extern "C"` shim a fuse-zstd maintainer would add.
 *
 * Error convention (mirrors src/errors.rs:4-10 and src/main.rs:467): functions return 0 on
 * success or a NEGATIVE errno-style code for call-level failures (-EINVAL, -ENOMEM, -ENODEV,
 * -EIO); per-item codec results are reported as non-negative FZG_E_* status codes.  The Rust
 * shim maps any non-zero decode result to libc::EFAULT (src/main.rs:467) and any encode failure
 * to EIO (src/errors.rs:9).
 *
 * There is no CPU fallback anywhere in this library: without a CUDA device every compute entry
 * point fails with -ENODEV.
 */
#ifndef FZGPU_H
#define FZGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-item codec status (same numbering as oracle/zstd_oracle.h FZO_*) */
enum {
    FZG_OK = 0,
    FZG_E_MAGIC = 1,       /* unknown frame descriptor / trailing garbage          */
    FZG_E_TRUNCATED = 2,   /* input ends inside a frame (zstd-rs: UnexpectedEof)    */
    FZG_E_UNSUPPORTED = 3, /* reserved FHD bit, dictionary id, window > 2^27        */
    FZG_E_CORRUPT = 4,     /* entropy / sequence / block inconsistency              */
    FZG_E_DSTSIZE = 5,     /* destination capacity too small                        */
    FZG_E_CHECKSUM = 6,    /* XXH64 content checksum mismatch                       */
    FZG_E_FCS = 7          /* produced size != Frame_Content_Size                   */
};

/* flags for the *_batch entry points */
enum {
    FZG_SRC_DEVICE = 1,         /* src[i] are device pointers (batch already resident in HBM)   */
    FZG_DST_DEVICE = 2,         /* dst[i] are device pointers                                   */
    FZG_NO_VERIFY_CHECKSUM = 4, /* skip XXH64 verification (default: verify, as libzstd does)   */
    FZG_PROFILE = 8,            /* record CUDA-event time per kernel (see fzg_last_timing)      */
    FZG_SEEK_TABLE = 16         /* encode: append a seek table (zstd seekable format) per file  */
};

/* Creates one context (stream, pinned staging, scratch) per listed CUDA device.  devices == NULL
 * means "device 0 .. n_devices-1"; n_devices == 0 means all visible devices.  Idempotent. */
int fzg_init(const int* devices, int n_devices);
void fzg_shutdown(void);
int fzg_device_count(void); /* devices with a context */

/*
 * Replaces  zstd::stream::copy_decode(source_file, target_file)      src/main.rs:463-467
 * Reads src_fd from its current offset to EOF (all concatenated / skippable frames), writes the
 * plain bytes to dst_fd at its current offset and leaves both offsets at the end, exactly as
 * copy_decode's io::copy does.  shard_key = the fuse-zstd inode (src/main.rs:744-753); the GPU is
 * shard_key % n_devices.  Returns 0, a positive FZG_E_* status, or -errno for I/O failures.
 */
int fzg_decode_fd(int src_fd, int dst_fd, uint64_t shard_key, uint64_t* out_size);

/*
 * Replaces  Encoder::new(w, level) + set_pledged_src_size(Some(n)) + include_checksum(true) +
 *           io::copy + finish()                                       src/main.rs:781-791
 * Reads src_size bytes from src_fd (current offset), writes zstd frames to dst_fd.  level follows
 * src/main.rs:1233-1241 (0 => default 3): 1 and 2 select the faster single-table matcher, every other
 * value the two-table one.  Every frame carries Frame_Content_Size and the XXH64
 * content checksum; the output is a concatenation of independent frames that stock libzstd
 * (>= 1.0) decodes to the input.
 */
int fzg_encode_fd(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t shard_key,
                  uint64_t* out_size);

/*
 * Batched decode -- the entry point configs 2-4 time.  n independent .zst files; src/dst are
 * host arrays of n pointers (host or device memory, see flags).  dst_len[i] receives the plain
 * size, status[i] an FZG_E_* code.  One launch sequence for all n items.  Blocking.
 */
int fzg_decode_batch(int device, size_t n, const void* const* src, const size_t* src_len,
                     void* const* dst, const size_t* dst_cap, size_t* dst_len, int* status, int flags);

/* Batched encode (config 5).  chunk_size = bytes per independent frame (0 => default). */
int fzg_encode_batch(int device, size_t n, const void* const* src, const size_t* src_len,
                     void* const* dst, const size_t* dst_cap, size_t* dst_len, int* status,
                     int level, size_t chunk_size, int flags);
size_t fzg_encode_bound(size_t src_len, size_t chunk_size);

/*
 * Host-side header walk (no GPU): sum of Frame_Content_Size over all frames and the number of
 * compressed bytes the frames occupy.  Lets `lookup`/`getattr` (src/main.rs:40-47) size a file
 * without decoding it.  *content_size = UINT64_MAX when some frame omits the field.
 */
int fzg_frame_info(const void* src, size_t len, uint64_t* content_size, uint64_t* compressed_size);

/*
 * Partial reads (the reference decodes the whole file on open and serves read(offset, size) from the
 * tmpfile, src/main.rs:495-513).  A file written with FZG_SEEK_TABLE ends in a skippable frame of the zstd seekable
 * format; these calls read that table and decode ONLY the frames that hold plain bytes [offset, offset + size).
 * *got = bytes produced (short at end of file).  Returns -ENOENT when the file has no valid seek table (the caller
 * falls back to fzg_decode_fd), a positive FZG_E_* status for a corrupt frame, -errno for I/O failures.
 */
int fzg_decode_range(int device, const void* src, size_t len, uint64_t offset, size_t size, void* dst, size_t* got);
int fzg_decode_range_fd(int src_fd, uint64_t shard_key, uint64_t offset, size_t size, void* dst, size_t* got);
/* host-only: parses the seek-table footer from the last bytes of a file (no GPU) */
int fzg_seek_footer(const void* tail, size_t tail_len, uint64_t file_size, uint32_t* n_frames, uint64_t* table_bytes);

/*
 * Batch formation + decoded-file cache.  fuse-zstd decodes one file per open() on one thread
 * (DESIGN.md:5-7 of the reference) and forgets the result with the last handle (src/file.rs:104-117); a GPU decodes ten
 * thousand files in the time of one.  fzg_cache_prefetch decodes the listed .zst files (e.g. the siblings readdir_wrapper,
 * src/main.rs:307-387, has just listed) as ONE batch and keeps the plain bytes, keyed by inode; fzg_cache_open is
 * open_wrapper's codec call (src/main.rs:463-467) with that cache in front of fzg_decode_fd (an entry is served only
 * while the source file still has the size and mtime it was decoded from).  Pinned slabs, reused oldest-first above the configured capacity.
 */
int fzg_cache_configure(size_t capacity_bytes);   /* default 1 GiB; 0 disables prefetching */
int fzg_cache_reserve(void);                      /* allocate the pinned slabs now (slow: once, at mount time) */
int fzg_cache_prefetch(int device, const char* const* paths, const uint64_t* keys, size_t n);        /* -> files added */
int fzg_cache_prefetch_async(int device, const char* const* paths, const uint64_t* keys, size_t n);  /* detached thread */
int fzg_cache_open(int src_fd, int dst_fd, uint64_t key, uint64_t* out_size, int* hit);
int fzg_cache_invalidate(uint64_t key);           /* after store_to_source_file / rename / unlink */
void fzg_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* bytes, uint64_t* files);
int fzg_cache_view(int src_fd, uint64_t key, const void** data, uint64_t* size, void** token);   /* the cached plain bytes in place (read-only opens:
                                                     no tmpfile copy of src/main.rs:462-466); 0, or -ENOENT when not cached / stale; pinned until ... */
void fzg_cache_unview(void* token);               /* ... the view is given back */
int fzg_cache_pending(uint64_t key);              /* 1 while `key` belongs to a prefetch batch in flight (set before fzg_cache_prefetch_async returns) */
void fzg_cache_wait(uint64_t key);                /* returns once `key` is in no prefetch batch in flight (several batches decode side by side
                                                     and are taken to the GPU together; an open waits for its file here, outside its own locks) */
void fzg_cache_drain(void);                       /* waits for the prefetch batches in flight (called by fzg_shutdown; a daemon calls it before exit) */

const char* fzg_strerror(int code);

/* ---- measurement hooks (bench.py); not part of the reference surface ---- */
typedef struct {
    float total_ms;        /* first kernel launch -> last kernel end, CUDA events on the ctx stream */
    float kernel_ms[16];   /* per pipeline stage, FZG_PROFILE only (see fzg_stage_name)             */
    int   launches;        /* kernels launched by the last *_batch call                              */
    uint64_t bytes_in, bytes_out;
} fzg_timing_t;
int fzg_last_timing(int device, fzg_timing_t* out);
const char* fzg_stage_name(int stage);
/* Host-resident decode of few large frames: device -> host copies of finished parts of a frame are queued while the frame is
 * still being executed (chunks of >= FZG_STREAM_OUT_MB MiB of output, default 256, at most half as many frames as SMs).
 * Returns how many such copies this context has queued since fzg_init (a counter for tests and traces). */
uint64_t fzg_streamed_copies(int device);
/* Host-resident batches are staged in HBM; the staging grows on demand, and growing it (cudaFree + cudaMalloc) stalls whatever
 * else the device is doing.  A daemon that knows its batch sizes reserves them once: `items` files, src_bytes compressed,
 * dst_bytes plain per call.  fzg_cache_reserve does this for the cache's own batches. */
int fzg_reserve_staging(int device, size_t items, size_t src_bytes, size_t dst_bytes);
void* fzg_stream(int device); /* cudaStream_t of the context (for external event timing) */

#ifdef __cplusplus
}
#endif
#endif
