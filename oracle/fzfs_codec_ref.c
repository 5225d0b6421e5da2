/*
 * fzfs_codec_ref.c -- TEST / MEASUREMENT INFRASTRUCTURE, not part of the product: the codec boundary of the fzfs host
 * (fuse-zstd_b200/csrc/fzfs_codec.h) bound to the reference's own codec calls, i.e. zstd-rs's copy_decode and Encoder
 * restated on the system libzstd (oracle/ref_libzstd.c; /root/reference/src/main.rs:463-467, 781-791).  Linked into
 * oracle/_ref/fzfs_ref only, the CPU arm of tools/mount_bench.py.  No readahead, no cache: one decode per open, on the
 * host's one FUSE thread, as fuse-zstd does (DESIGN.md:5-7 of the reference).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../fuse-zstd_b200/csrc/fzfs_codec.h"

int fzr_available(void);
size_t fzr_compress_bound(size_t n);
int fzr_copy_decode_fd(int src_fd, int dst_fd, uint64_t* out_len);
int fzr_writer_encode(const void* src, size_t src_len, void* dst, size_t dst_cap, int level, int pledge, int checksum, int window_log, size_t* out_len);

int fzfs_codec_init(size_t cache_bytes) { (void)cache_bytes; return fzr_available() ? 0 : -ENOSYS; }
const char* fzfs_codec_name(void) { return "libzstd (reference restatement, one thread)"; }
int fzfs_codec_threads(void) { return 1; }          /* fuser's session loop: one request at a time */

int fzfs_decode(int src_fd, int dst_fd, uint64_t ino, uint64_t* out_size) { (void)ino; return fzr_copy_decode_fd(src_fd, dst_fd, out_size); }

static int read_full(int fd, uint8_t* p, size_t n) { size_t g = 0; while (g < n) { ssize_t r = read(fd, p + g, n - g); if (r < 0) { if (errno == EINTR) continue; return -1; } if (r == 0) break; g += (size_t)r; } return g == n ? 0 : -1; }
static int write_full(int fd, const uint8_t* p, size_t n) { while (n) { ssize_t w = write(fd, p, n); if (w < 0) { if (errno == EINTR) continue; return -1; } p += w; n -= (size_t)w; } return 0; }

int fzfs_encode(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t ino, uint64_t* out_size)
{
    (void)ino;
    uint8_t* in = (uint8_t*)malloc(src_size ? src_size : 1);
    const size_t cap = fzr_compress_bound(src_size) + 64;
    uint8_t* out = (uint8_t*)malloc(cap);
    size_t n = 0; int rc = -1;
    if (in && out && read_full(src_fd, in, src_size) == 0 && fzr_writer_encode(in, src_size, out, cap, level ? level : 3, 1, 1, 0, &n) == 0 && write_full(dst_fd, out, n) == 0) rc = 0;
    free(in); free(out);
    if (out_size) *out_size = n;
    return rc;
}
int fzfs_prefetch(const char* const* paths, const uint64_t* inos, size_t n) { (void)paths; (void)inos; (void)n; return 0; }
void fzfs_invalidate(uint64_t ino) { (void)ino; }
void fzfs_codec_shutdown(void) { }
int fzfs_view(int src_fd, uint64_t ino, const void** data, uint64_t* size, void** pin) { (void)src_fd; (void)ino; (void)data; (void)size; (void)pin; return -1; }
void fzfs_unview(void* pin) { (void)pin; }
void fzfs_wait(uint64_t ino) { (void)ino; }
int fzfs_pending(uint64_t ino) { (void)ino; return 0; }
