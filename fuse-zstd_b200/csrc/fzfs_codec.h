/*
 * fzfs_codec.h -- the codec boundary of the fzfs host: exactly the two call sites fuse-zstd has
 * (/root/reference/src/main.rs:463-467 and :781-791) plus the hooks batch formation needs.
 * Implementations: fzfs_codec_gpu.cpp (product: libfzgpu.so) and oracle/fzfs_codec_ref.c (measurement baseline: the
 * reference's libzstd calls).  All return 0 or a non-zero error; the host maps decode failures to EFAULT and encode
 * failures to EIO as the reference does.
 */
#ifndef FZFS_CODEC_H
#define FZFS_CODEC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int fzfs_codec_init(size_t cache_bytes);            /* 0 = no decoded-file cache / readahead */
const char* fzfs_codec_name(void);
int fzfs_codec_threads(void);                       /* serving threads of the host when --threads is not given */
int fzfs_decode(int src_fd, int dst_fd, uint64_t ino, uint64_t* out_size);                       /* copy_decode */
int fzfs_encode(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t ino, uint64_t* out_size);   /* Encoder ... finish */
int fzfs_prefetch(const char* const* paths, const uint64_t* inos, size_t n);                     /* returns at once */
void fzfs_invalidate(uint64_t ino);
int fzfs_view(int src_fd, uint64_t ino, const void** data, uint64_t* size, void** pin);        /* read-only open: the plain bytes in place, or non-zero (use fzfs_decode) */
void fzfs_unview(void* pin);
int fzfs_pending(uint64_t ino);                     /* non-zero while `ino` is in a readahead batch in flight */
void fzfs_wait(uint64_t ino);                       /* returns once `ino` is in no readahead batch in flight (called outside the host's lock) */
void fzfs_codec_shutdown(void);                    /* before the daemon exits: nothing of the codec may still be running */
#ifdef __cplusplus
}
#endif
#endif
