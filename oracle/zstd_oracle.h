/*
 * zstd_oracle.h -- TEST INFRASTRUCTURE ONLY (see zstd_oracle.c header).
 *
 * CPU restatement of the Zstandard decode arithmetic that fuse-zstd reaches at
 *   /root/reference/src/main.rs:463-467   zstd::stream::copy_decode(src, dst)
 * through zstd 0.13.2 -> zstd-safe 7.2.1 -> zstd-sys 2.0.13+zstd.1.5.6
 * (/root/reference/Cargo.lock:2371-2390; sources NOT under /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link
 * or call this.  The product library (libfzgpu.so) never does.
 */
#ifndef FZ_ORACLE_H
#define FZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: identical numbering to include/fzgpu.h FZG_E_* */
enum {
    FZO_OK = 0,
    FZO_E_MAGIC = 1,       /* unknown frame descriptor / bad magic / trailing garbage */
    FZO_E_TRUNCATED = 2,   /* input ends inside a frame */
    FZO_E_UNSUPPORTED = 3, /* reserved bit, dictionary id, window > 2^27 */
    FZO_E_CORRUPT = 4,     /* any entropy / sequence / block inconsistency */
    FZO_E_DSTSIZE = 5,     /* destination capacity too small */
    FZO_E_CHECKSUM = 6,    /* XXH64 trailer mismatch */
    FZO_E_FCS = 7          /* produced size != Frame_Content_Size */
};

/* Optional per-block stage recorder (used by tests to compare GPU intermediates). */
typedef struct {
    uint32_t lit_len;    /* literal run length */
    uint32_t match_len;  /* match length (>=3) */
    uint32_t offset;     /* resolved match distance (after repeat-offset handling) */
    uint32_t of_value;   /* raw Offset_Value as decoded from the bitstream */
} fzo_seq_t;

typedef struct {
    /* caller-provided capacity, filled by the decoder; any pointer may be NULL */
    uint8_t*   literals;      size_t literals_cap;  size_t literals_len;
    fzo_seq_t* seqs;          size_t seqs_cap;      size_t seqs_len;
    uint32_t*  block_nseq;    size_t blocks_cap;    size_t blocks_len;   /* per block */
    uint32_t*  block_litsize;                                        /* per block */
    uint8_t*   block_type;                                           /* 0 raw 1 rle 2 compressed */
    uint8_t*   block_littype;  /* literals type | (streams==4)<<2 | (FSE-compressed weights)<<3 */
    uint8_t*   block_modes;    /* Symbol_Compression_Modes byte (0 when nseq == 0) */
} fzo_trace_t;

/* Decode a whole buffer of concatenated frames (zstd + skippable) as
 * zstd::stream::copy_decode does.  Returns FZO_*; *out_len = bytes produced. */
int fzo_decode(const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* out_len);
int fzo_decode_trace(const void* src, size_t src_len, void* dst, size_t dst_cap, size_t* out_len,
                     fzo_trace_t* trace);

/* Sum of Frame_Content_Size over all frames; returns FZO_OK, or an error if a
 * header is malformed / FCS is absent (*content_size = UINT64_MAX then). */
int fzo_frame_info(const void* src, size_t src_len, uint64_t* content_size, uint64_t* n_frames);

uint64_t fzo_xxh64(const void* data, size_t len, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
