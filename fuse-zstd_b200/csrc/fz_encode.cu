/*
 * fz_encode.cu -- sm_100a zstd encoder (replaces the Encoder flow of
 * /root/reference/src/main.rs:781-791).  Placeholder until the encoder stage lands: the entry
 * points exist so the ABI is complete, and report -ENOSYS.
 */
#include "fz_host.h"

int fzh_encode_setup(void) { return 0; }
int fzh_encode_run(FzCtx*, uint32_t, uint32_t, int, size_t, int) { return -38; /* -ENOSYS */ }
size_t fzh_encode_bound(size_t src_len, size_t chunk)
{
    if (chunk == 0) chunk = 1u << 20;
    size_t frames = src_len / chunk + 1;
    return src_len + src_len / 128 + frames * 32 + 64;
}
const char* fzh_encode_stage_name(int) { return ""; }
