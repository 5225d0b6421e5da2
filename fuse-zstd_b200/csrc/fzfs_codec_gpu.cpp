/* fzfs_codec_gpu.cpp -- the product codec of the fzfs host: libfzgpu.so (include/fzgpu.h).  No CPU path. */
#include "fzfs_codec.h"
#include <stdio.h>
#include <stdlib.h>
#include "../../include/fzgpu.h"

static bool g_cache = false;

extern "C" int fzfs_codec_init(size_t cache_bytes)
{
    const int rc = fzg_init(nullptr, 0);
    if (rc) return rc;
    g_cache = cache_bytes != 0;
    if (int e = fzg_cache_configure(cache_bytes)) return e;
    const int slabs = g_cache ? fzg_cache_reserve() : 0;             // pinned memory is slow to allocate: before the mount appears
    return slabs < 0 ? slabs : 0;
}
extern "C" const char* fzfs_codec_name(void) { return "fzgpu (CUDA, sm_100a)"; }
// readers of cached files are served from pinned memory by whichever thread is free: the host's request loop is the bottleneck
// of the read path through the mount (profiles/r02_notes.md, section 7), so it gets several threads
extern "C" int fzfs_codec_threads(void) { return 8; }
extern "C" int fzfs_decode(int src_fd, int dst_fd, uint64_t ino, uint64_t* out_size)
{
    return g_cache ? fzg_cache_open(src_fd, dst_fd, ino, out_size, nullptr) : fzg_decode_fd(src_fd, dst_fd, ino, out_size);
}
extern "C" int fzfs_encode(int src_fd, int dst_fd, int level, uint64_t src_size, uint64_t ino, uint64_t* out_size)
{
    return fzg_encode_fd(src_fd, dst_fd, level, src_size, ino, out_size);
}
extern "C" int fzfs_prefetch(const char* const* paths, const uint64_t* inos, size_t n)
{
    if (!g_cache || n == 0) return 0;
    const int devs = fzg_device_count();                             // a directory's batch goes to one GPU: sharded by inode like the opens
    return fzg_cache_prefetch_async(devs > 0 ? (int)(inos[0] % (uint64_t)devs) : 0, paths, inos, n);
}
extern "C" void fzfs_invalidate(uint64_t ino) { if (g_cache) fzg_cache_invalidate(ino); }
extern "C" void fzfs_codec_shutdown(void)
{
    if (g_cache && getenv("FZFS_CACHE_STATS")) {
        uint64_t h = 0, m = 0, b = 0, f = 0; fzg_cache_stats(&h, &m, &b, &f);
        fprintf(stderr, "fzfs: cache: %llu opens served from it, %llu not; %llu files / %.1f MB held at exit\n", (unsigned long long)h, (unsigned long long)m, (unsigned long long)f, b / 1e6);
    }
    fzg_shutdown();                                                  // drains the prefetch threads first
}
extern "C" int fzfs_view(int src_fd, uint64_t ino, const void** data, uint64_t* size, void** pin)
{
    return g_cache ? fzg_cache_view(src_fd, ino, data, size, pin) : -1;
}
extern "C" void fzfs_unview(void* pin) { fzg_cache_unview(pin); }
extern "C" void fzfs_wait(uint64_t ino) { if (g_cache) fzg_cache_wait(ino); }
extern "C" int fzfs_pending(uint64_t ino) { return g_cache ? fzg_cache_pending(ino) : 0; }
