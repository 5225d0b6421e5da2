"""fuse-zstd_b200 -- B200-native zstd codec behind fuse-zstd's codec boundary.

The product is libfzgpu.so (hand-written CUDA for sm_100a + a C ABI, see include/fzgpu.h).  This
package is the thin host-side mirror used by tests and benchmarks:

  codec   -- ctypes bindings of the C ABI (batched decode / encode, fd entry points)
  stream  -- the reference-facing names: copy_decode, decode_all, Encoder (zstd-rs API as used at
             /root/reference/src/main.rs:463-467 and :781-791)
  corpus  -- deterministic synthetic JSON corpus (bench/test support)
  shard   -- file -> GPU / rank partitioning (by inode, no collective) + the max-over-ranks timing helper

There is no CPU fallback: importing `codec` without the built library raises, and every compute
call without a CUDA device fails with ENODEV.
"""
import importlib as _importlib

__all__ = ["codec", "stream", "corpus", "shard"]


def __getattr__(name):
    if name in __all__:
        return _importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
