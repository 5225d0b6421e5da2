"""The reference-facing surface: the zstd-rs names fuse-zstd calls, bound to the GPU codec.

  copy_decode(source, destination)       /root/reference/src/main.rs:463-467
  Encoder(writer, level) + set_pledged_src_size + include_checksum + write + finish
                                         /root/reference/src/main.rs:781-791
  decode_all(source)                     /root/reference/tests/utils.rs:12-17

Arguments are Python file objects (anything with fileno()), standing in for the dup'd
std::fs::File handles the reference passes.  Error behaviour follows the reference: a decode
failure of any kind is OSError(EFAULT) (src/main.rs:467); encode failures carry the raw errno or
EIO (src/errors.rs:4-10).
"""
import os
import tempfile

from . import codec


def copy_decode(source, destination, inode=0):
    """Decode every frame of `source` (from its current offset to EOF) into `destination`."""
    source.flush() if hasattr(source, "flush") and source.writable() else None
    destination.flush()
    return codec.decode_fd(source.fileno(), destination.fileno(), inode)


def decode_all(source, inode=0):
    with tempfile.TemporaryFile() as tmp:
        copy_decode(source, tmp, inode)
        tmp.seek(0)
        return tmp.read()


class Encoder:
    """zstd::stream::Encoder as fuse-zstd drives it: buffered writes, one encode at finish()."""

    def __init__(self, writer, level=0, inode=0):
        if not 0 <= level <= 19:
            level = 0                       # src/main.rs:1283-1296: out of range => default
        self.writer, self.level, self.inode = writer, level, inode
        self.pledged, self.checksum = None, False
        self._spool = tempfile.TemporaryFile()

    def set_pledged_src_size(self, size):
        self.pledged = size

    def include_checksum(self, flag):
        self.checksum = bool(flag)          # the GPU encoder always writes the checksum (src/main.rs:789)

    def write(self, data):
        self._spool.write(data)
        return len(data)

    def finish(self):
        self._spool.flush()
        size = self._spool.tell()
        if self.pledged is not None and self.pledged != size:
            raise OSError(5, "pledged source size mismatch")   # libzstd: Src size is incorrect -> EIO
        self._spool.seek(0)
        os.lseek(self._spool.fileno(), 0, os.SEEK_SET)
        self.writer.flush()
        n = codec.encode_fd(self._spool.fileno(), self.writer.fileno(), self.level, size, self.inode)
        self._spool.close()
        return n
