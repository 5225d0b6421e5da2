"""The reference-facing surface: the zstd-rs names fuse-zstd calls, bound to the GPU codec.

  copy_decode(source, destination)       /root/reference/src/main.rs:463-467
  Encoder(writer, level) + set_pledged_src_size + include_checksum + write + finish
                                         /root/reference/src/main.rs:781-791
  decode_all(source)                     /root/reference/tests/utils.rs:12-17
  read_range(source, offset, size)       the read path, /root/reference/src/main.rs:495-513, without the whole-file
                                         decode at open: only the frames the range touches (needs the seek table
                                         FZG_SEEK_TABLE writes)

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


def read_range(source, offset, size, inode=0):
    """Plain bytes [offset, offset + size) of the .zst file `source`.  Files that carry a seek table are decoded
    partially; any other file is decoded whole (what the reference does at open) and sliced."""
    import errno
    rc, data = codec.decode_range_fd(source.fileno(), offset, size, inode)
    if rc == 0:
        return data
    if rc != -errno.ENOENT:
        raise OSError(errno.EFAULT, "decode failed: %s" % codec.strerror(rc))        # src/main.rs:467
    os.lseek(source.fileno(), 0, os.SEEK_SET)
    return decode_all(source, inode)[offset:offset + size]
