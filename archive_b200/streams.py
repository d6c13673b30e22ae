"""InputMemoryStream / OutputMemoryStream -- the in-memory byte streams the reference's codec hot
path is written against (lib/src/util/input_memory_stream.dart:8-134,
lib/src/util/output_memory_stream.dart:8-137).  Only the members the codecs touch are mirrored."""
from __future__ import annotations

import os

LITTLE_ENDIAN, BIG_ENDIAN = 0, 1


class InputMemoryStream:
    def __init__(self, data=b"", byte_order: int = LITTLE_ENDIAN, offset: int | None = None, length: int | None = None):
        data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
        offset = offset or 0
        if length is None or offset + length > len(data):
            length = len(data) - offset  # input_memory_stream.dart:17-22
        self.buffer = memoryview(data)[offset:offset + length]
        self.position = 0
        self.byte_order = byte_order

    @property
    def length(self) -> int:  # bytes LEFT (input_memory_stream.dart:56)
        return len(self.buffer) - self.position

    @property
    def is_eos(self) -> bool:
        return self.position >= len(self.buffer)

    def set_position(self, v: int):
        self.position = v

    def read_byte(self) -> int:
        b = self.buffer[self.position]  # IndexError past the end, as Dart's RangeError
        self.position += 1
        return b

    def read_bytes(self, count: int) -> "InputMemoryStream":
        s = InputMemoryStream(self.buffer, self.byte_order, self.position, count)
        self.position += len(s.buffer)
        return s

    def to_uint8_list(self) -> bytes:
        return bytes(self.buffer[self.position:])


class OutputMemoryStream:
    default_buffer_size = 0x8000

    def __init__(self, size: int | None = None, byte_order: int = LITTLE_ENDIAN):
        self._buf = bytearray()
        self.byte_order = byte_order

    @property
    def length(self) -> int:
        return len(self._buf)

    def clear(self):
        del self._buf[:]

    def flush(self):
        pass

    def write_byte(self, v: int):
        self._buf.append(v & 0xff)

    def write_bytes(self, data, length: int | None = None):
        self._buf += bytes(data if length is None else data[:length])

    def write_uint32(self, v: int):
        self._buf += int(v & 0xffffffff).to_bytes(4, "big" if self.byte_order == BIG_ENDIAN else "little")

    def write_uint16(self, v: int):
        self._buf += int(v & 0xffff).to_bytes(2, "big" if self.byte_order == BIG_ENDIAN else "little")

    def get_bytes(self) -> bytes:
        return bytes(self._buf)


class InputFileStream:
    """InputFileStream (lib/src/util/input_file_stream.dart:11-221): a window [file_offset, file_offset + file_size) of a
    file with a read position.  The reference reads through a FileBuffer cache (file_buffer.dart:10, `buffer_size` bytes);
    here small reads are os.pread calls and the codecs never pull bytes through this class at all: they hand the path and
    the byte range to the library (b200z_file_codec, include/b200z.h), which stages the file in page-locked segments."""

    def __init__(self, path: str, byte_order: int = LITTLE_ENDIAN, buffer_size: int = 1024, _share=None,
                 _file_offset: int = 0, _file_size: int | None = None):
        self.path = path
        self.byte_order = byte_order
        self.buffer_size = buffer_size
        self._fd_box = _share if _share is not None else [os.open(path, os.O_RDONLY)]  # shared with sub-streams (:71-80)
        self._file_offset = _file_offset
        self._file_size = os.fstat(self._fd_box[0]).st_size if _file_size is None else _file_size
        self._position = 0

    @classmethod
    def from_file_stream(cls, other: "InputFileStream", position: int | None = None, length: int | None = None,
                         buffer_size: int | None = None) -> "InputFileStream":
        """InputFileStream.fromFileStream (:71-80).  As in the reference, `length` is taken as given (not clipped)."""
        return cls(other.path, other.byte_order, buffer_size or other.buffer_size, _share=other._fd_box,
                   _file_offset=other._file_offset + (position or 0),
                   _file_size=other._file_size if length is None else length)

    def open(self) -> bool:
        if self._fd_box[0] is None:
            self._fd_box[0] = os.open(self.path, os.O_RDONLY)
        return True

    def close_sync(self):
        if self._fd_box[0] is not None:
            os.close(self._fd_box[0])
            self._fd_box[0] = None
        self._position = 0
        self._file_size = 0

    close = close_sync

    @property
    def length(self) -> int:  # bytes LEFT (:101)
        return self._file_size - self._position

    file_remaining = length

    @property
    def position(self) -> int:
        return self._position

    @position.setter
    def position(self, v: int):
        self.set_position(v)

    def set_position(self, v: int):  # (:110-116)
        if v < self._position:
            self.rewind(self._position - v)
        elif v > self._position:
            self.skip(v - self._position)

    @property
    def is_eos(self) -> bool:
        return self._position >= self._file_size

    def reset(self):
        self._position = 0

    def skip(self, length: int):
        self._position += length

    def rewind(self, length: int = 1):
        self._position = max(0, self._position - length)

    def subset(self, position: int | None = None, length: int | None = None, buffer_size: int | None = None):
        return InputFileStream.from_file_stream(self, position=position, length=length, buffer_size=buffer_size)

    def _pread(self, n: int, at: int) -> bytes:
        n = max(0, min(n, self._file_size - at))  # FileBuffer reads stop at the stream's end (file_buffer.dart:98-117)
        return os.pread(self._fd_box[0], n, self._file_offset + at) if n else b""

    def _read_uint(self, size: int) -> int:
        if self.is_eos:
            return 0
        b = self._pread(size, self._position).ljust(size, b"\0")
        self._position += size
        return int.from_bytes(b, "big" if self.byte_order == BIG_ENDIAN else "little")

    def read_byte(self) -> int:  # 0 at the end of the stream (:144-151), unlike InputMemoryStream
        return self._read_uint(1)

    def read_uint16(self) -> int:
        return self._read_uint(2)

    def read_uint24(self) -> int:
        return self._read_uint(3)

    def read_uint32(self) -> int:
        return self._read_uint(4)

    def read_uint64(self) -> int:
        return self._read_uint(8)

    def read_bytes(self, count: int) -> "InputFileStream":  # (:196-207)
        if self.is_eos:
            return InputFileStream.from_file_stream(self, length=0)
        count = min(count, self._file_size - self._position)
        s = InputFileStream.from_file_stream(self, position=self._position, length=count)
        self._position += s.length
        return s

    def peek_bytes(self, count: int, offset: int = 0) -> "InputFileStream":  # input_stream.dart:115-121
        keep = self._position
        self.skip(offset)
        s = self.read_bytes(count)
        self._position = keep
        return s

    def to_uint8_list(self) -> bytes:  # the REST of the stream, position unchanged (:210-216)
        return b"" if self.is_eos else self._pread(self.length, self._position)

    def file_range(self):
        """(path, first byte, byte count) of what is left: what the file entry point is given instead of bytes."""
        return self.path, self._file_offset + self._position, max(0, self.length)


class OutputFileStream:
    """OutputFileStream (lib/src/util/output_file_stream.dart:11-235): the file is created / truncated when the stream is
    made (FileHandle(path, mode: write), _file_handle_io.dart:23-37); small writes collect in a `buffer_size` buffer."""
    default_buffer_size = 1024 * 1024

    def __init__(self, path: str, byte_order: int = LITTLE_ENDIAN, buffer_size: int | None = None):
        self.path = path
        self.byte_order = byte_order
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)  # createSync(recursive: true)
        self._fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        self._length = 0
        self._buf = bytearray()
        self._cap = self.default_buffer_size if buffer_size is None else max(1, buffer_size)

    @property
    def is_open(self) -> bool:
        return self._fd is not None

    @property
    def length(self) -> int:
        return self._length

    def flush(self):
        if self._buf:
            if self.is_open:
                os.write(self._fd, self._buf)
            del self._buf[:]

    def close_sync(self):
        if not self.is_open:
            return
        self.flush()
        os.close(self._fd)
        self._fd = None

    close = clear = close_sync

    def write_byte(self, v: int):
        self._buf.append(v & 0xff)
        if len(self._buf) == self._cap:
            self.flush()
        self._length += 1

    def write_bytes(self, data, length: int | None = None):  # (:110-127)
        data = bytes(data if length is None else data[:length])
        n = len(data)
        if len(self._buf) + n >= self._cap:
            self.flush()
        if len(self._buf) + n < self._cap:
            self._buf += data
        else:
            os.write(self._fd, data)
        self._length += n

    def write_stream(self, stream):  # (:130-144): 1 MiB pieces
        size = stream.length
        while size > 0:
            piece = stream.read_bytes(min(size, 1024 * 1024)).to_uint8_list()
            self.write_bytes(piece)
            size -= min(size, 1024 * 1024)

    def write_uint16(self, v: int):
        self.write_bytes(int(v & 0xffff).to_bytes(2, "big" if self.byte_order == BIG_ENDIAN else "little"))

    def write_uint32(self, v: int):
        self.write_bytes(int(v & 0xffffffff).to_bytes(4, "big" if self.byte_order == BIG_ENDIAN else "little"))

    def write_uint64(self, v: int):
        self.write_bytes(int(v & 0xffffffffffffffff).to_bytes(8, "big" if self.byte_order == BIG_ENDIAN else "little"))

    def subset(self, start: int, end: int | None = None) -> bytes:  # (:198-234) bytes already written
        self.flush()
        pos = self._length
        if start < 0:
            start += pos
        if end is None:
            end = pos
        elif end < 0:
            end += pos
        return os.pread(self._fd, max(0, end - start), start)

    def file_tail(self):
        """(path, byte offset at which the next byte goes), everything buffered flushed: for the file entry point."""
        self.flush()
        return self.path, self._length

    def advanced(self, n: int):
        """The library has written n bytes at file_tail()'s offset."""
        self._length += n
        os.lseek(self._fd, self._length, os.SEEK_SET)
