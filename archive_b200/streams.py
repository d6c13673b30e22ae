"""InputMemoryStream / OutputMemoryStream -- the in-memory byte streams the reference's codec hot
path is written against (lib/src/util/input_memory_stream.dart:8-134,
lib/src/util/output_memory_stream.dart:8-137).  Only the members the codecs touch are mirrored."""
from __future__ import annotations

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
