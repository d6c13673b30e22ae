"""Host-side mirror of the reference's codec classes (same names, argument meaning and error
behaviour), calling the sm_100a kernels through the C ABI.

Mirrors (paths relative to /root/reference/):
  Inflate            lib/src/codecs/zlib/inflate.dart:12-116
  ZLibDecoder(Web)   lib/src/codecs/zlib_decoder.dart:14-35, codecs/zlib/_zlib_decoder_web.dart:14-107
  GZipDecoder(Web)   lib/src/codecs/gzip_decoder.dart:14-30, codecs/zlib/_gzip_decoder_web.dart:14-58
  inflateBuffer      lib/src/codecs/zlib/inflate_buffer.dart:7

In the Dart package these classes stay Dart and bind libb200z.so with dart:ffi (dart/);
no Dart SDK exists in the build image, so the parity tests drive this Python mirror instead.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _ffi
from .streams import BIG_ENDIAN, InputFileStream, InputMemoryStream, OutputFileStream, OutputMemoryStream


def _rest(input):
    """The rest of an input stream as one bytes-like object (InputStream.toUint8List)."""
    if isinstance(input, InputFileStream):
        return input.to_uint8_list()
    return input.buffer[input.position:]


def _consume(input):
    """decodeStream / encodeStream read their input to its end."""
    if isinstance(input, InputFileStream):
        input.skip(max(0, input.length))
    else:
        input.position = len(input.buffer)


def _both_files(input, output) -> bool:
    return isinstance(input, InputFileStream) and isinstance(output, OutputFileStream)


def _file_codec(op: int, input: InputFileStream, output: OutputFileStream, a0: int = 0, a1: int = 0, a2: int = 0) -> int:
    """InputFileStream -> codec -> OutputFileStream without the bytes passing through the host language: the library gets
    the two paths and byte ranges (b200z_file_codec, include/b200z.h; csrc/b200z_file.cu)."""
    L = _ffi.ensure_init()
    path, off, n = input.file_range()
    opath, ooff = output.file_tail()
    used, got = C.c_uint64(0), C.c_uint64(0)
    rc = L.b200z_file_codec(op, os.fsencode(path), off, n, os.fsencode(opath), ooff, a0, a1, a2 & 0xFFFFFFFF,
                            C.byref(used), C.byref(got))
    output.advanced(got.value)
    input.skip(n)
    return rc


def _stream_result(rc: int) -> bool:
    if rc == _ffi.E_DATA:
        return False
    _ffi.check(rc)
    return True


def _grow_call(fn, in_addr, in_len, first_cap):
    """Call fn(out_addr, cap) -> (rc, out_len); retry with a larger buffer on E_NOSPC."""
    cap = max(first_cap, 1 << 12)
    while True:
        out = (C.c_uint8 * cap)()
        rc, n = fn(C.addressof(out), cap)
        if rc == _ffi.E_NOSPC and cap < (1 << 40):
            cap = max(cap * 2, n + (n >> 3))
            continue
        return rc, out, n


class Inflate:
    """`Inflate(bytes)` / `Inflate.stream(input, output:)`: all work happens in the constructor
    (inflate.dart:23-40); bad data never raises -- decoding stops and the partial output is kept
    (inflate.dart:150-151); `RangeError` cases raise DartRangeError."""

    def __init__(self, data=None, output: OutputMemoryStream | None = None, uncompressed_size: int | None = None,
                 _input: InputMemoryStream | None = None):
        self._input = _input if _input is not None else InputMemoryStream(data if data is not None else b"")
        self._output = output if output is not None else OutputMemoryStream(size=uncompressed_size)
        self.status = _ffi.U_EOS
        self._inflate(uncompressed_size)

    @classmethod
    def stream(cls, input: InputMemoryStream | None, output: OutputMemoryStream | None = None,
               uncompressed_size: int | None = None) -> "Inflate":
        return cls(None, output=output, uncompressed_size=uncompressed_size,
                   _input=input if input is not None else InputMemoryStream(b""))

    def _inflate(self, size_hint):
        L = _ffi.ensure_init()
        view = _rest(self._input)
        if len(view) == 0:
            return
        addr, n, keep = _ffi.as_buffer(view)
        out_len, used, ust = C.c_size_t(0), C.c_size_t(0), C.c_int32(0)

        def call(out_addr, cap):
            rc = L.b200z_inflate_raw(addr, n, out_addr, cap, C.byref(out_len), C.byref(used), C.byref(ust))
            return rc, out_len.value

        rc, out, got = _grow_call(call, addr, n, size_hint or 4 * n + 1024)
        self.status = ust.value
        if got:
            self._output.write_bytes(C.string_at(out, got))
        self._input.position += used.value  # (a file stream's setter skips forward)
        _ffi.check(rc)

    def get_bytes(self) -> bytes:
        return self._output.get_bytes()


def inflate_buffer(data) -> bytes:  # inflate_buffer.dart:7 (web variant: Inflate(data).getBytes())
    return Inflate(data).get_bytes()


class _FramedDecoder:
    _fn = None
    _has_raw = True

    def decode_bytes(self, data, verify: bool = False, raw: bool = False) -> bytes:
        """decodeBytes ignores decodeStream's bool and returns whatever was written
        (_gzip_decoder_web.dart:19-24, _zlib_decoder_web.dart:21-28)."""
        out = OutputMemoryStream()
        self.decode_stream(InputMemoryStream(data), out, verify=verify, raw=raw)
        return out.get_bytes()

    def decode_stream(self, input: InputMemoryStream, output: OutputMemoryStream, verify: bool = False,
                      raw: bool = False) -> bool:
        if _both_files(input, output):
            a0 = int(bool(verify)) | (2 if raw and self._file_op == _ffi.FILE_GZIP_DECODE else 0)  # B200Z_GZIP_RAW
            return _stream_result(_file_codec(self._file_op, input, output, a0, int(raw)))
        L = _ffi.ensure_init()
        view = _rest(input)
        addr, n, keep = _ffi.as_buffer(view)
        out_len = C.c_size_t(0)
        rc, out, got = _grow_call(lambda oa, cap: self._call(L, addr, n, verify, raw, oa, cap, out_len),
                                  addr, n, self._first_cap(L, addr, n))
        if got:
            output.write_bytes(C.string_at(out, got))
        _consume(input)
        return _stream_result(rc)


class ZLibDecoderWeb(_FramedDecoder):
    _file_op = _ffi.FILE_ZLIB_DECODE

    def _first_cap(self, L, addr, n):
        return 4 * n + 1024

    def _call(self, L, addr, n, verify, raw, oa, cap, out_len):
        rc = L.b200z_zlib_decode(addr, n, int(verify), int(raw), oa, cap, C.byref(out_len))
        return rc, out_len.value


class GZipDecoderWeb(_FramedDecoder):
    _file_op = _ffi.FILE_GZIP_DECODE

    def _first_cap(self, L, addr, n):
        return L.b200z_gzip_bound(addr, n) or 4 * n + 1024

    def _call(self, L, addr, n, verify, raw, oa, cap, out_len):
        rc = L.b200z_gzip_decode(addr, n, int(bool(verify)) | (2 if raw else 0), oa, cap, C.byref(out_len))  # B200Z_GZIP_RAW
        return rc, out_len.value


# The platform-dispatched names (zlib_decoder.dart:14, gzip_decoder.dart:14) bind the same backend:
# this is the `platformZLibDecoder` / `platformGZipDecoder` seam (_zlib_decoder.dart:1, _gzip_decoder.dart:1).
ZLibDecoder = ZLibDecoderWeb
GZipDecoder = GZipDecoderWeb


class BZip2Decoder:
    """BZip2Decoder().decodeBytes / decodeStream (lib/src/codecs/bzip2_decoder.dart:12-88): CRCs are compared only
    when `verify`; decodeStream returns False on any data error and keeps the blocks decoded before it."""

    def decode_bytes(self, data, verify: bool = False) -> bytes:
        out = OutputMemoryStream()
        self.decode_stream(InputMemoryStream(data), out, verify=verify)
        return out.get_bytes()

    def decode_stream(self, input: InputMemoryStream, output: OutputMemoryStream, verify: bool = False) -> bool:
        if _both_files(input, output):
            return _stream_result(_file_codec(_ffi.FILE_BZIP2_DECODE, input, output, int(verify)))
        L = _ffi.ensure_init()
        view = _rest(input)
        addr, n, keep = _ffi.as_buffer(view)
        out_len = C.c_size_t(0)

        def call(oa, cap):
            rc = L.b200z_bzip2_decode(addr, n, int(verify), oa, cap, C.byref(out_len))
            return rc, out_len.value

        rc, out, got = _grow_call(call, addr, n, 6 * n + 4096)
        if got:
            output.write_bytes(C.string_at(out, got))
        _consume(input)
        return _stream_result(rc)


class BZip2Encoder:
    """BZip2Encoder().encodeBytes / encode / encodeStream (lib/src/codecs/bzip2_encoder.dart:15-81): one "BZh9"
    stream; encodeStream returns True."""

    def encode_bytes(self, data) -> bytes:
        out = OutputMemoryStream()
        self.encode_stream(InputMemoryStream(data), out)
        return out.get_bytes()

    encode = encode_bytes

    def encode_stream(self, input: InputMemoryStream, output: OutputMemoryStream) -> bool:
        if _both_files(input, output):
            _ffi.check(_file_codec(_ffi.FILE_BZIP2_ENCODE, input, output))
            return True
        L = _ffi.ensure_init()
        view = _rest(input)
        addr, n, keep = _ffi.as_buffer(view)
        cap = L.b200z_bzip2_bound(n)
        out = (C.c_uint8 * cap)()
        out_len = C.c_size_t(0)
        _ffi.check(L.b200z_bzip2_encode(addr, n, C.addressof(out), cap, C.byref(out_len)))
        output.write_bytes(C.string_at(out, out_len.value))
        _consume(input)
        return True


class Deflate:
    """`Deflate(bytes, level: 6, windowBits: 15)` (lib/src/codecs/zlib/deflate.dart:25-100): raw DEFLATE produced in the
    constructor, `get_bytes()` / `take_bytes()`, and `crc32` of the consumed input.  Invalid parameters make the
    reference's `getBytes()` throw LateInitializationError (`_init` returned false, :107-118): B200ZError(E_ARG) here."""

    def __init__(self, data=b"", level: int = 6, window_bits: int = 15, output: OutputMemoryStream | None = None):
        self._output = output if output is not None else OutputMemoryStream()
        self.level = level
        self.crc32 = 0
        L = _ffi.ensure_init()
        addr, n, keep = _ffi.as_buffer(data)
        cap = L.b200z_deflate_bound(n)
        out = (C.c_uint8 * cap)()
        out_len, crc = C.c_size_t(0), C.c_uint32(0)
        rc = L.b200z_deflate_raw(addr, n, level, window_bits, C.addressof(out), cap, C.byref(out_len), C.byref(crc))
        _ffi.check(rc)
        self.crc32 = crc.value
        self._output.write_bytes(C.string_at(out, out_len.value))

    @classmethod
    def stream(cls, input: InputMemoryStream, level: int = 6, window_bits: int = 15, output: OutputMemoryStream | None = None):
        """`Deflate.stream(input, level:, windowBits:, output:)` (deflate.dart:59-67): consumes the rest of `input`."""
        d = cls(_rest(input), level=level, window_bits=window_bits, output=output)
        _consume(input)
        return d

    def finish(self):  # deflate.dart:69 -- everything is already flushed when the constructor returns
        return None

    def get_bytes(self) -> bytes:
        return self._output.get_bytes()

    def take_bytes(self) -> bytes:
        b = self._output.get_bytes()
        self._output.clear()
        return b


class ZLibEncoderWeb:
    """ZLibEncoderWeb().encodeBytes(bytes, level:, windowBits:, raw:) -- _zlib_encoder_web.dart:17-73"""

    def encode_bytes(self, data, level: int | None = None, window_bits: int | None = None, raw: bool = False) -> bytes:
        L = _ffi.ensure_init()
        addr, n, keep = _ffi.as_buffer(data)
        cap = L.b200z_deflate_bound(n)
        out = (C.c_uint8 * cap)()
        out_len = C.c_size_t(0)
        rc = L.b200z_zlib_encode(addr, n, 6 if level is None else level, 15 if window_bits is None else window_bits, int(raw),
                                 C.addressof(out), cap, C.byref(out_len))
        _ffi.check(rc)
        return C.string_at(out, out_len.value)

    def encode_stream(self, input: InputMemoryStream, output: OutputMemoryStream, level: int | None = None,
                      window_bits: int | None = None, raw: bool = False) -> None:
        """_zlib_encoder_web.dart:30-73: the rest of `input` is consumed."""
        if _both_files(input, output):
            _ffi.check(_file_codec(_ffi.FILE_ZLIB_ENCODE, input, output, 6 if level is None else level,
                                   15 if window_bits is None else window_bits, int(raw)))
            return
        output.write_bytes(self.encode_bytes(_rest(input), level=level, window_bits=window_bits, raw=raw))
        _consume(input)


class GZipEncoderWeb:
    """GZipEncoderWeb().encodeBytes(bytes, level:) -- _gzip_encoder_web.dart:17-100.  The reference stamps MTIME with the
    wall clock (:82); pass `mtime` for reproducible bytes."""

    def encode_bytes(self, data, level: int | None = None, mtime: int | None = None) -> bytes:
        import time
        L = _ffi.ensure_init()
        addr, n, keep = _ffi.as_buffer(data)
        cap = L.b200z_deflate_bound(n)
        out = (C.c_uint8 * cap)()
        out_len = C.c_size_t(0)
        rc = L.b200z_gzip_encode(addr, n, 6 if level is None else level, int(time.time()) if mtime is None else mtime,
                                 C.addressof(out), cap, C.byref(out_len))
        _ffi.check(rc)
        return C.string_at(out, out_len.value)

    def encode_stream(self, input: InputMemoryStream, output: OutputMemoryStream, level: int | None = None,
                      mtime: int | None = None) -> None:
        """_gzip_encoder_web.dart:30-100: the rest of `input` is consumed."""
        if _both_files(input, output):
            import time
            _ffi.check(_file_codec(_ffi.FILE_GZIP_ENCODE, input, output, 6 if level is None else level, 0,
                                   int(time.time()) if mtime is None else mtime))
            return
        output.write_bytes(self.encode_bytes(_rest(input), level=level, mtime=mtime))
        _consume(input)


ZLibEncoder = ZLibEncoderWeb
GZipEncoder = GZipEncoderWeb
