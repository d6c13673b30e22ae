// The plug point of the reference is the conditional-export seam that binds
// `platformZLibDecoder` / `platformGZipDecoder` (lib/src/codecs/zlib/_zlib_decoder.dart:1,
// _gzip_decoder.dart:1) to an object implementing ZLibDecoderBase
// (lib/src/codecs/zlib/_zlib_decoder_base.dart:5-13).  These classes implement that interface on top
// of libb200z.so; `Inflate` mirrors lib/src/codecs/zlib/inflate.dart:12-116 and BZip2Decoder mirrors
// lib/src/codecs/bzip2_decoder.dart:12-21.
import 'dart:ffi';
import 'dart:typed_data';

import 'package:archive/archive.dart' as ar;
import 'package:ffi/ffi.dart';

import 'b200z_ffi.dart';

Uint8List _drain(ar.InputStream input) =>
    input is ar.InputMemoryStream ? input.toUint8List() : input.toUint8List();

/// Same surface as the reference's `Inflate`: all work in the constructor, never throws on bad data.
class Inflate {
  final ar.OutputStream _output;
  int status = 1;

  Inflate(List<int> bytes, {ar.OutputStream? output, int? uncompressedSize})
      : _output = output ?? ar.OutputMemoryStream(size: uncompressedSize) {
    _run(ar.InputMemoryStream(bytes), uncompressedSize);
  }

  Inflate.stream(ar.InputStream? input, {ar.OutputStream? output, int? uncompressedSize})
      : _output = output ?? ar.OutputMemoryStream(size: uncompressedSize) {
    if (input != null) _run(input, uncompressedSize);
  }

  void _run(ar.InputStream input, int? sizeHint) {
    final z = B200Z.instance;
    final data = _drain(input);
    if (data.isEmpty) return;
    final inp = z.toNative(data);
    final consumed = calloc<Size>();
    final ust = calloc<Int32>();
    try {
      final (out, _) = z.grow(sizeHint ?? data.length * 4 + 1024,
          (o, cap, outLen) => z.inflateRaw(inp, data.length, o, cap, outLen, consumed, ust));
      status = ust.value;
      _output.writeBytes(out);
      input.skip(consumed.value); // inflate.dart:337-340: stream left on the first unread byte
    } finally {
      calloc.free(consumed);
      calloc.free(ust);
      z.hostFree(inp);
    }
  }

  Uint8List getBytes() => _output.getBytes();
}

class _B200ZLibDecoder extends ar.ZLibDecoderBase {
  const _B200ZLibDecoder();

  @override
  Uint8List decodeBytes(List<int> data, {bool verify = false, bool raw = false}) {
    final z = B200Z.instance;
    final inp = z.toNative(data);
    try {
      final (out, _) = z.grow(data.length * 4 + 1024,
          (o, cap, outLen) => z.zlibDecode(inp, data.length, verify ? 1 : 0, raw ? 1 : 0, o, cap, outLen));
      return out;
    } finally {
      z.hostFree(inp);
    }
  }

  @override
  bool decodeStream(ar.InputStream input, ar.OutputStream output, {bool verify = false, bool raw = false}) {
    final z = B200Z.instance;
    final data = _drain(input);
    final inp = z.toNative(data);
    try {
      final (out, ok) = z.grow(data.length * 4 + 1024,
          (o, cap, outLen) => z.zlibDecode(inp, data.length, verify ? 1 : 0, raw ? 1 : 0, o, cap, outLen));
      output.writeBytes(out);
      input.skip(data.length);
      return ok;
    } finally {
      z.hostFree(inp);
    }
  }
}

class _B200GZipDecoder extends ar.ZLibDecoderBase {
  const _B200GZipDecoder();

  @override
  Uint8List decodeBytes(List<int> data, {bool verify = false, bool raw = false}) {
    final output = ar.OutputMemoryStream();
    decodeStream(ar.InputMemoryStream(data), output, verify: verify, raw: raw);
    return output.getBytes();
  }

  @override
  bool decodeStream(ar.InputStream input, ar.OutputStream output, {bool verify = false, bool raw = false}) {
    final z = B200Z.instance;
    final data = _drain(input);
    final inp = z.toNative(data);
    try {
      final bound = z.gzipBound(inp, data.length);
      final (out, ok) = z.grow(bound > 0 ? bound : data.length * 4 + 1024,
          (o, cap, outLen) => z.gzipDecode(inp, data.length, verify ? 1 : 0, o, cap, outLen));
      output.writeBytes(out);
      input.skip(data.length);
      return ok;
    } finally {
      z.hostFree(inp);
    }
  }
}

/// Drop these four names into a `_zlib_decoder_b200.dart` / `_gzip_decoder_b200.dart` selected by the
/// reference's conditional export (INTEGRATION.md).
const platformZLibDecoder = _B200ZLibDecoder();
const platformGZipDecoder = _B200GZipDecoder();

class BZip2Decoder {
  Uint8List decodeBytes(List<int> data, {bool verify = false}) {
    final z = B200Z.instance;
    final inp = z.toNative(data);
    try {
      final (out, _) = z.grow(data.length * 6 + 1024,
          (o, cap, outLen) => z.bzip2Decode(inp, data.length, verify ? 1 : 0, o, cap, outLen));
      return out;
    } finally {
      z.hostFree(inp);
    }
  }

  bool decodeStream(ar.InputStream input, ar.OutputStream output, {bool verify = false}) {
    final z = B200Z.instance;
    final data = _drain(input);
    final inp = z.toNative(data);
    try {
      final (out, ok) = z.grow(data.length * 6 + 1024,
          (o, cap, outLen) => z.bzip2Decode(inp, data.length, verify ? 1 : 0, o, cap, outLen));
      output.writeBytes(out);
      input.skip(data.length);
      return ok;
    } finally {
      z.hostFree(inp);
    }
  }
}
