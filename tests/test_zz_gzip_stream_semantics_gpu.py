"""GZip / zlib framing semantics of the reference that only show on unusual or damaged streams, on the sm_100a path against
the oracle:
  * the members of a gzip stream share ONE OutputStream (_gzip_decoder_web.dart:38: Inflate.stream(input, output: output)),
    so a member's back-references may reach into the members before it (output_memory_stream.dart:79-98 checks against the
    whole stream) -- members compressed against a preset dictionary that equals the preceding output decode fine;
  * a stream that ends inside a DEFLATE block makes the reference read its trailer past the end of the input: RangeError,
    not `false` (_gzip_decoder_web.dart:40-41, _zlib_decoder_web.dart:86).
The file sorts last with the other late additions; each test has a hard time limit."""
import os
import struct
import zlib

import pytest

import oracle_lib as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


def member(chunk: bytes, zdict: bytes | None = None, hint: bool = False) -> bytes:
    co = (zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_DEFAULT_STRATEGY, zdict) if zdict
          else zlib.compressobj(6, zlib.DEFLATED, -15, 9))
    body = co.compress(chunk) + co.flush()
    trailer = struct.pack("<II", zlib.crc32(chunk), len(chunk))
    if hint:
        total = 10 + 2 + 6 + len(body) + 8
        return (b"\x1f\x8b\x08\x04" + bytes(4) + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1)
                + body + trailer)
    return b"\x1f\x8b\x08\x00" + bytes(4) + b"\x00\xff" + body + trailer


def run(a, dec, z, **kw):
    out = a.OutputMemoryStream()
    try:
        ok = dec.decode_stream(a.InputMemoryStream(z), out, **kw)
        return (orc.OK if ok else orc.FALSE), out.get_bytes()
    except a.DartRangeError:
        return orc.THROW, out.get_bytes()


def chained(text: bytes, cuts, hint):
    """Members whose matches reach into the output of the members before them (preset dictionary = that output)."""
    ms, done = [], b""
    for lo, hi in cuts:
        ms.append(member(text[lo:hi], zdict=done[-32768:] or None, hint=hint))
        done += text[lo:hi]
    return ms, done


def test_members_share_one_output_stream(a):
    from archive_b200 import synth
    text = synth.text(90000, stream=50).tobytes()
    cuts = [(0, 20000), (15000, 40000), (30000, 60000), (100, 9000), (50000, 90000)]  # overlapping: plenty to copy from
    for hint in (False, True):
        ms, want = chained(text, cuts, hint)
        blob = b"".join(ms)
        assert sum(len(m) for m in ms[1:]) < sum(len(member(text[lo:hi])) for lo, hi in cuts[1:])  # the dictionary is used
        assert orc.gzip_decode(blob) == (orc.OK, want)
        assert run(a, a.GZipDecoder(), blob) == (orc.OK, want), hint
        # on its own, such a member reaches before the start of its output: RangeError, as in the reference
        ost, _ = orc.gzip_decode(ms[1])
        assert ost == orc.THROW and run(a, a.GZipDecoder(), ms[1])[0] == orc.THROW
    # a plain member between chained ones, and the zlib streams of ZLibDecoder, which do NOT share (each Inflate has a
    # buffer of its own, _zlib_decoder_web.dart:82-84)
    ms, want = chained(text, cuts[:3], True)
    blob = ms[0] + member(text[:5000], hint=True) + ms[1]
    ost, oout = orc.gzip_decode(blob)
    assert run(a, a.GZipDecoder(), blob) == (ost, oout)


def test_chained_members_through_the_file_path(a, tmp_path, monkeypatch):
    from archive_b200 import synth
    text = synth.text(400000, stream=51).tobytes()
    cuts = [(i * 9000, i * 9000 + 16000) for i in range(40)]
    ms, want = chained(text, cuts, True)
    plain = [member(text[lo:hi], hint=True) for lo, hi in cuts[:12]]
    blob = b"".join(plain + ms[1:])  # a hinted run the segment pipeline takes, then members that need what came before
    ost, oout = orc.gzip_decode(blob)
    mem = run(a, a.GZipDecoder(), blob)
    assert mem == (ost, oout)
    for seg in ("64", "100", None):
        if seg:
            monkeypatch.setenv("B200Z_FILE_SEG_KB", seg)
        else:
            monkeypatch.delenv("B200Z_FILE_SEG_KB", raising=False)
        src, dst = str(tmp_path / "c.gz"), str(tmp_path / "c.out")
        open(src, "wb").write(blob)
        out = a.OutputFileStream(dst)
        try:
            ok = a.GZipDecoder().decode_stream(a.InputFileStream(src), out)
            st = orc.OK if ok else orc.FALSE
        except a.DartRangeError:
            st = orc.THROW
        out.close_sync()
        assert (st, open(dst, "rb").read()) == mem, seg


def test_truncated_streams_throw_like_the_reference(a):
    from archive_b200 import synth
    text = synth.text(3 * 8192, stream=49).tobytes()
    blob = b"".join(member(text[i:i + 8192], hint=(i == 0)) for i in range(0, len(text), 8192))
    cuts = list(range(0, len(blob), 211)) + [len(blob) - k for k in range(1, 12)]
    for cut in cuts:
        z = blob[:cut]
        ost, oout = orc.gzip_decode(z)
        st, got = run(a, a.GZipDecoder(), z)
        assert st == ost and (st == orc.THROW or got == oout), ("gzip", cut, st, ost)
        if st == orc.THROW:
            assert got == oout[:len(got)]  # what was written before the exception is a prefix of the reference's
    zb = zlib.compress(text[:9000], 6) + zlib.compress(text[9000:20000], 9)
    for cut in list(range(0, len(zb), 173)) + [len(zb) - k for k in range(1, 8)]:
        for verify in (False, True):
            ost, oout = orc.zlib_decode(zb[:cut], verify=verify)
            st, got = run(a, a.ZLibDecoder(), zb[:cut], verify=verify)
            assert st == ost and (st == orc.THROW or got == oout), ("zlib", cut, verify, st, ost)


def test_impossible_size_fields_are_not_hints(a):
    """ISIZE is never looked at by the reference (_gzip_decoder_web.dart:40-41 reads and drops it).  Here it is a size hint
    that is verified afterwards -- but a value DEFLATE cannot reach from the member's bytes (more than 1032:1) must not even
    size a buffer: the member is decoded the hint-free way and the bytes are the reference's."""
    from archive_b200 import _ffi, synth
    text = synth.text(20 * 8192, stream=54).tobytes()
    ms = [member(text[i:i + 8192], hint=True) for i in range(0, len(text), 8192)]
    for k in (0, 7, 19):
        bad = bytearray(ms[k])
        bad[-4:] = struct.pack("<I", 0xFFFFFFF0)
        blob = b"".join(ms[:k] + [bytes(bad)] + ms[k + 1:])
        L = _ffi.ensure_init()
        addr, n, keep = _ffi.as_buffer(blob)
        assert L.b200z_gzip_bound(addr, n) == 0  # unknown, not 4 GiB
        assert orc.gzip_decode(blob) == (orc.OK, text)
        assert run(a, a.GZipDecoder(), blob) == (orc.OK, text), k
