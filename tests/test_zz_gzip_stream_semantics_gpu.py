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


def _gz(chunk, flags=0, extra=b"", name=b"", comment=b"", hcrc=False):
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 9)
    body = co.compress(chunk) + co.flush()
    f = flags | (4 if extra else 0) | (8 if name else 0) | (16 if comment else 0) | (2 if hcrc else 0)
    h = b"\x1f\x8b\x08" + bytes([f]) + bytes(4) + b"\x00\xff"
    if extra:
        h += struct.pack("<H", len(extra)) + extra
    if name:
        h += name + b"\0"
    if comment:
        h += comment + b"\0"
    if hcrc:
        h += b"\xab\xcd"
    return h + body + struct.pack("<II", zlib.crc32(chunk), len(chunk))


def test_gzip_header_variants_and_zlib_fallback(a):
    """_readHeader (_gzip_decoder_web.dart:60-138) with every optional field, size subfields that lie, reserved flag bits,
    junk between members -- and the zlib fall-back (:31-37), whose streams reach the output one stream late
    (_zlib_decoder_web.dart:82-84): a zlib stream followed by something that is not a zlib header is lost."""
    from archive_b200 import synth
    text = synth.text(40000, stream=55).tobytes()
    t = [text[i * 5000:(i + 1) * 5000] for i in range(8)]
    variants = [
        _gz(t[0], name=b"file.txt") + _gz(t[1], comment=b"a comment") + _gz(t[2], hcrc=True),
        _gz(t[0], extra=b"XY\x03\x00abc") + _gz(t[1], extra=b"BC\x02\x00\xff\xff") + _gz(t[2]),
        _gz(t[0], extra=b"AB\x01\x00zBC\x02\x00\x10\x00") + _gz(t[1]),
        _gz(t[0], flags=0xE0) + _gz(t[1], flags=0x20, name=b"n"),
        _gz(t[0]) + b"\x00\x00\x00" + _gz(t[1]),
        _gz(t[0]) + zlib.compress(t[1]) + _gz(t[2]),            # the zlib stream's bytes never reach the output
        _gz(t[0]) + zlib.compress(t[1]) + zlib.compress(t[2]),  # ... here the first one does
        _gz(t[0], extra=b"Q" * 300) + _gz(t[1], name=b"x" * 2000),
        _gz(t[0])[:-8] + _gz(t[1]),
        b"\x1f\x8b\x09\x00" + bytes(6) + _gz(t[0])[10:],
        _gz(t[0], name=b"unterminated")[:22],
        _gz(b""), _gz(b"") + _gz(t[3]), b"\x1f", b"\x1f\x8b", b"\x1f\x8b\x08",
    ]
    for i, z in enumerate(variants):
        ost, oout = orc.gzip_decode(z)
        st, got = run(a, a.GZipDecoder(), z)
        assert st == ost and (st == orc.THROW or got == oout), (i, st, ost, len(got), len(oout))
    # `raw` is handed on to the zlib decoder (:31-37): raw DEFLATE without any header inflates through GZipDecoder(raw: true)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    rawz = co.compress(t[4]) + co.flush()
    for z in (rawz, rawz + rawz, rawz[:-2]):
        for raw in (False, True):
            ost, oout = orc.gzip_decode(z, raw=raw)
            st, got = run(a, a.GZipDecoder(), z, raw=raw)
            assert st == ost and (st == orc.THROW or got == oout), (raw, st, ost, len(got), len(oout))
    assert orc.gzip_decode(rawz, raw=True)[1] == t[4]


def test_zlib_streams_reach_the_output_one_stream_late(a):
    from archive_b200 import synth
    text = synth.text(30000, stream=56).tobytes()
    s1, s2 = zlib.compress(text[:12000]), zlib.compress(text[12000:])
    raw = []
    for lvl, strat in ((6, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED)):
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
        raw.append(co.compress(text[:15000]) + co.flush())
    cases = [(s1 + s2, {}), (s1 + b"\x78", {}), (s1 + b"\x00\x00", {}), (s1 + b"\x78\x9d" + s2[2:], {}),  # bad FCHECK behind s1
             (s1 + s2[:-2], {}), (s1 + s2[:40], {}), (s1[:-1] + b"\x00" + s2, {"verify": True}),            # wrong Adler-32 in s1
             (raw[0] + raw[1], {"raw": True}), (raw[0] + raw[1] + b"\x00", {"raw": True}), (raw[0] + b"junkjunk", {"raw": True}),
             (raw[0][:-3], {"raw": True}), (b"", {"raw": True}), (b"", {})]
    for i, (z, kw) in enumerate(cases):
        for verify in (False, True):
            k = dict(kw)
            k.setdefault("verify", verify)
            ost, oout = orc.zlib_decode(z, **k)
            st, got = run(a, a.ZLibDecoder(), z, **k)
            assert st == ost and (st == orc.THROW or got == oout), (i, k, st, ost, len(got), len(oout))


def test_inflate_leaves_the_input_where_the_reference_does(a):
    """Inflate.stream(input): output AND the position the input is left at (inflate.dart:337-340 gives whole unread bytes
    back after a block; a read that runs out of input has pulled every byte, :166-168,192-195)."""
    import random
    from archive_b200 import synth
    text = synth.text(15000, stream=57).tobytes()
    rng = random.Random(11)
    for lvl, strat in ((6, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_HUFFMAN_ONLY)):
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
        z0 = co.compress(text) + co.flush()
        cands = [z0 + b"TRAILER!", z0] + [z0[:k] for k in range(0, len(z0), max(1, len(z0) // 12))]
        for z in cands:
            ost, oout, ocons = orc.inflate(z)
            ims = a.InputMemoryStream(z)
            try:
                got, st = a.Inflate.stream(ims).get_bytes(), orc.OK
            except a.DartRangeError:
                got, st = None, orc.THROW
            assert (st == orc.THROW) == (ost == orc.THROW), (lvl, len(z))
            if st != orc.THROW:
                assert got == oout and ims.position == ocons, (lvl, strat, len(z), ims.position, ocons)
