"""Pins the oracle against every golden vector / known-answer test the reference's own tests hold for the
inflate path: test/inflate_test.dart, test/gzip_test.dart, test/zlib_test.dart,
test/adler32_test.dart, test/crc32_test.dart.  CPU only."""
import hashlib
import json
import os
import zlib

import oracle_lib as orc

G = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))


def rd(n):
    return open(os.path.join(G, n), "rb").read()


def test_inflate_data_bin():  # test/inflate_test.dart:14-20
    st, out, used = orc.inflate(rd("inflate_data.bin"))
    assert st == orc.OK
    assert len(out.decode("utf8")) == 5259
    assert hashlib.sha256(out).hexdigest() == MAN["inflate_data.bin"]["sha256"]
    assert used == 1102


def test_gzip_fixtures():  # test/gzip_test.dart:63-93
    for name in ("cat.jpg.gz", "test2.tar.gz", "a.txt.gz"):
        st, out = orc.gzip_decode(rd(name))
        assert st == orc.OK
        assert len(out) == MAN[name]["size"]
        assert hashlib.sha256(out).hexdigest() == MAN[name]["sha256"]
    assert orc.gzip_decode(rd("cat.jpg.gz"))[1] == rd("cat.jpg")
    assert orc.gzip_decode(rd("a.txt.gz"))[1] == ("this is a test\nof the\nzip archive\nformat.\n" * 3).encode()


def test_git_vector():  # test/inflate_test.dart:57-60,64-179 (disabled there; "only 148 bytes consumed")
    data = rd("git_inflate_input.bin")
    st, out, used = orc.inflate(data[2:])  # skip the 2-byte zlib header
    assert out == rd("git_expected_output.bin")
    assert used + 2 + 4 == MAN["git_inflate_input.bin"]["consumed"]


def test_zlib_multistream():  # test/zlib_test.dart:15-23 (ZLibDecoderWeb, verify: true)
    data = zlib.compress(bytes([1, 2, 3])) + zlib.compress(bytes([4, 5, 6]))
    st, out = orc.zlib_decode(data, verify=True)
    assert st == orc.OK and out == bytes([1, 2, 3, 4, 5, 6])


def test_gzip_multimember():  # test/gzip_test.dart:44-52
    import gzip
    data = gzip.compress(bytes([1, 2, 3])) + gzip.compress(bytes([4, 5, 6]))
    st, out = orc.gzip_decode(data, verify=True)
    assert st == orc.OK and out == bytes([1, 2, 3, 4, 5, 6])


def test_roundtrip_buffers():  # test/zlib_test.dart:25-55 / test/deflate_test.dart:12-44 (decode side, system zlib encodes)
    buf = bytes(i % 256 for i in range(0xfffff))
    for level in (0, 1, 9):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        z = co.compress(buf) + co.flush()
        st, out, used = orc.inflate(z)
        assert st == orc.OK and out == buf and used == len(z)


def test_checksum_kats():  # test/adler32_test.dart:6-24, test/crc32_test.dart:6-24
    k = MAN["kat"]
    ten = bytes([1, 2, 3, 4, 5, 6, 7, 8, 9, 0])
    assert orc.adler32(b"") == k["adler32"]["empty"]
    assert orc.adler32(bytes([1])) == k["adler32"]["one"]
    assert orc.adler32(ten) == k["adler32"]["ten"]
    assert orc.crc32(b"") == k["crc32"]["empty"]
    assert orc.crc32(bytes([1])) == k["crc32"]["one"]
    assert orc.crc32(ten) == k["crc32"]["ten"]
    a, c = orc.adler32(b""), orc.crc32(b"")
    for _ in range(10000):
        a, c = orc.adler32(ten, a), orc.crc32(ten, c)
    assert a == k["adler32"]["hundred_k"] and c == k["crc32"]["hundred_k"]


def test_quirk_q1_short_tail():
    """inflate.dart:192-195: _readCodeByTable wants maxCodeLength bits; a raw stream that ends exactly at its
    last byte loses the EOB (and, with long codes, trailing symbols).  Pads of >= 2 bytes never trigger it."""
    text = (b"abcdefghij" * 40 + bytes(range(256))) * 3
    co = zlib.compressobj(9, zlib.DEFLATED, -15)
    z = co.compress(text) + co.flush()
    st, out, used = orc.inflate(z + b"\0\0")
    assert out == text and used == len(z)
    st, out2, _ = orc.inflate(z)
    assert text.startswith(out2)  # prefix, possibly short
