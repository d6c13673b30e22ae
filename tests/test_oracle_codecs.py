"""Oracle checks for the encoder and BZip2 rows of SURVEY.md section 8(c).  The reference's tests pin these only
by round trip (F6: "parity unpinned"), so on top of the round trips the oracle is cross-checked against the
native libraries it descends from: stock zlib (identical once the reference's extra early-flush heuristic
is switched off) and libbz2.  CPU only."""
import bz2
import gzip
import hashlib
import json
import os
import random
import zlib

import pytest

import oracle_lib as orc
from archive_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))


def rd(n):
    return open(os.path.join(G, n), "rb").read()


def zl(d, level, wb=15):
    co = zlib.compressobj(level, zlib.DEFLATED, -wb, 8)
    return co.compress(d) + co.flush()


@pytest.fixture(scope="module")
def corpus():
    rng = random.Random(5)
    t = synth.text(1 << 20).tobytes()
    return {
        "text": t, "empty": b"", "one": b"a", "zeros": b"\0" * 200000,
        "rand": bytes(rng.getrandbits(8) for _ in range(100000)),
        "mod256": bytes(i % 256 for i in range(0xfffff)),  # test/deflate_test.dart:12-44
        "mix": t[:50000] + bytes(rng.getrandbits(8) for _ in range(50000)) + t[:50000],
    }


def test_deflate_roundtrip_all_levels(corpus):  # test/deflate_test.dart:12-44 (levels 0/1/9 there)
    for name, d in corpus.items():
        for level in range(10):
            st, z, crc = orc.deflate(d, level)
            assert st == orc.OK
            assert zlib.decompress(z, -15) == d, (name, level)
            assert orc.inflate(z + b"\0\0")[1] == d, (name, level)
            assert crc == zlib.crc32(d)  # Deflate.crc32 (deflate.dart:31,1231)


def test_deflate_equals_stock_zlib_without_the_heuristic(corpus):
    """with TRUNCATE_BLOCK (deflate.dart:549-562) off, levels 1-9 must be byte-identical to
    zlib.compressobj(level, DEFLATED, -wbits, memLevel 8); any mismatch is an oracle bug."""
    orc.set_truncate_heuristic(False)
    try:
        for name in ("text", "mix", "rand", "zeros", "one", "empty"):
            d = corpus[name]
            for level in range(1, 10):
                for wb in (15, 11, 9):
                    assert orc.deflate(d, level, wb)[1] == zl(d, level, wb), (name, level, wb)
    finally:
        orc.set_truncate_heuristic(True)


def test_deflate_heuristic_changes_text_blocks(corpus):
    """F4: with the heuristic ON (the reference's behaviour) text at level >= 4 differs from stock zlib."""
    assert orc.deflate(corpus["text"], 6)[1] != zl(corpus["text"], 6)
    assert orc.deflate(corpus["rand"], 6)[1] == zl(corpus["rand"], 6)


def test_deflate_invalid_params():
    assert orc.deflate(b"abc", 10)[0] == orc.THROW
    assert orc.deflate(b"abc", 6, 8)[0] == orc.THROW


def test_zlib_gzip_encoder_framing(corpus):  # _zlib_encoder_web.dart:27-73, _gzip_encoder_web.dart:27-100
    d = corpus["text"][:100000]
    st, z = orc.zlib_encode(d, 6)
    assert z[:2] == b"\x78\x01"  # quirk Q4: FLEVEL 0 at every level
    assert zlib.decompress(z) == d
    assert orc.zlib_decode(z, verify=True) == (orc.OK, d)
    st, g = orc.gzip_encode(d, 6, mtime=0)
    assert g[:10] == b"\x1f\x8b\x08\x00\0\0\0\0\x00\xff"
    assert gzip.decompress(g) == d
    assert orc.gzip_decode(g)[1] == d
    # test/zlib_test.dart:15-23 and test/gzip_test.dart:44-52 with the oracle's OWN encoders
    two = orc.zlib_encode(bytes([1, 2, 3]))[1] + orc.zlib_encode(bytes([4, 5, 6]))[1]
    assert orc.zlib_decode(two, verify=True) == (orc.OK, bytes([1, 2, 3, 4, 5, 6]))
    two = orc.gzip_encode(bytes([1, 2, 3]))[1] + orc.gzip_encode(bytes([4, 5, 6]))[1]
    assert orc.gzip_decode(two, verify=True) == (orc.OK, bytes([1, 2, 3, 4, 5, 6]))


def test_bzip2_decode_fixtures():  # test/bzip2_test.dart:8-12, io_test (test2.tar.bz2), zip_bzip2.zip members
    for name in ("test.bz2", "test2.tar.bz2"):
        st, out = orc.bzip2_decode(rd(name), verify=True)
        assert st == orc.OK
        assert len(out) == MAN[name]["size"] and hashlib.sha256(out).hexdigest() == MAN[name]["sha256"]
    assert orc.bzip2_decode(rd("test2.tar.bz2"))[1] == rd("test2.tar")


def test_bzip2_decode_vs_libbz2(corpus):
    rng = random.Random(9)
    big = synth.text(2_500_000).tobytes()  # > 2 blocks at level 9, many at level 1
    for d in (big, corpus["empty"], corpus["one"], b"aaaa" * 100000, corpus["rand"], big[:1000],
              bytes([251]) * 70000, bytes(rng.choice(b"ab") for _ in range(50000))):
        for level in (1, 9):
            z = bz2.compress(d, level)
            assert orc.bzip2_decode(z, verify=True) == (orc.OK, d)


def test_bzip2_decode_errors():
    z = bz2.compress(synth.text(100000).tobytes())
    assert orc.bzip2_decode(b"BZx9" + z[4:])[0] == orc.FALSE  # bad signature -> false (bzip2_decoder.dart:29-33)
    assert orc.bzip2_decode(z[:len(z) // 2])[0] == orc.THROW  # readByte past the end
    bad = bytearray(z)
    bad[len(z) // 2] ^= 0x10
    st, out = orc.bzip2_decode(bytes(bad), verify=True)
    assert st in (orc.FALSE, orc.THROW)
    # stops at the first end-of-stream block: a second concatenated stream is NOT decoded (:83-84)
    st, out = orc.bzip2_decode(z + z)
    assert st == orc.OK and out == bz2.decompress(z)


def test_bzip2_encode_single_block_equals_libbz2(corpus):
    """the Dart encoder descends from libbzip2; for inputs that fit ONE block the oracle's output
    must equal bz2.compress(x, 9) byte for byte (main sort, fallback sort < 10000, periodic blocks that exhaust the
    work budget, RLE1 corner cases)."""
    rng = random.Random(11)
    t = corpus["text"]
    cases = [t[:880000], b"", b"a", b"ab" * 5, b"aaaa" * 100000, bytes([251]) * 70000, corpus["rand"], t[:1000], t[:9999],
             t[:10000], rd("cat.jpg"), bytes(rng.choice(b"ab") for _ in range(50000)), t[:37] * 20000, b"\0" * 800000,
             b"".join(bytes([rng.randrange(4)]) * rng.choice([1, 3, 4, 5, 8, 9, 10, 255, 256, 259, 260]) for _ in range(3000))]
    for d in cases:
        st, z = orc.bzip2_encode(d)
        assert st == orc.OK and z == bz2.compress(d, 9), len(d)


def test_bzip2_encode_multiblock_roundtrip_and_block_boundary_quirk():
    """test/bzip2_test.dart:14-25 is a round trip.  Across block boundaries the reference flushes the pending RLE1 run
    at the end of EVERY block (_writeBlock, bzip2_encoder.dart:97-103) where libbzip2 carries it over, so multi-block
    output legitimately differs from libbz2 while decoding to the same bytes."""
    d = synth.text(2_000_000, stream=3).tobytes()
    st, z = orc.bzip2_encode(d)
    assert st == orc.OK and bz2.decompress(z) == d
    assert orc.bzip2_decode(z, verify=True) == (orc.OK, d)
    assert z != bz2.compress(d, 9)
    assert z[:4] == b"BZh9"
