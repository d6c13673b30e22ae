"""Parity tests for the sm_100a Deflate encoder: compressed bytes IDENTICAL to the oracle (the line-by-line
restatement of deflate.dart) at the same level, plus the reference's own round-trip tests (test/deflate_test.dart:12-44,
test/zlib_test.dart:25-55, test/gzip_test.dart) with the GPU inflate on the other side.  The reference's tests pin the
encoder by round trip only ("parity unpinned"); byte identity is against the oracle."""
import gzip
import random
import zlib

import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


@pytest.fixture(scope="module")
def corpus():
    from archive_b200 import synth
    rng = random.Random(3)
    t = synth.text(3 << 20, stream=5).tobytes()
    return {
        "text": t, "empty": b"", "a": b"a", "ab": b"ab", "abc": b"abc", "zeros": b"\0" * 300000,
        "rand": bytes(rng.getrandbits(8) for _ in range(200000)), "mod256": bytes(i % 256 for i in range(0xfffff)),
        "short": t[:1000], "mix": t[:50000] + bytes(rng.getrandbits(8) for _ in range(50000)) + t[:50000],
        "rep": t[:700] * 500, "far": t[:40000] + bytes(rng.getrandbits(8) for _ in range(32500)) + t[:40000],
        "tail": t[:65536 - 3], "tail2": t[:65536 + 261], "seg": t[:32768 * 3 + 5],
    }


def test_byte_identical_to_oracle_levels_4_to_9(a, corpus):
    for name, d in corpus.items():
        for level in (4, 5, 6, 7, 8, 9):
            if level in (8, 9) and len(d) > (1 << 20):
                d = d[:1 << 20]
            df = a.Deflate(d, level=level)
            st, oz, ocrc = orc.deflate(d, level)
            assert df.get_bytes() == oz, (name, level)
            assert df.crc32 == ocrc == zlib.crc32(d), (name, level)


def test_level_0_stored(a, corpus):
    for name in ("text", "empty", "a", "mod256", "tail2"):
        d = corpus[name]
        assert a.Deflate(d, level=0).get_bytes() == orc.deflate(d, 0)[1], name


def test_reference_roundtrips(a):  # test/deflate_test.dart:12-44 (levels 0 / 1 / 9)
    buf = bytes(i % 256 for i in range(0xfffff))
    for level in (0, 1, 9):
        z = a.Deflate(buf, level=level).get_bytes()
        assert a.Inflate(z + b"\0\0").get_bytes() == buf
        assert zlib.decompress(z, -15) == buf


def test_encoder_framing(a, corpus):  # test/zlib_test.dart:15-55, test/gzip_test.dart:16-52
    d = corpus["text"][:200000]
    z = a.ZLibEncoder().encode_bytes(d)
    assert z == orc.zlib_encode(d)[1] and z[:2] == b"\x78\x01"
    assert a.ZLibDecoder().decode_bytes(z, verify=True) == d == zlib.decompress(z)
    g = a.GZipEncoder().encode_bytes(d, mtime=0)
    assert g == orc.gzip_encode(d, 6, 0)[1]
    assert a.GZipDecoder().decode_bytes(g, verify=True) == d == gzip.decompress(g)
    two = a.ZLibEncoder().encode_bytes(bytes([1, 2, 3])) + a.ZLibEncoder().encode_bytes(bytes([4, 5, 6]))
    assert a.ZLibDecoderWeb().decode_bytes(two, verify=True) == bytes([1, 2, 3, 4, 5, 6])
    two = a.GZipEncoder().encode_bytes(bytes([1, 2, 3])) + a.GZipEncoder().encode_bytes(bytes([4, 5, 6]))
    assert a.GZipDecoderWeb().decode_bytes(two, verify=True) == bytes([1, 2, 3, 4, 5, 6])
    raw = a.ZLibEncoder().encode_bytes(d, raw=True)
    assert raw == orc.deflate(d, 6)[1]


def test_levels_1_to_3_serial_strategy(a, corpus):
    """deflate_fast (levels 1-3) is serial per stream on the device too; byte-identical, just not fast."""
    for name in ("short", "mix", "zeros", "empty", "abc", "tail2"):
        d = corpus[name][:200000]
        for level in (1, 2, 3):
            assert a.Deflate(d, level=level).get_bytes() == orc.deflate(d, level)[1], (name, level)


def test_window_bits(a, corpus):
    """windowBits 9..14 (deflate.dart:107-118): shorter match distances, and blocks whose start has slid out of the
    window can no longer be stored (buf == -1, :677-680) -- incompressible data exercises that rule."""
    for name in ("rand", "mix", "short", "rep"):
        d = corpus[name][:150000]
        for wb in (9, 11, 12, 14):
            for level in (1, 6):
                assert a.Deflate(d, level=level, window_bits=wb).get_bytes() == orc.deflate(d, level, wb)[1], (name, wb, level)


def test_invalid_and_unsupported_parameters(a):
    for kw in (dict(level=10), dict(level=-1), dict(window_bits=8), dict(window_bits=16)):
        with pytest.raises(a.B200ZError) as ei:
            a.Deflate(b"abc", **kw)
        assert ei.value.code == -2
        assert orc.deflate(b"abc", kw.get("level", 6), kw.get("window_bits", 15))[0] == orc.THROW


def test_config3_shape_32MiB(a):
    """BASELINE config 3 at test size: level 6 on 32 MiB of the synthetic text, output identical to the oracle."""
    from archive_b200 import synth
    d = synth.text(32 << 20, stream=30).tobytes()
    z = a.Deflate(d, level=6).get_bytes()
    assert z == orc.deflate(d, 6)[1]
    assert zlib.decompress(z, -15) == d
