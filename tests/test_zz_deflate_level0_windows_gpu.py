"""Deflate(level: 0) with windowBits < 15 on the sm_100a path against the oracle: `_deflateStored` (deflate.dart:691-737) cuts a
stored block as soon as it is 2^windowBits - 262 bytes long ("flush if we may have to slide"), so the block list -- and with
it every byte of the output -- depends on windowBits (a 100 000-byte input: 100 015 bytes at windowBits 15, 100 975 at 9).
A seeded survey of level x windowBits pairs rides along.  Late addition: sorts last, hard time limit."""
import random
import zlib

import pytest

import oracle_lib as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


def test_level0_small_windows(a):
    from archive_b200 import synth
    text = synth.text(150000, stream=60).tobytes()
    for wb in (9, 10, 11, 12, 13, 14, 15):
        for n in (0, 1, 249, 250, 251, 400, 511, 512, 513, 1024, 5000, 32768, 65531, 65536, 100000, 150000):
            d = text[:n]
            st, want, crc = orc.deflate(d, 0, wb)
            dfl = a.Deflate(d, level=0, window_bits=wb)
            assert dfl.get_bytes() == want and dfl.crc32 == crc == zlib.crc32(d), (wb, n, len(want), len(dfl.get_bytes()))
            assert zlib.decompressobj(-15).decompress(want) == d
    assert len(orc.deflate(text[:100000], 0, 9)[1]) == 100975 and len(orc.deflate(text[:100000], 0, 15)[1]) == 100015
    # the framed encoders go the same way (ZLibEncoder(level: 0, windowBits: 9))
    d = text[:40000]
    assert a.ZLibEncoder().encode_bytes(d, level=0, window_bits=9) == orc.zlib_encode(d, 0, 9)[1]


def test_level_and_window_survey(a):
    from archive_b200 import synth
    rng = random.Random(21)
    text = synth.text(120000, stream=59).tobytes()

    def gen(kind, n):
        if kind == 0:
            return text[rng.randrange(1000):][:n]
        if kind == 1:
            return bytes(rng.randrange(256) for _ in range(n))
        if kind == 2:
            return bytes([rng.randrange(4)]) * n
        if kind == 3:
            return (bytes(rng.randrange(256) for _ in range(rng.randrange(1, 300))) * (n // 2 + 1))[:n]
        if kind == 4:
            return b"".join(bytes([rng.randrange(3)]) * rng.choice([1, 2, 3, 258, 259, 600]) for _ in range(n // 50 + 1))[:n]
        return ((text[:rng.randrange(10, 5000)] + bytes(rng.randrange(256) for _ in range(rng.randrange(0, 50)))) * (n // 1000 + 1))[:n]

    for t in range(60):
        n = rng.choice([0, 1, 2, 3, 100, 5000, 32768, 32769, 40000, 65536, 70000, 100000])
        d = gen(t % 6, n)
        level, wb = rng.randrange(0, 10), rng.choice([15, 15, 15, 9, 10, 12])
        assert a.Deflate(d, level=level, window_bits=wb).get_bytes() == orc.deflate(d, level, wb)[1], (t, t % 6, n, level, wb)
