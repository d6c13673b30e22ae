"""Parity tests for the sm_100a BZip2 decode path (K6-K8), through the C ABI, against the oracle and libbz2.
Fixtures: test/bzip2_test.dart:8-12 (test.bz2), io_test's test2.tar.bz2; seeded synthetic data incl. multi-block
streams (no reference fixture has one), degenerate run patterns for the RLE1 automaton, and bad data."""
import bz2
import hashlib
import json
import os
import random

import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))


def rd(n):
    return open(os.path.join(G, n), "rb").read()


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


def test_fixtures(a):
    for name in ("test.bz2", "test2.tar.bz2"):
        out = a.BZip2Decoder().decode_bytes(rd(name), verify=True)
        assert hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name
        assert out == orc.bzip2_decode(rd(name))[1]
    assert a.BZip2Decoder().decode_bytes(rd("test2.tar.bz2")) == rd("test2.tar")


def test_roundtrip_cat_jpg(a):  # test/bzip2_test.dart:14-25 (decode side; libbz2 encodes)
    z = bz2.compress(rd("cat.jpg"), 9)
    assert a.BZip2Decoder().decode_bytes(z, verify=True) == rd("cat.jpg")


def test_synthetic_vs_oracle(a):
    from archive_b200 import synth
    rng = random.Random(31)
    big = synth.text(2_300_000, stream=7).tobytes()
    cases = [big, b"", b"a", b"ab" * 5, b"aaaa", b"aaaaa", b"aaaa" * 100000, bytes([251]) * 70000,
             bytes([4]) * 1000 + bytes([5]) * 9 + bytes([4]) * 5, bytes(rng.getrandbits(8) for _ in range(200000)),
             big[:1000], bytes(rng.choice(b"ab") for _ in range(50000)),
             b"".join(bytes([rng.randrange(4)]) * rng.choice([1, 3, 4, 5, 8, 9, 10, 255, 256, 259, 260]) for _ in range(3000))]
    for d in cases:
        for level in (1, 9):
            z = bz2.compress(d, level)
            ost, oout = orc.bzip2_decode(z, verify=True)
            out = a.BZip2Decoder().decode_bytes(z, verify=True)
            assert ost == orc.OK and out == oout == d, (len(d), level)


def test_stops_at_first_eos_and_bad_data(a):
    from archive_b200 import synth
    d = synth.text(300000, stream=8).tobytes()
    z = bz2.compress(d, 1)  # 4 blocks
    out = a.OutputMemoryStream()
    assert a.BZip2Decoder().decode_stream(a.InputMemoryStream(z + z), out) is True
    assert out.get_bytes() == d  # the second stream is not decoded (bzip2_decoder.dart:83-84)
    # bad signature -> false, nothing written
    out = a.OutputMemoryStream()
    assert a.BZip2Decoder().decode_stream(a.InputMemoryStream(b"BZx1" + z[4:]), out) is False
    assert out.get_bytes() == b""
    # a flipped bit inside a block: whatever the reference makes of it (here: the block still decodes, to garbage, and
    # fails its CRC -- its bytes are already written when the CRC is compared, bzip2_decoder.dart:58-66), the GPU path
    # must produce the same bytes and the same verdict
    bad = bytearray(z)
    bad[len(z) * 5 // 8] ^= 0x10
    ost, oout = orc.bzip2_decode(bytes(bad), verify=True)
    out = a.OutputMemoryStream()
    try:
        ok = a.BZip2Decoder().decode_stream(a.InputMemoryStream(bytes(bad)), out, verify=True)
        assert ok is False and ost == orc.FALSE
        assert out.get_bytes() == oout
    except a.DartRangeError:
        assert ost == orc.THROW
    assert out.get_bytes()[:90000] == d[:90000]  # the block before the damage is intact
    # truncated stream: readByte past the end throws in the reference
    with pytest.raises(a.DartRangeError):
        a.BZip2Decoder().decode_bytes(z[:len(z) // 2])
    assert orc.bzip2_decode(z[:len(z) // 2])[0] == orc.THROW


def test_scale_multiblock(a):
    """Config-4 shape at test size: one BZh9 stream of many 900 KB blocks; every block CRC and the combined CRC
    are verified on the device path (verify=True) and the bytes equal the source."""
    from archive_b200 import synth
    d = synth.text(24 << 20, stream=12).tobytes()
    z = bz2.compress(d, 9)
    out = a.BZip2Decoder().decode_bytes(z, verify=True)
    assert len(out) == len(d) and out == d


def test_decode_paths_agree(a, monkeypatch):
    """The ways K7 / K8 can be run give the same bytes and verdicts: the fast entropy kernel (default) or the exact kernel alone
    (B200Z_BZ2_FAST=0); the RLE1 output pass in one piece or in groups with early copies (B200Z_BZ2_EMIT_GROUPS); the whole
    of K8 in groups (B200Z_BZ2_GROUPS).  Long runs (records beside the bytes the fast kernel writes itself), a damaged
    block in the middle and a stream that ends early are in the set."""
    from archive_b200 import synth
    text = synth.text(3_000_000, stream=14).tobytes()
    runs = (b"\0" * 200_000 + text[:150_000] + b"ab" * 90_000 + bytes([9]) * 70_000) * 3
    z_text, z_runs = bz2.compress(text, 9), bz2.compress(runs, 3)
    bad = bytearray(bz2.compress(text, 2))
    bad[len(bad) * 5 // 8] ^= 0x10
    streams = [z_text, z_runs, bytes(bad), z_text[:len(z_text) * 2 // 3]]

    def run(z):
        out = a.OutputMemoryStream()
        try:
            ok = a.BZip2Decoder().decode_stream(a.InputMemoryStream(z), out, verify=True)
        except a.DartRangeError:
            ok = "throw"
        return ok, out.get_bytes()

    want = [run(z) for z in streams]
    assert want[0] == (True, text) and want[1] == (True, runs)
    for env in ({"B200Z_BZ2_FAST": "0"}, {"B200Z_BZ2_EMIT_GROUPS": "1"}, {"B200Z_BZ2_EMIT_GROUPS": "3"}, {"B200Z_BZ2_GROUPS": "2"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        assert [run(z) for z in streams] == want, env
        for k in env:
            monkeypatch.delenv(k)


def test_randomised_blocks(a):
    """The obsolete randomised-block bit (no encoder has set it since bzip2 0.9.5): the reference decodes such blocks with its
    own variant of the de-randomisation (bzip2_decoder.dart:492-608, SURVEY Q6 -- it differs from libbzip2, so libbz2 is no
    witness here; the oracle restates it).  Setting the bit on a normal stream changes the decoded bytes (CRCs then mismatch,
    which only `verify` looks at)."""
    from archive_b200 import synth
    rng = random.Random(8)
    cases = [b"hello hello hello, randomised world! " * 40,
             bytes(rng.randrange(256) for _ in range(5000)),
             b"a" * 3000 + b"bcd" * 500 + bytes(range(256)) * 4,
             synth.text(1_200_000, stream=970).tobytes()]  # two blocks: the first one randomised
    for src in cases:
        z = bytearray(bz2.compress(src, 9))
        z[14] |= 0x80  # bit 112 = 32 (stream header) + 48 (block magic) + 32 (block CRC): the "randomised" flag
        z = bytes(z)
        st, want = orc.bzip2_decode(z, verify=False)
        got = a.BZip2Decoder().decode_bytes(z, verify=False)
        assert got == want and want != src
        ok = a.BZip2Decoder().decode_stream(a.InputMemoryStream(z), a.OutputMemoryStream(), verify=True)
        assert ok is False  # block CRC no longer matches
