"""Parity tests proper: the sm_100a inflate path, called through the C ABI (ctypes -> libb200z.so), against
the oracle on the same bytes -- golden fixtures, seeded synthetic units, the edge cases the reference tests
(empty / ragged inputs, stored / fixed / dynamic blocks, multi-member, multi-stream, bad data), and
size-independent properties at scale.  Bit-exact: byte/integer work has no tolerance."""
import ctypes as C
import gzip
import hashlib
import json
import os
import random
import zlib

import numpy as np
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


def inflate_batch(units, caps):
    """units: list[bytes] of raw DEFLATE; -> list[(status, bytes, in_used)] via b200z_inflate_batch."""
    from archive_b200 import _ffi
    L = _ffi.ensure_init()
    n = len(units)
    in_off = np.zeros(n, dtype=np.uint64)
    in_len = np.array([len(u) for u in units], dtype=np.uint32)
    pos = 0
    rng = random.Random(5)
    blob = bytearray()
    for i, u in enumerate(units):
        blob += bytes(rng.randrange(4))  # ragged alignment between units
        in_off[i] = len(blob)
        blob += u
    blob = bytes(blob)
    out_cap = np.array(caps, dtype=np.uint32)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(out_cap.astype(np.uint64))[:-1]
    out_bytes = int(out_cap.astype(np.uint64).sum())
    out = np.zeros(max(out_bytes, 1), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    used = np.zeros(n, dtype=np.uint32)
    addr, nb, keep = _ffi.as_buffer(blob)
    p = lambda arr: arr.ctypes.data
    rc = L.b200z_inflate_batch(addr, nb, p(in_off), p(in_len), p(out), out_bytes, p(out_off), p(out_cap), p(out_len),
                               p(status), p(used), n)
    assert rc == 0, _ffi.last_error()
    res = []
    for i in range(n):
        o = int(out_off[i])
        res.append((int(status[i]), out[o:o + int(out_len[i])].tobytes(), int(used[i])))
    return res


# ------------------------------------------------------------------ reference fixtures through the class API
def test_inflate_data_bin(a):  # test/inflate_test.dart:14-20
    out = a.Inflate(rd("inflate_data.bin")).get_bytes()
    assert len(out.decode("utf8")) == 5259
    assert hashlib.sha256(out).hexdigest() == MAN["inflate_data.bin"]["sha256"]
    assert out == orc.inflate(rd("inflate_data.bin"))[1]


def test_gzip_fixtures(a):  # test/gzip_test.dart:63-93
    for name in ("cat.jpg.gz", "test2.tar.gz", "a.txt.gz"):
        out = a.GZipDecoder().decode_bytes(rd(name))
        assert hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name
        assert out == orc.gzip_decode(rd(name))[1]
    assert a.GZipDecoderWeb().decode_bytes(rd("cat.jpg.gz")) == rd("cat.jpg")


def test_git_vector(a):  # test/inflate_test.dart:57-179: first zlib stream, 148 bytes consumed
    data = rd("git_inflate_input.bin")
    inp = a.InputMemoryStream(data[2:])
    inf = a.Inflate.stream(inp)
    assert inf.get_bytes() == rd("git_expected_output.bin")
    assert inp.position + 2 + 4 == 148


def test_zlib_multistream_verify(a):  # test/zlib_test.dart:15-23
    data = zlib.compress(bytes([1, 2, 3])) + zlib.compress(bytes([4, 5, 6]))
    assert a.ZLibDecoderWeb().decode_bytes(data, verify=True) == bytes([1, 2, 3, 4, 5, 6])
    assert a.ZLibDecoderWeb().decode_bytes(data, verify=True) == orc.zlib_decode(data, verify=True)[1]
    # a wrong Adler drops the failing stream only (_zlib_decoder_web.dart:88-96)
    bad = bytearray(data)
    bad[-1] ^= 0xff
    out = a.OutputMemoryStream()
    ok = a.ZLibDecoderWeb().decode_stream(a.InputMemoryStream(bytes(bad)), out, verify=True)
    st, oout = orc.zlib_decode(bytes(bad), verify=True)
    assert ok is False and st == orc.FALSE
    assert out.get_bytes() == oout == bytes([1, 2, 3])


def test_gzip_multimember(a):  # test/gzip_test.dart:44-52
    data = gzip.compress(bytes([1, 2, 3])) + gzip.compress(bytes([4, 5, 6]))
    assert a.GZipDecoderWeb().decode_bytes(data, verify=True) == bytes([1, 2, 3, 4, 5, 6])


def test_roundtrip_levels(a):  # test/deflate_test.dart:12-44 / inflate_test.dart:22-54 (decode side)
    buf = bytes(i % 256 for i in range(0xfffff))
    for level in (0, 1, 9):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        z = co.compress(buf) + co.flush()
        inp = a.InputMemoryStream(z)
        assert a.Inflate.stream(inp).get_bytes() == buf
        assert inp.position == len(z)
    z = zlib.compress(bytes(i % 256 for i in range(10000)))
    assert a.ZLibDecoder().decode_bytes(z, verify=True) == bytes(i % 256 for i in range(10000))


# ------------------------------------------------------------------ batch kernel vs oracle
def corpus(rng, n):
    words = [bytes(rng.choice(b"etaoinshrdlu") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b" "
    return bytes(b[:n])


def test_batch_mixed_units_vs_oracle():
    rng = random.Random(17)
    units, expect = [], []
    for it in range(300):
        t = corpus(rng, rng.choice([0, 1, 2, 3, 100, 5000, 65536, 70000]))
        if it % 7 == 0:
            t = bytes(rng.getrandbits(8) for _ in range(len(t) // 4))
        if it % 11 == 0:
            t = b"\0" * len(t)
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, rng.choice([1, 8, 9]),
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE]))
        h = len(t) // 2
        z = co.compress(t[:h]) + co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH])) + \
            co.compress(t[h:]) + co.flush()
        z += rng.choice([b"", b"\0\0", b"trailing garbage"])
        units.append(z)
        expect.append(orc.inflate(z))
    res = inflate_batch(units, [len(e[1]) + rng.choice([0, 0, 5]) for e in expect])
    for i, ((st, out, used), (ost, oout, oused)) in enumerate(zip(res, expect)):
        assert ost == orc.OK
        assert out == oout, i
        assert st in (0, 1, -1), (i, st)
        if st == 0:
            assert used == oused, i


def test_batch_bad_data_vs_oracle():
    rng = random.Random(23)
    units = []
    base = []
    for it in range(12):
        t = corpus(rng, rng.randint(1, 4000))
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, 8, rng.choice([0, 4]))
        base.append(co.compress(t) + co.flush())
    for z in base:
        for cut in range(0, len(z), max(1, len(z) // 25)):
            units.append(z[:cut])
        for _ in range(25):
            zz = bytearray(z)
            zz[rng.randrange(len(zz))] ^= 1 << rng.randrange(8)
            units.append(bytes(zz))
    for _ in range(100):
        units.append(bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 300))))
    units = [u for u in units if len(u) > 0]
    res = inflate_batch(units, [1 << 16] * len(units))
    seen = set()
    for i, (st, out, used) in enumerate(res):
        ost, oout, oused = orc.inflate(units[i])
        seen.add((st, ost))
        if st == -2:  # cap reached: prefix
            assert oout[:len(out)] == out
        elif ost == orc.OK:
            if st in (0, 1, -1):
                assert out == oout, i
            else:
                assert st in (-3, -4) and oout[:len(out)] == out, (i, st)
        elif ost == orc.RUNAWAY:
            assert st == -4, (i, st)
        else:
            assert st in (-3, -4, -5), (i, st)
    assert (0, 0) in seen and (-1, 0) in seen


def test_empty_inputs(a):
    assert a.Inflate(b"").get_bytes() == b""
    assert a.GZipDecoder().decode_bytes(b"") == b""
    assert a.ZLibDecoder().decode_bytes(b"") == b""
    assert a.inflate_buffer(b"\x03\x00") == b""


def test_synthetic_units_config1_and_config2_shapes():
    """BASELINE configs 1 and 2 at test size: 64 KiB units of the section-8(d) text, one fixed-Huffman block
    (config 1) and one dynamic block per gzip member (config 2), each compared byte-for-byte with the oracle."""
    from archive_b200 import synth
    text = synth.text(64 * 65536, stream=3)
    chunks = [text[i:i + 65536].tobytes() for i in range(0, len(text), 65536)]
    fixed = [synth.deflate_raw(c, 6, 9, zlib.Z_FIXED) for c in chunks[:8]]
    assert fixed[0][0] & 7 == 0b011
    dyn = [synth.deflate_raw(c, 6, 9) for c in chunks]
    assert dyn[0][0] & 7 == 0b101
    units = fixed + dyn
    units = [u + b"\0\0" for u in units]  # >= 2 pad bytes: the reference's short-read quirk Q1 cannot fire
    res = inflate_batch(units, [65536] * len(units))
    for i, (st, out, used) in enumerate(res):
        ost, oout, oused = orc.inflate(units[i])
        assert st == 0 and out == oout and used == len(units[i]) - 2 == oused, i
        assert out == (chunks[i] if i < 8 else chunks[i - 8])
    # and WITHOUT padding the quirk must be reproduced bit for bit: the EOB (and whatever follows a short
    # read) is lost exactly as in inflate.dart:192-195
    raw_units = [u[:-2] for u in units]
    res = inflate_batch(raw_units, [65536] * len(raw_units))
    for i, (st, out, used) in enumerate(res):
        ost, oout, oused = orc.inflate(raw_units[i])
        assert out == oout and st in (0, -1), i


def test_gzip_members_hinted_and_unhinted(a):
    from archive_b200 import synth
    text = synth.text(48 * 65536 + 12345, stream=4)
    hinted = b"".join(synth.gzip_members(text, workers=1))
    plain = b"".join(synth.gzip_members(text, hint=False, workers=1))
    for blob in (hinted, plain):
        out = a.GZipDecoderWeb().decode_bytes(blob)
        assert out == text.tobytes()
    # a lying BSIZE hint must not change the result (the library re-checks every hint)
    liar = bytearray(hinted)
    liar[16] ^= 0x40
    assert a.GZipDecoderWeb().decode_bytes(bytes(liar)) == text.tobytes()
    st, oout = orc.gzip_decode(hinted[:400000])
    assert st in (orc.OK, orc.THROW)


@pytest.mark.parametrize("n_units", [4096])
def test_scale_property_crc_of_members(a, n_units):
    """Full-size property (size-independent): every decoded 64 KiB unit must hash to the CRC-32 its own gzip
    trailer carries, and the whole output to the text it was made from."""
    from archive_b200 import synth
    text = synth.text(n_units * 65536, stream=9)
    members = synth.gzip_members(text)
    out = a.GZipDecoder().decode_bytes(b"".join(members))
    assert len(out) == len(text)
    assert out == text.tobytes()
    for i in (0, 1, n_units // 2, n_units - 1):
        m = members[i]
        crc = int.from_bytes(m[-8:-4], "little")
        assert zlib.crc32(out[i * 65536:(i + 1) * 65536]) == crc
