"""The CUDA kernel's per-stream decode logic (archive_b200/csrc/inflate_decode.cuh), compiled for the host
by tests/host_emul, against the oracle: fixtures, every block type, truncations, bit flips, random bytes.
This is a LOGIC check that runs without a GPU; the parity tests proper are the -m gpu tests."""
import os
import random
import zlib

import oracle_lib as orc

G = os.path.join(os.path.dirname(__file__), "golden")
DONE, EOS, STOP, NOSPC, RANGE, BADCODE, UTHROW = 0, 1, -1, -2, -3, -4, -5


def check(z: bytes, tag=""):
    ust, out, used, ntok = orc.emul_inflate(z, 1 << 18)
    ost, oout, oused = orc.inflate(z)
    if ost == orc.OK:
        if ust in (DONE, EOS, STOP):
            assert out == oout, tag
            if ust == DONE:
                assert used == oused, tag
        else:  # BADCODE (documented divergence) / RANGE on a truncated extra-bits read: prefix of the reference
            assert ust in (BADCODE, RANGE), (tag, ust)
            assert oout[:len(out)] == out, tag
    elif ost == orc.RUNAWAY:
        assert ust == BADCODE, (tag, ust)
    else:
        assert ust in (RANGE, UTHROW, BADCODE), (tag, ust)
        if ust != BADCODE:
            assert oout[:len(out)] == out, tag
    return ust


def corpus(rng, n):
    words = [bytes(rng.choice(b"etaoinshrdlu") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b" "
    return bytes(b[:n])


def test_fixture_and_git_vector():
    assert check(open(os.path.join(G, "inflate_data.bin"), "rb").read()) == DONE
    assert check(open(os.path.join(G, "git_inflate_input.bin"), "rb").read()[2:]) == DONE
    cat = open(os.path.join(G, "cat.jpg.gz"), "rb").read()
    assert cat[3] == 0x08  # FNAME only
    assert check(cat[cat.index(b"\0", 10) + 1:]) == DONE


def test_block_types_levels_and_flushes():
    rng = random.Random(3)
    for it in range(40):
        t = corpus(rng, rng.randint(1, 70000))
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, rng.choice([1, 8, 9]),
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE]))
        h = len(t) // 2
        z = co.compress(t[:h]) + co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH])) + \
            co.compress(t[h:]) + co.flush()
        for pad in (b"", b"\0", b"\0\0", b"12345678"):
            check(z + pad, f"it{it} pad{len(pad)}")


def test_empty_and_tiny():
    for z in (b"", b"\x03", b"\x03\x00", b"\x01", b"\x01\x00\x00\xff\xff", b"\x00\x00\x00\xff\xff\x03\x00",
              b"\x01\x01\x00\xfe\xff\x41", b"\x01\x02\x00\xfd\xffAB", b"\x01\x05\x00\xfa\xffABCDE",
              b"\x07", b"\x05", b"\xff" * 8):
        check(z, repr(z))


def test_truncations_and_bitflips():
    rng = random.Random(7)
    for it in range(25):
        t = corpus(rng, rng.randint(1, 5000))
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, 8, rng.choice([0, 4, 2]))
        z = co.compress(t) + co.flush()
        for cut in range(0, len(z), max(1, len(z) // 40)):
            check(z[:cut], f"trunc {it}:{cut}")
        for _ in range(25):
            zz = bytearray(z)
            zz[rng.randrange(len(zz))] ^= 1 << rng.randrange(8)
            check(bytes(zz), f"flip {it}")
        for _ in range(8):
            check(bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 200))), f"rand {it}")


def test_max_lengths_and_distances():
    # long runs (len 258, dist 1), far matches (dist ~32K), 15-bit codes from a skewed alphabet
    rng = random.Random(11)
    far = bytes(rng.getrandbits(8) for _ in range(300))
    t = b"\0" * 70000 + far + bytes(rng.getrandbits(8) for _ in range(32000)) + far + b"x" * 1000
    skew = bytearray()
    for s in range(40):
        skew += bytes([s]) * (1 << max(0, 16 - s))
    rng.shuffle(skew)
    for data in (t, bytes(skew)):
        for lvl in (1, 9):
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9)
            z = co.compress(data) + co.flush()
            assert check(z + b"\0\0") == DONE
