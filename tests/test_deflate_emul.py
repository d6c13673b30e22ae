"""CPU-tier coverage of the device Deflate ENCODER (archive_b200/csrc/deflate_kernels.cu: hash chains, _longestMatch table,
speculative lazy-match walk + stitch, block cut, trees, bit emission, and the host driver that sequences them), run on the CUDA
execution-model emulation (tests/host_emul) and compared byte for byte with the oracle's line-by-line Deflate
(deflate.dart:25-100, 997-1118) at levels 1..9 and window sizes 9..15."""
import random
import zlib

import oracle_lib as orc


def same(d, level, wbits=15):
    rc, out, stats = orc.emul_deflate_raw(d, level, wbits)
    ost, oout, _ = orc.deflate(d, level, wbits)
    assert rc == 0 and ost == orc.OK and out == oout, (len(d), level, wbits, rc, len(out), len(oout))
    return out, stats


def test_levels_and_window_bits_on_text():
    from archive_b200 import synth
    d = synth.text(150_000, stream=5).tobytes()
    for level in range(1, 10):
        out, stats = same(d, level)
        assert zlib.decompress(out, -15) == d
        assert stats[1] >= 2  # more than one block
    for wbits in (9, 10, 11, 12, 13, 14):
        for level in (1, 4, 6, 9):
            same(d[:70_000], level, wbits)


def test_edge_sizes_and_degenerate_data():
    rng = random.Random(3)
    cases = [b"", b"a", b"ab", b"abc", b"aaaa", b"a" * 258, b"a" * 259, b"a" * 100000, b"ab" * 40000, bytes(range(256)) * 200,
             bytes(rng.randrange(256) for _ in range(40000)),  # incompressible: stored blocks win
             bytes(rng.randrange(2) for _ in range(60000)),
             b"".join(bytes([rng.randrange(256)]) * rng.choice([1, 2, 3, 4, 257, 258, 259, 260, 600]) for _ in range(800))]
    for d in cases:
        for level in (1, 3, 4, 6, 9):
            same(d, level)


def test_matches_at_the_window_edge():
    """Copies placed exactly at / one past the farthest distance _longestMatch may use (wSize - MIN_LOOKAHEAD,
    deflate.dart:1120-1206) and at TOO_FAR (4096) for length-3 matches."""
    from archive_b200 import synth
    rng = random.Random(9)
    for wbits in (15, 12):
        w = 1 << wbits
        base = bytearray(synth.text(3 * w + 5000, stream=21 + wbits).tobytes())
        for _ in range(40):
            dist = rng.choice([w - 262, w - 261, w - 263, w - 1, w, w + 1, 4096, 4097, 4095, w // 2])
            p = rng.randrange(dist, len(base) - 300)
            n = rng.choice([3, 3, 4, 5, 20, 258, 259])
            base[p:p + n] = base[p - dist:p - dist + n]
        for level in (1, 2, 3, 4, 5, 6, 7, 8, 9):
            same(bytes(base), level, wbits)


def test_fuzz():
    from archive_b200 import synth
    rng = random.Random(0xDEF1)
    n_cases = 0
    for _ in range(60):
        k = rng.randrange(6)
        n = rng.choice([0, 1, 3, 100, 5000, 33000, 66000, rng.randrange(1, 150000)])
        if k == 0:
            d = synth.text(max(n, 1), stream=rng.randrange(1000)).tobytes()[:n]
        elif k == 1:
            d = bytes(rng.randrange(256) for _ in range(min(n, 30000)))
        elif k == 2:
            d = bytes(rng.randrange(rng.choice([2, 3, 4])) for _ in range(n))
        elif k == 3:
            d = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 300))) * (n // 100 + 1)
        elif k == 4:
            d = b"".join(bytes([rng.randrange(256)]) * rng.choice([1, 2, 3, 4, 257, 258, 259, 260, 600]) for _ in range(n // 50 + 1))
        else:
            m = rng.randrange(1, 7)
            d = bytes(((i * m) >> 2) & 0xFF for i in range(n))
        for level in rng.sample(range(1, 10), 3):
            same(d, level, rng.choice([15, 15, 15, 9, 10, 12, 14]))
            n_cases += 1
    assert n_cases == 180


def test_fast_levels_ring_and_rebase():
    """Levels 1-3 (k_defl_fast_batch: a warp per member, head / prev as 16-bit offsets from a base that moves every
    32 KiB, a 64 KiB byte ring refilled 16 KiB at a time): inputs long enough for several re-bases and ring wraps, matches
    at the far edge of every window size, long runs (258-byte matches, interior positions not inserted)."""
    import random
    from archive_b200 import synth
    rng = random.Random(12)
    text = synth.text(230_000, stream=21).tobytes()
    far = bytearray(rng.getrandbits(8) for _ in range(40_000))
    for k in range(0, 200_000, 32_500):  # the same 40 bytes again just inside / outside every window
        far += bytes(far[100:140]) + bytes(rng.getrandbits(8) for _ in range(32_460))
    cases = [text, bytes(far), b"\0" * 150_000, (b"abcdefghij" * 30 + bytes(rng.getrandbits(8) for _ in range(50))) * 400,
             text[:65_536 + 300], text[:131_072 + 2]]
    for d in cases:
        for level in (1, 2, 3):
            same(d, level)
    for wbits in (9, 12, 14):
        same(text[:140_000], 1, wbits)
        same(bytes(far[:150_000]), 3, wbits)


def test_fast_levels_window_turns():
    """Levels 1-3 take 32 positions per turn (k_defl_fast_batch: every lane assumes the window positions before its own were
    entered in the hash chains; a step start whose same-hash predecessor lies inside a long match ends the window).  Data
    made of short repeats, runs and small alphabets has such predecessors all the time."""
    import random
    rng = random.Random(99)
    for it in range(24):
        kind, n = it % 4, rng.choice([300, 1000, 5000, 40000])
        if kind == 0:
            d = bytes(rng.choice(b"ab") for _ in range(n))
        elif kind == 1:
            d = b"".join(bytes([rng.randrange(3)]) * rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 20, 300]) for _ in range(n // 8))
        elif kind == 2:
            d = (b"abcabcabd" * 7 + bytes(rng.getrandbits(8) for _ in range(5))) * (n // 70)
        else:
            w = [bytes(rng.choice(b"etaoin shrdlu") for _ in range(rng.randrange(2, 9))) for _ in range(40)]
            d = b"".join(rng.choice(w) for _ in range(n // 5))
        for level in (1, 2, 3):
            same(d, level, rng.choice([15, 15, 9, 12]))
