"""CPU-tier coverage of the device BZip2 DECODE kernels (K6 magic scan, K7 entropy decode, K8 inverse BWT / RLE / CRC of
archive_b200/csrc/bzip2_kernels.cu): the kernels run on the CUDA execution-model emulation (tests/host_emul) and must give
the oracle's bytes and verdict -- fixtures of test/bzip2_test.dart:8-12 and io_test, seeded synthetic data, the RLE1
automaton's degenerate runs, randomised blocks, damaged and truncated streams."""
import bz2
import hashlib
import json
import os
import random

import oracle_lib as orc

G = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))


def rd(n):
    return open(os.path.join(G, n), "rb").read()


def same(z, verify=True):
    ost, oout = orc.bzip2_decode(z, verify=verify)
    est, eout, _ = orc.emul_bzip2_decode(z, verify=verify)
    assert (est, eout) == (ost, oout), (len(z), ost, est, len(oout), len(eout))
    return est, eout


def test_fixtures():
    for name in ("test.bz2", "test2.tar.bz2"):
        st, out = same(rd(name))
        assert st == orc.OK and hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name
    assert same(rd("test2.tar.bz2"))[1] == rd("test2.tar")
    z = bz2.compress(rd("cat.jpg"), 9)  # test/bzip2_test.dart:14-25, decode side
    assert same(z) == (orc.OK, rd("cat.jpg"))


def test_synthetic_vs_oracle():
    from archive_b200 import synth
    rng = random.Random(31)
    big = synth.text(450_000, stream=7).tobytes()
    cases = [big, b"", b"a", b"ab" * 5, b"aaaa", b"aaaaa", b"aaaa" * 100000, bytes([251]) * 70000,
             bytes([4]) * 1000 + bytes([5]) * 9 + bytes([4]) * 5, bytes(rng.getrandbits(8) for _ in range(120000)),
             big[:1000], bytes(rng.choice(b"ab") for _ in range(50000)), b"abcdefgh" * 30000, bytes(range(256)) * 300,
             b"".join(bytes([rng.randrange(4)]) * rng.choice([1, 3, 4, 5, 8, 9, 10, 255, 256, 259, 260]) for _ in range(3000))]
    for d in cases:
        for level in (1, 9):
            z = bz2.compress(d, level)
            assert same(z) == (orc.OK, d), (len(d), level)


def test_stops_at_first_eos_and_bad_data():
    from archive_b200 import synth
    d = synth.text(300000, stream=8).tobytes()
    z = bz2.compress(d, 1)  # 4 blocks
    assert same(z + z) == (orc.OK, d)  # the second stream is not decoded (bzip2_decoder.dart:83-84)
    assert same(b"BZx1" + z[4:]) == (orc.FALSE, b"")
    assert same(b"BZh:" + z[4:]) == (orc.FALSE, b"")
    for n in (0, 1, 2, 3):
        assert same(z[:n])[0] == orc.THROW
    assert same(z[:4]) == (orc.OK, b"")
    bad = bytearray(z)
    bad[len(z) * 5 // 8] ^= 0x10
    st, out = same(bytes(bad))
    assert st != orc.OK and out[:90000] == d[:90000]  # the blocks before the damage are intact
    same(bytes(bad), verify=False)
    assert same(z[:len(z) // 2])[0] == orc.THROW
    for cut in range(1, 24):  # the end-of-stream magic / combined CRC cut short
        same(z[:-cut])
    bad = bytearray(z)
    bad[-2] ^= 1  # combined CRC
    assert same(bytes(bad))[0] == orc.FALSE
    assert same(bytes(bad), verify=False) == (orc.OK, d)


def test_randomised_blocks():
    from archive_b200 import synth
    rng = random.Random(8)
    cases = [b"hello hello hello, randomised world! " * 40, bytes(rng.randrange(256) for _ in range(5000)),
             b"a" * 3000 + b"bcd" * 500 + bytes(range(256)) * 4, synth.text(250_000, stream=970).tobytes()]
    for src in cases:
        for level in (1, 9):
            z = bytearray(bz2.compress(src, level))
            z[14] |= 0x80  # the first block's "randomised" flag
            st, out = same(bytes(z), verify=False)
            assert out != src
            assert same(bytes(z), verify=True)[0] == orc.FALSE


def test_run_of_four_at_block_end():
    """Two damaged streams (found by the fuzz below, kept as fixtures) whose first block ends on 4 equal bytes with no count
    byte behind them: the reference reads the count one step past the block, writes the run and only then returns -1
    (bzip2_decoder.dart:708-716, 628-631) -- more bytes than the block holds, and decodeStream is false."""
    for name in ("bz2_run_at_block_end_a.bz2", "bz2_run_at_block_end_b.bz2"):
        for verify in (False, True):
            st, out = same(rd(name), verify=verify)
            assert st == orc.FALSE and hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name


DAMAGED = ("bz2_mtfval_quirk_a.bz2", "bz2_mtfval_quirk_b.bz2", "bz2_mtfval_quirk_c.bz2", "bz2_short_cycle.bz2",
           "bz2_rand_overrun_a.bz2", "bz2_rand_overrun_b.bz2", "bz2_run_at_block_end_a.bz2", "bz2_run_at_block_end_b.bz2")


def test_damaged_fixtures():
    """Small damaged streams kept as fixtures (tests/golden/manifest.json says what each one is): the reference's verdict and
    bytes, from the oracle, for the paths damaged data takes -- the literal entropy kernel (_getMtfVal's unchecked -1), a
    short inverse-BWT cycle, overrunning run-length walks."""
    for name in DAMAGED:
        z = rd(name)
        st, out = same(z, verify=False)
        assert st == MAN[name]["status"] and hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name
        same(z, verify=True)
        if "mtfval" in name:
            assert st == orc.OK and orc.emul_bzip2_last_quirk() >= 0
    same(rd("bz2_mtfval_quirk_a.bz2"), verify=False)
    assert orc.emul_bzip2_last_quirk() == 1  # it did take the literal path


def test_stream_ends_inside_a_block_signature():
    """_readBlockType (bzip2_decoder.dart:90-111) compares byte by byte: a wrong byte ends the stream with `false` before the
    missing ones throw."""
    from archive_b200 import synth
    d = synth.text(20000, stream=3).tobytes()
    z = bz2.compress(d, 1)
    verdicts = set()
    for cut in range(5, 11):
        for back in range(1, 4):
            for bit in range(8):
                t = bytearray(z[:-cut])
                t[-back] ^= 1 << bit
                verdicts.add(same(bytes(t), verify=False)[0])
    assert verdicts >= {orc.FALSE, orc.THROW}


def test_fuzz_damage():
    """Seeded bit flips, byte overwrites, truncations and the randomised flag on small streams of several shapes: whatever the
    reference makes of the damage, the kernels agree -- garbage that fails its CRC; a block the reference KEEPS decoding after
    a bad Huffman code (_getMtfVal's -1 is only checked on its first call, bzip2_decoder.dart:273-275, 387: the literal
    kernel); a run of 4 whose count lies past the block; a stream that ends inside a block signature; a read past the end."""
    from archive_b200 import synth
    rng = random.Random(0xB2)

    def mk():
        k = rng.randrange(5)
        if k == 0:
            return synth.text(rng.randrange(1000, 100000), stream=rng.randrange(1000)).tobytes()
        if k == 1:
            return bytes(rng.randrange(rng.choice([2, 3, 7, 256])) for _ in range(rng.randrange(1, 40000)))
        if k == 2:
            return b"".join(bytes([rng.randrange(3)]) * rng.choice([1, 2, 4, 5, 255, 256, 1000]) for _ in range(rng.randrange(1, 800)))
        if k == 3:
            return bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40))) * rng.randrange(1, 3000)
        return b"".join(bytes([rng.randrange(256)]) * rng.choice([3, 4, 4, 4, 5, 259, 260]) for _ in range(rng.randrange(1, 500)))

    seen = {"literal": 0, "false_with_bytes": 0, "throw": 0, "ok": 0}
    for _ in range(60):
        z = bz2.compress(mk(), rng.choice([1, 1, 1, 2, 9]))
        for _ in range(10):
            bad = bytearray(z)
            kind = rng.randrange(4)
            if kind == 0:
                for _k in range(rng.choice([1, 1, 2, 5])):
                    bad[rng.randrange(4, len(bad))] ^= 1 << rng.randrange(8)
            elif kind == 1:
                p = rng.randrange(4, len(bad))
                bad[p:p + rng.randrange(1, 9)] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
            elif kind == 2:
                bad = bad[:rng.randrange(0, len(bad))]
            else:
                if len(bad) > 14:
                    bad[14] |= 0x80
                if rng.random() < 0.5 and len(bad) > 30:
                    bad[rng.randrange(15, len(bad))] ^= 1 << rng.randrange(8)
            same(bytes(bad), verify=True)
            st, out = same(bytes(bad), verify=False)
            lit = orc.emul_bzip2_last_quirk()
            seen["literal"] += lit > 0
            seen["false_with_bytes"] += st == orc.FALSE and len(out) > 0
            seen["throw"] += st == orc.THROW
            seen["ok"] += st == orc.OK
    assert all(seen.values()), seen  # every path of interest was exercised


def test_fast_kernel_takes_clean_blocks(monkeypatch):
    """k_bz2_entropy_fast (two-warp pipeline: parallel bit walk + symbolic move-to-front) must finish every clean block --
    and the exact kernel alone (B200Z_BZ2_FAST=0) must still give the same bytes."""
    from archive_b200 import synth
    rng = random.Random(5)
    d1 = synth.text(260_000, stream=11).tobytes()
    cases = [(d1, 1, 3), (bytes(rng.getrandbits(8) for _ in range(150000)), 1, 2), (b"ab" * 60000, 9, 1),
             (bytes(range(256)) * 500, 9, 1), (bytes([7]) * 300000, 9, 1), (b"x", 9, 1),
             (b"".join(bytes([rng.randrange(6)]) * rng.choice([1, 2, 3, 4, 5, 7, 50, 255, 300, 70000]) for _ in range(400)), 9, None)]
    for d, level, nblocks in cases:
        z = bz2.compress(d, level)
        assert same(z) == (orc.OK, d)
        nfast = orc.emul_bzip2_last_fast()
        if nblocks is not None:
            assert nfast == nblocks, (len(d), level, nfast)
        else:
            assert nfast >= 1
    monkeypatch.setenv("B200Z_BZ2_FAST", "0")
    assert same(bz2.compress(d1, 1)) == (orc.OK, d1)
    assert orc.emul_bzip2_last_fast() == 0


def test_fast_kernel_long_codes_and_deep_lists():
    """Codes longer than the 10-bit look-up table (the walker's and the decoder's limit / base walk) and move-to-front
    positions deep in the list (the symbolic lists' lazy initialisation, up to all 64 words): 250 distinct bytes, most of
    the text a few frequent ones, the rest rare -- the rare list positions get codes of 11 .. 15 bits (694 of the first block's 6 x 188 codes are longer than 10)."""
    rng = random.Random(77)
    common = bytes(rng.choice(b"etaoinshr ") for _ in range(64))
    out = bytearray()
    while len(out) < 260_000:
        out += bytes(rng.choice(common) for _ in range(rng.randrange(20, 400)))
        if rng.random() < 0.7:
            out += bytes([rng.randrange(250)]) * rng.choice([1, 1, 2, 5])
    d = bytes(out)
    for level, nblocks in ((1, 3), (3, 1)):
        z = bz2.compress(d, level)
        assert same(z) == (orc.OK, d)
        assert orc.emul_bzip2_last_fast() == nblocks  # all of them by the fast kernel
    # every byte value in turn, again and again: list positions 255 all the time
    d2 = bytes(range(256)) * 700 + bytes(reversed(range(256))) * 300
    z2 = bz2.compress(d2, 9)
    assert same(z2) == (orc.OK, d2) and orc.emul_bzip2_last_fast() == 1
