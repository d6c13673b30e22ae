"""k_inflate_decode / k_inflate_expand (archive_b200/csrc/inflate_kernels.cu) executed on the CUDA execution-model
emulation (tests/host_emul/cuda_emu.h): several lanes per stream.  The speculative helper lanes, the piece stitching and
the expand kernel's range checks must give exactly what the one-lane-per-stream decode gives (which
tests/test_decode_logic_emul.py pins against the oracle): bytes, out_len, status and in_used -- on valid streams of every
block type, on multi-block streams, and on truncated / corrupted ones."""
import ctypes as C
import os
import random
import zlib

import numpy as np
import pytest

import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_E = None


def run(units, caps, upw, lpu):
    global _E
    if _E is None:
        _E = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libinflate_emul.so"))
    n = len(units)
    blob = bytearray(64)
    in_off, in_len = [], []
    for u in units:
        in_off.append(len(blob))
        in_len.append(len(u))
        blob += u
        blob += b"\xA5" * (3 + (len(blob) * 7) % 13)  # neighbours are never zero padding, offsets are unaligned
    blob += bytes(64)
    out_off, o = [], 0
    for c in caps:
        out_off.append(o)
        o += c
    inb = (C.c_uint8 * len(blob)).from_buffer(blob)
    outb = (C.c_uint8 * (o + 64))()
    arr = lambda t, l: (t * len(l))(*l)
    ol, st, iu, pc = (C.c_uint32 * n)(), (C.c_int32 * n)(), (C.c_uint32 * n)(), (C.c_uint32 * n)()
    _E.emu_inflate_batch(inb, arr(C.c_uint64, in_off), arr(C.c_uint32, in_len), outb, arr(C.c_uint64, out_off),
                         arr(C.c_uint32, caps), ol, st, iu, n, C.c_size_t(o), upw, lpu, pc)
    outs = [bytes(outb[out_off[i]:out_off[i] + min(ol[i], caps[i])]) for i in range(n)]
    return outs, list(ol), list(st), list(iu), list(pc)


def same_as_one_lane(units, caps, tag):
    ref = run(units, caps, 8, 1)
    for upw, lpu in ((8, 4), (4, 8), (16, 2)):
        got = run(units, caps, upw, lpu)
        for k, name in enumerate(("bytes", "out_len", "status")):
            assert got[k] == ref[k], (tag, upw, lpu, name)
        # in_used is what the framing layer continues from; it has no meaning after "output full" / a thrown RangeError
        keep = [i for i, st in enumerate(ref[2]) if st not in (-2, -3)]
        assert [got[3][i] for i in keep] == [ref[3][i] for i in keep], (tag, upw, lpu, "in_used")
    return ref


def text(rng, n):
    words = [bytes(rng.choice(b"etaoinshrdlucmfwypvbgkqjxz") for _ in range(rng.randint(2, 10))) for _ in range(3000)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + rng.choice([b" ", b" ", b" ", b", ", b".\n"])
    return bytes(b[:n])


def test_valid_members_are_stitched_from_helper_pieces():
    from archive_b200 import synth
    t = synth.text(12 * 65536)
    plain = [t[i * 65536:(i + 1) * 65536].tobytes() for i in range(12)]
    units = [synth.deflate_raw(p) + bytes(8) for p in plain]  # + gzip trailer room, as the framing layer passes it
    ref = same_as_one_lane(units, [65536] * 12, "cfg2")
    assert ref[0] == plain and set(ref[2]) == {0}
    pieces = run(units, [65536] * 12, 8, 4)[4]
    assert min(pieces) >= 4  # own piece + one per helper: every helper was adopted


def test_block_types_multi_block_and_sizes():
    rng = random.Random(11)
    units, plain = [], []
    for it in range(24):
        n = rng.choice([100, 3000, 20000, 70000, 150000])
        p = text(rng, n) if it % 3 else bytes(rng.randrange(256) for _ in range(n))
        co = zlib.compressobj(rng.choice([0, 1, 6, 9]), zlib.DEFLATED, -15, rng.choice([1, 8, 9]),
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE]))
        h = len(p) // 3
        z = co.compress(p[:h]) + co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_NO_FLUSH])) + \
            co.compress(p[h:]) + co.flush()
        units.append(z + rng.choice([b"", b"\0\0", bytes(8)]))
        plain.append(p)
    caps = [len(p) for p in plain]
    ref = same_as_one_lane(units, caps, "types")
    for i, p in enumerate(plain):
        if ref[2][i] == 0:
            assert ref[0][i] == p
        else:  # quirk Q1: no bytes after the stream, the last symbols are not decoded (inflate.dart:192-195)
            assert p.startswith(ref[0][i]) and orc.inflate(units[i])[1] == ref[0][i]


def test_capacity_too_small_and_exact():
    rng = random.Random(5)
    p = text(rng, 90000)
    z = zlib.compress(p, 6)[2:-4] + bytes(8)
    caps = [90000, 89999, 50000, 20000, 1, 0, 90001, 45000]
    same_as_one_lane([z] * len(caps), caps, "caps")


def test_corrupted_and_truncated_streams():
    rng = random.Random(9)
    base = []
    for it in range(4):
        p = text(rng, rng.choice([30000, 66000]))
        base.append(zlib.compress(p, rng.choice([1, 6, 9]))[2:-4] + bytes(8))
    units = []
    for z in base:
        for _ in range(10):
            zz = bytearray(z)
            for _ in range(rng.choice([1, 1, 3])):
                zz[rng.randrange(len(zz) - 8)] ^= 1 << rng.randrange(8)
            units.append(bytes(zz))
        for _ in range(4):
            units.append(z[:rng.randrange(1, len(z))])
    caps = [70000] * len(units)
    ref = same_as_one_lane(units, caps, "fuzz")
    assert len(set(ref[2])) >= 3  # several different ways of failing were exercised


def test_second_level_tables_and_pool_overflow():
    """Alphabets with many codes longer than the 9/8-bit root tables: the second-level pool (96 + 32 entries per stream)
    overflows and the remaining long codes take the exact step."""
    rng = np.random.default_rng(3)
    units, plain = [], []
    for k in range(6):
        w = 1.0 / np.arange(1, 257) ** (1.0 + 0.15 * k)  # skewed byte frequencies: code lengths up to 15
        p = rng.choice(256, size=60000, p=w / w.sum()).astype(np.uint8).tobytes()
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY if k % 2 else zlib.Z_DEFAULT_STRATEGY)
        units.append(co.compress(p) + co.flush() + bytes(8))
        plain.append(p)
    ref = same_as_one_lane(units, [60000] * 6, "longcodes")
    assert ref[0] == plain
