"""Parity tests for the sm_100a BZip2 encoder (bzip2_enc_*.{cu,inl}), through the C ABI, against the oracle restatement
of bzip2_encoder.dart (byte-identical output), libbz2 (decodes every stream; identical bytes for single-block inputs)
and the repo's own decoder (round trip, test/bzip2_test.dart:14-25)."""
import bz2
import os
import random

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


def enc(a, data):
    return a.BZip2Encoder().encode_bytes(data)


def test_small_and_periodic(a):
    import test_bzip2_enc_emul as cases
    for name, data in list(cases.SMALL.items()) + list(cases.PERIODIC.items()):
        z = enc(a, data)
        assert z == orc.bzip2_encode(data)[1], name
        assert bz2.decompress(z) == data, name


def test_roundtrip_cat_jpg(a):  # test/bzip2_test.dart:14-25
    src = open(os.path.join(G, "cat.jpg"), "rb").read()
    z = enc(a, src)
    assert z == orc.bzip2_encode(src)[1]
    assert z == bz2.compress(src, 9)  # single block: identical to libbzip2
    assert a.BZip2Decoder().decode_bytes(z, verify=True) == src


def test_multi_block_text_identical_to_oracle(a):
    from archive_b200 import synth
    src = synth.text(12 << 20, stream=300).tobytes()
    z = enc(a, src)
    assert z == orc.bzip2_encode(src)[1]
    assert bz2.decompress(z) == src


def test_block_cut_rules(a):
    r = np.random.default_rng(5)
    x = r.integers(0, 256, 1_900_000, dtype=np.uint8)
    x[899_000:901_500] = 7
    x[1000:1600] = 9
    zeros = np.zeros(3_000_000, dtype=np.uint8)
    zeros[1_234_567] = 1
    quad = r.integers(0, 4, 3_000_000, dtype=np.uint8)
    for name, data in (("runs", x.tobytes()), ("zeros", zeros.tobytes()), ("quad", quad.tobytes()),
                       ("onebyte", bytes(100_000_000) + b"tail")):
        z = enc(a, data)
        assert z == orc.bzip2_encode(data)[1], name


def test_repetitive_blocks(a):
    """Worst case for the doubling sort (every rotation stays unresolved for ~17 rounds) and the periodic fallback."""
    rep = b"0123456789abcdefghij" * 150_000  # 3 MB, period 20: blocks are NOT periodic (cut at 899 982), but nearly
    z = enc(a, rep)
    assert z == orc.bzip2_encode(rep)[1]
    per = b"abcdefgh" * 112_497  # 899 976 bytes in one block, period 8 -> identical rotations
    z = enc(a, per)
    assert z == orc.bzip2_encode(per)[1]


def test_large_roundtrip(a):
    """128 MiB: too slow for the single-threaded oracle; libbz2 and the repo's decoder must give the input back, and
    the stream's block count must be the reference's (block cuts are data independent of the sort)."""
    from archive_b200 import synth
    src = synth.text(128 << 20, stream=400).tobytes()
    z = enc(a, src)
    assert a.BZip2Decoder().decode_bytes(z, verify=True) == src
    assert bz2.decompress(z[:]) == src
