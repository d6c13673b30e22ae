"""The device BZip2 encoder (archive_b200/csrc/bzip2_enc_*.{cu,inl}) executed on the CUDA execution-model emulation
(tests/host_emul/cuda_emu.h) and compared byte for byte with the oracle restatement of bzip2_encoder.dart.
This is the CPU-tier cover of the kernels' logic; tests/test_bzip2_enc_gpu.py runs the same comparison on the B200."""
import bz2
import random

import numpy as np
import pytest

import oracle_lib as orc


def both(data):
    rc, got, st = orc.emul_bzip2_encode(data)
    st_o, ref = orc.bzip2_encode(data)
    assert st_o == orc.OK
    assert rc == 0
    assert got == ref
    return got, st


SMALL = {
    "empty": b"",
    "one": b"a",
    "abc": b"abcabcabd",
    "rand1k": bytes(random.Random(1).randrange(256) for _ in range(1000)),
    "text": (b"the quick brown fox jumps over the lazy dog. " * 200)[:7001],
    # every RLE1 run shape of _addPairToBlock (:2033-2071): 1,2,3 | 4 | 5 | 255 | 256 | 259 | 600
    "runs": b"a" * 3 + b"b" * 4 + b"c" * 5 + b"d" * 255 + b"e" * 256 + b"f" * 259 + b"g" * 600 + b"xyz",
    "rand20k": bytes(random.Random(2).randrange(256) for _ in range(20000)),
    "lowent": bytes(random.Random(3).choice(b"ab") for _ in range(30000)),
}


@pytest.mark.parametrize("name", list(SMALL))
def test_small(name):
    got, st = both(SMALL[name])
    assert bz2.decompress(got) == SMALL[name]
    assert st[1] == 0  # no periodic block


# periodic blocks: the order among identical rotations is whatever the reference's sort leaves (k_serial_sort)
PERIODIC = {
    "hello3": b"hello hello hello world, hello!" * 3,  # nblock < 10000: _fallbackSort
    "xyxy": b"xyxy" * 500,
    "aaa": b"aaa",
    "abc90k": b"abc" * 30000,  # _mainSort, budget exhausted, _fallbackSort
    "zeros": bytes(255 * 4000),  # RLE1 turns it into (0,0,0,0,251) x 4000
}


@pytest.mark.parametrize("name", list(PERIODIC))
def test_periodic(name):
    got, st = both(PERIODIC[name])
    assert st[1] == 1


def test_multi_block_cut_rules():
    """Block cuts (_writeBlock :83-110): runs straddling the cut, the byte that closes the last run, a long run that
    fills a block by itself."""
    r = np.random.default_rng(5)
    a = r.integers(0, 256, 1_900_000, dtype=np.uint8)
    a[899_000:901_500] = 7
    a[1000:1600] = 9
    got, st = both(a.tobytes())
    assert st[0] == 3
    z = np.zeros(3_000_000, dtype=np.uint8)
    z[1_234_567] = 1
    both(z.tobytes())
    q = r.integers(0, 4, 1_000_000, dtype=np.uint8)  # many runs of 4 and more: RLE1 expands, block cut moves
    got, st = both(q.tobytes())
    assert st[0] == 2
    # 46 MB of one byte: the first block is a single run cut by the 899 981-byte rule, the rest follows
    both(bytes(46_000_000) + b"tail")
