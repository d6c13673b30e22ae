"""The data-parallel REFORMULATION of Deflate levels 4-9 that the CUDA encoder implements (match table under two
chain budgets -> 4-state parse graph -> block cuts from prefix sums -> per-block trees), run as a CPU model
(tests/host_emul/deflate_model.cpp) and compared byte-for-byte with the oracle's line-by-line restatement of
deflate.dart.  Proves the reformulation; the GPU tests then check the kernels against the same oracle."""
import ctypes as C
import os
import random
import subprocess

import pytest

import oracle_lib as orc
from archive_b200 import synth

HERE = os.path.join(os.path.dirname(__file__), "host_emul")


@pytest.fixture(scope="module")
def model():
    so, src = os.path.join(HERE, "libdeflate_model.so"), os.path.join(HERE, "deflate_model.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", src, "-o", so], check=True)
    M = C.CDLL(so)

    def run(d, level):
        out, n, nt, nb = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_uint64(), C.c_uint64()
        assert M.model_deflate(d, C.c_size_t(len(d)), level, C.byref(out), C.byref(n), C.byref(nt), C.byref(nb)) == 0
        return C.string_at(out, n.value), nt.value, nb.value

    return run


def test_model_equals_oracle(model):
    rng = random.Random(3)
    t = synth.text(2 << 20).tobytes()
    cases = {
        "text": t, "empty": b"", "a": b"a", "ab": b"ab", "abc": b"abc", "zeros": b"\0" * 300000,
        "rand": bytes(rng.getrandbits(8) for _ in range(150000)), "mod256": bytes(i % 256 for i in range(0xfffff)),
        "short": t[:1000], "mix": t[:50000] + bytes(rng.getrandbits(8) for _ in range(50000)) + t[:50000],
        "rep": t[:700] * 500, "far": t[:40000] + bytes(rng.getrandbits(8) for _ in range(32500)) + t[:40000],
        "tail": t[:65536 - 3], "tail2": t[:65536 + 261],
    }
    for name, d in cases.items():
        for level in range(4, 10):
            z, ntok, nblk = model(d, level)
            assert z == orc.deflate(d, level)[1], (name, level)


def test_model_exercises_the_early_flush_heuristic(model):
    """On text the reference's TRUNCATE_BLOCK rule fires (blocks of 8192 symbols), so there are clearly more
    blocks than ceil(tokens / 16383)."""
    t = synth.text(4 << 20).tobytes()
    z, ntok, nblk = model(t, 6)
    assert nblk > (ntok + 16382) // 16383 + 5
    assert z == orc.deflate(t, 6)[1]


def test_model_window_bits_and_the_stored_block_rule():
    """windowBits 9..15: MAX_DIST shrinks, and a block whose start has slid out of the 2 x wSize window (buf == -1,
    deflate.dart:677-680 + the slide in _fillWindow :816-857) may not be stored.  Incompressible data hits that rule."""
    so = os.path.join(HERE, "libdeflate_model.so")
    M = C.CDLL(so)

    def run(d, level, wb):
        out, n, nt, nb = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_uint64(), C.c_uint64()
        assert M.model_deflate_wb(d, C.c_size_t(len(d)), level, wb, C.byref(out), C.byref(n), C.byref(nt), C.byref(nb)) == 0
        return C.string_at(out, n.value)

    rng = random.Random(1)
    t = synth.text(400000).tobytes()
    rnd = bytes(rng.getrandbits(8) for _ in range(100000))
    for d in (t, t[:5000] * 40, rnd, rnd[:30000] + t[:100000] + rnd[:50000], rnd[:16383], rnd[:16384], rnd[:33000]):
        for wb in (9, 10, 12, 14, 15):
            for level in (4, 6, 9):
                assert run(d, level, wb) == orc.deflate(d, level, wb)[1], (len(d), wb, level)
