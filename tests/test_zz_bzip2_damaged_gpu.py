"""BZip2 decode of DAMAGED streams on the sm_100a path, through the C ABI, against the oracle: the reference keeps decoding
after a bad Huffman code (bzip2_decoder.dart:273-387), walks short inverse-BWT cycles and lets runs overrun the block
(:497-499, :628-631), and the bytes it has written by then count.  These cases came out of the CPU-tier fuzz
(tests/test_bzip2_dec_emul.py).  The file sorts last on purpose: valid-stream parity is established before damaged data is
thrown at the device, and each test carries a hard time limit."""
import bz2
import hashlib
import json
import os
import random

import pytest

import oracle_lib as orc
from archive_b200 import shard

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
G = os.path.join(os.path.dirname(__file__), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))


def rd(n):
    return open(os.path.join(G, n), "rb").read()


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


DAMAGED = ("bz2_mtfval_quirk_a.bz2", "bz2_mtfval_quirk_b.bz2", "bz2_mtfval_quirk_c.bz2", "bz2_short_cycle.bz2",
           "bz2_rand_overrun_a.bz2", "bz2_rand_overrun_b.bz2", "bz2_run_at_block_end_a.bz2", "bz2_run_at_block_end_b.bz2")


def _decode(a, z, verify):
    """-> (oracle-style status, bytes written) of BZip2Decoder.decodeStream through the C ABI."""
    out = a.OutputMemoryStream()
    try:
        ok = a.BZip2Decoder().decode_stream(a.InputMemoryStream(z), out, verify=verify)
    except a.DartRangeError:
        return orc.THROW, out.get_bytes()
    return (orc.OK if ok else orc.FALSE), out.get_bytes()


def test_damaged_fixtures(a):
    """Damaged streams the CPU-tier fuzz found (tests/test_bzip2_dec_emul.py; manifest.json says what each one is): the
    reference keeps decoding after a bad Huffman code (literal entropy kernel), walks a short inverse-BWT cycle, lets a run
    overrun the block -- and the bytes it has written by then count."""
    for name in DAMAGED:
        z = rd(name)
        for verify in (False, True):
            ost, oout = orc.bzip2_decode(z, verify=verify)
            st, out = _decode(a, z, verify)
            assert st == ost and (st == orc.THROW or out == oout), (name, verify, st, ost, len(out), len(oout))
        st, out = _decode(a, z, False)
        assert st == MAN[name]["status"] and hashlib.sha256(out).hexdigest() == MAN[name]["sha256"], name


def test_fuzz_damage_vs_oracle(a):
    """Seeded damage (bit flips, overwrites, truncation, the randomised flag) of small streams: same verdict and bytes as the
    oracle, whatever they are."""
    rng = random.Random(0xB200)
    n = 0
    for r in range(24):
        k = r % 3
        if k == 0:
            src = bytes(rng.randrange(rng.choice([3, 7, 256])) for _ in range(rng.randrange(200, 30000)))
        elif k == 1:
            src = b"".join(bytes([rng.randrange(3)]) * rng.choice([1, 2, 4, 5, 255, 256, 1000]) for _ in range(rng.randrange(1, 400)))
        else:
            src = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40))) * rng.randrange(1, 2000)
        z = bz2.compress(src, rng.choice([1, 1, 9]))
        for _ in range(8):
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
            bad = bytes(bad)
            for verify in (False, True):
                ost, oout = orc.bzip2_decode(bad, verify=verify)
                st, out = _decode(a, bad, verify)
                assert st == ost and (st == orc.THROW or out == oout), (r, kind, verify, st, ost, len(out), len(oout))
            n += 1
    assert n == 192


def test_shards_agree_on_damaged_streams():
    """The damaged-stream fixtures (tests/golden/manifest.json) through the sharded path: bytes a block has written before the
    reference notices an overrun are kept (B200Z_BZ2_OVERRUN), a block decoded by the literal entropy kernel is reported
    like any other."""
    import hashlib
    import json
    import os
    import oracle_lib as orc
    G = os.path.join(os.path.dirname(__file__), "golden")
    man = json.load(open(os.path.join(G, "manifest.json")))
    for name in ("bz2_mtfval_quirk_a.bz2", "bz2_short_cycle.bz2", "bz2_rand_overrun_a.bz2", "bz2_run_at_block_end_a.bz2"):
        z = open(os.path.join(G, name), "rb").read()
        for world in (1, 2):
            parts = [shard.bzip2_decode_sharded(z, rank=r, world=world) for r in range(world)]
            reports = [x for p in parts for x in p["reports"]]
            out = bytearray()
            total = None
            for r, p in enumerate(parts):
                others = [x for q, pp in enumerate(parts) if q != r for x in pp["reports"]]
                mine = shard.bzip2_decode_sharded(z, verify=False, rank=r, world=world, reports_in=others)
                total = mine["total"]
                out.extend(b"\0" * max(0, total - len(out)))
                for off, v in mine["pieces"]:
                    out[off:off + len(v)] = v
                kind = mine["kind"]
            st = {"ok": orc.OK, "data": orc.FALSE, "throw": orc.THROW}[kind]
            assert st == man[name]["status"] and total == man[name]["size"], (name, world, kind, total)
            assert hashlib.sha256(bytes(out)).hexdigest() == man[name]["sha256"], (name, world)
