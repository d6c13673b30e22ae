"""BZip2 block sharding: the host-side chain walk over merged per-rank block reports (CPU), its exchange
between two gloo ranks (CPU), and -- on the GPU -- b200z_bzip2_decode_shard played for 1, 2 and 3 ranks on one device."""
import bz2
import os
import socket

import pytest

from archive_b200 import shard


def rep(start, end, out, crc, stored=None, status=0, flags=0, rank=0, off=0):
    return (start, end, out, crc, crc if stored is None else stored, status, flags, rank, off)


def comb(crcs):
    c = 0
    for x in crcs:
        c = (((c << 1) | (c >> 31)) & 0xFFFFFFFF) ^ x
    return c


def test_chain_walk_rules():
    n = 1000
    good = [rep(32, 3000, 100, 11), rep(3000, 6000, 200, 22, rank=1), rep(6000, 6080, 0, 0, stored=comb([11, 22]), flags=shard.BZ2_EOS)]
    assert shard.bz2_walk_chain(good, n, True)[::2] == ("ok", 300)
    # a magic-looking pattern inside a block is reported but is not on the chain
    junk = good + [rep(4000, 4100, 7, 1, status=-1)]
    assert shard.bz2_walk_chain(junk, n, True)[::2] == ("ok", 300)
    # gap: the second block does not start where the first ended -> false, first block kept
    gap = [good[0], rep(3001, 6000, 200, 22), good[2]]
    assert shard.bz2_walk_chain(gap, n, False)[::2] == ("data", 100)
    # CRC mismatch is only seen with verify, and the bad block's bytes are already written (bzip2_decoder.dart:58-66)
    bad = [good[0], rep(3000, 6000, 200, 22, stored=23, rank=1), good[2]]
    assert shard.bz2_walk_chain(bad, n, False)[::2] == ("ok", 300)
    assert shard.bz2_walk_chain(bad, n, True)[::2] == ("data", 300)
    # combined CRC
    eos_bad = good[:2] + [rep(6000, 6080, 0, 0, stored=5, flags=shard.BZ2_EOS)]
    assert shard.bz2_walk_chain(eos_bad, n, True)[0] == "data" and shard.bz2_walk_chain(eos_bad, n, False)[0] == "ok"
    # block that read past the end: RangeError; randomised: false
    assert shard.bz2_walk_chain([good[0], rep(3000, 0, 0, 0, status=-2)], n, False)[::2] == ("throw", 100)
    assert shard.bz2_walk_chain([good[0], rep(3000, 6000, 0, 0, flags=shard.BZ2_RANDOMISED)], n, False)[::2] == ("data", 100)
    # stream that just ends after a block (no EOS): the loop stops at isEOS
    assert shard.bz2_walk_chain([rep(32, 7995, 100, 11)], n, True)[::2] == ("ok", 100)
    assert shard.bz2_walk_chain([rep(32, 7990, 100, 11)], n, True)[::2] == ("throw", 100)  # 10 bits left: _readBlockType throws


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [rep(32, 3000, 100, 11, rank=0)] if rank == 0 else [rep(3000, 6000, 200, 22, rank=1)]
    mine.append(rep(6000, 6080, 0, 0, stored=comb([11, 22]), flags=shard.BZ2_EOS, rank=rank))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    kind, chain, n_out = shard.bz2_walk_chain([r for part in gathered for r in part], 1000, True)
    if rank == 0:
        q.put((kind, n_out, [r[7] for r in chain]))
    dist.destroy_process_group()


def test_two_rank_report_exchange_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(60)
    assert got == ("ok", 300, [0, 1])


@pytest.mark.gpu
def test_shards_reassemble_on_one_gpu():
    """Each 'rank' is a call of b200z_bzip2_decode_shard on the same device; the merged reports must give back the
    stream, identically for every world size, and find corruption where the unsharded decoder finds it."""
    import archive_b200
    from archive_b200 import synth
    src = synth.text(9_500_000, stream=77).tobytes()
    z = bz2.compress(src, 9)
    for world in (1, 2, 3):
        parts = [shard.bzip2_decode_sharded(z, rank=r, world=world) for r in range(world)]
        reports = [x for p in parts for x in p["reports"]]
        kind, chain, n_out = shard.bz2_walk_chain(reports, len(z), True)
        assert kind == "ok" and n_out == len(src)
        out = bytearray(n_out)
        for r, p in enumerate(parts):
            others = [x for q, pp in enumerate(parts) if q != r for x in pp["reports"]]
            mine = shard.bzip2_decode_sharded(z, verify=True, rank=r, world=world, reports_in=others)
            assert mine["kind"] == "ok" and mine["total"] == len(src)
            for off, v in mine["pieces"]:
                out[off:off + len(v)] = v
        assert bytes(out) == src
    bad = bytearray(z)
    bad[len(z) // 2] ^= 0x10
    ref_ok = archive_b200.BZip2Decoder().decode_bytes(bytes(bad), verify=True)
    parts = [shard.bzip2_decode_sharded(bytes(bad), rank=r, world=2) for r in range(2)]
    kind, chain, n_out = shard.bz2_walk_chain([x for p in parts for x in p["reports"]], len(bad), True)
    assert kind == "data" and n_out == len(ref_ok)


def test_chain_walk_overrun_and_short_signature():
    # a block whose run-length walk overran: its bytes count, then the stream is false -- whatever `verify` says
    n = 1000
    blocks = [rep(32, 3000, 100, 11), rep(3000, 6000, 250, 22, flags=shard.BZ2_OVERRUN), rep(6000, 7000, 50, 33)]
    for verify in (False, True):
        kind, chain, n_out = shard.bz2_walk_chain(blocks, n, verify)
        assert (kind, n_out, len(chain)) == ("data", 350, 2)
    # the stream ends inside the next block signature (bzip2_decoder.dart:90-111): a byte that fits neither magic -> false,
    # a proper prefix of either magic -> RangeError; without the stream at hand the walk can only say RangeError
    head = [rep(32, 7968, 10, 5)]
    for tail, want in ((b"\x31\x41\x59", "throw"), (b"\x17\x72", "throw"), (b"\x31\x42", "data"), (b"\x00", "data"),
                       (b"\x17\x72\x45\x38\x50", "throw"), (b"\x31\x41\x59\x26\x53\x58"[:5], "throw")):
        data = bytes(996) + tail
        assert shard.bz2_walk_chain(head, len(data), False, data=data)[0] == want, tail
        assert shard.bz2_walk_chain(head, len(data), False)[0] == "throw"
    # not byte aligned: the signature starts 3 bits into a byte
    bits = "0" * (8 * 996 + 3) + "".join(f"{b:08b}" for b in b"\x31\x41") + "00000"
    data = int(bits, 2).to_bytes(len(bits) // 8, "big")
    assert shard.bz2_walk_chain([rep(32, 8 * 996 + 3, 10, 5)], len(data), False, data=data)[0] == "throw"
    bits = "0" * (8 * 996 + 3) + "".join(f"{b:08b}" for b in b"\x31\x40") + "00000"
    data = int(bits, 2).to_bytes(len(bits) // 8, "big")
    assert shard.bz2_walk_chain([rep(32, 8 * 996 + 3, 10, 5)], len(data), False, data=data)[0] == "data"


def test_zip_member_packing():
    sizes = [10, 500, 20, 499, 498, 1, 0, 300]
    bins = shard.pack_members(sizes, 3)
    assert sorted(i for b in bins for i in b) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(sizes)  # largest-first greedy: within one member of each other
    assert bins == shard.pack_members(sizes, 3)  # deterministic: every rank computes the same plan
    assert shard.pack_members(sizes, 1) == [list(range(len(sizes)))]
    assert shard.pack_members([], 4) == [[], [], [], []]


@pytest.mark.gpu
def test_zip_members_sharded_on_one_gpu():
    import io
    import zipfile
    from archive_b200 import synth
    txt = synth.text(40 * 200_000, stream=990).tobytes()
    buf = io.BytesIO()
    want = []
    with zipfile.ZipFile(buf, "w") as z:
        for i in range(40):
            body = txt[i * 200_000:i * 200_000 + 1000 * (i % 7) * (i % 5) * 8 + i]
            z.writestr(f"f{i}", body, compress_type=zipfile.ZIP_DEFLATED if i % 3 else zipfile.ZIP_STORED)
            want.append(body)
    data = buf.getvalue()
    for world in (1, 3):
        got = {}
        for r in range(world):
            ents, part = shard.zip_extract_sharded(data, r, world)
            assert not set(part) & set(got)
            got.update(part)
        assert [got[i] for i in range(40)] == want
