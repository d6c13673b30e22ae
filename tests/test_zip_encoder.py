"""ZipEncoder mirror (archive_b200/zip.py; zip_encoder.dart:66-583).  CPU tier: the container logic (headers, central
directory, zip64 end records, DOS times, name normalisation) with a stand-in compressor, read back by CPython's zipfile and by
b200z_zip_list.  GPU tier: members compressed on the device, byte-identical payloads to the oracle's Deflate / BZip2Encoder, and a
ZipDecoder round trip (test/zip_test.dart:400-470 does the same round trip)."""
import bz2
import io
import struct
import time
import zipfile
import zlib

import pytest

from archive_b200.zip import Archive, ArchiveFile, ZipDecoder, ZipEncoder


def standin(content, method, level):  # test infrastructure: CPython codecs instead of the device
    crc = zlib.crc32(content)
    if method == "deflate":
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        return co.compress(content) + co.flush(), crc
    if method == "bzip2":
        return bz2.compress(content, 9), crc
    return bytes(content), crc


def make_archive():
    arc = Archive()
    t0 = int(time.mktime((2024, 5, 17, 13, 37, 42, 0, 0, -1)))
    for i, (name, body, comp) in enumerate([("a.txt", b"hello zip " * 100, None), ("dir\\b.bin", bytes(range(256)) * 9, "none"),
                                            ("c.bz2src", b"bzip me " * 500, "bzip2"), ("empty", b"", None)]):
        f = ArchiveFile(name, len(body))
        f.content, f.compression, f.last_mod_time, f.mode = body, comp, t0 + 2 * i, 0o100644 if i % 2 == 0 else 0o100600
        arc.add(f)
    d = ArchiveFile("folder", 0, is_file=False)
    d.last_mod_time, d.mode = t0, 0o40755
    arc.add(d)
    return arc, t0


def test_container_layout_with_standin():
    arc, t0 = make_archive()
    data = ZipEncoder(compress=standin).encode_bytes(arc, level=6, comment="made by a test")
    z = zipfile.ZipFile(io.BytesIO(data))
    assert z.comment == b"made by a test"
    infos = z.infolist()
    assert [i.filename for i in infos] == ["a.txt", "dir/b.bin", "c.bz2src", "empty", "folder/"]
    assert [i.compress_type for i in infos] == [8, 0, 12, 8, 8]
    for i, f in zip(infos, arc.files):
        assert i.flag_bits == 0x800 and i.create_version == 20 and i.extract_version == 20 and i.create_system == 0
        assert i.external_attr == (f.mode << 16) & 0xFFFFFFFF
        if f.is_file:
            assert z.read(i) == f.content and i.CRC == zlib.crc32(f.content) and i.file_size == len(f.content)
    assert infos[0].date_time == (2024, 5, 17, 13, 37, 42) and infos[1].date_time == (2024, 5, 17, 13, 37, 44)
    # `modified` overrides every member's time (zip_encoder.dart:188-191)
    data2 = ZipEncoder(compress=standin).encode_bytes(arc, level=1, modified=t0 + 3600)
    assert all(i.date_time == (2024, 5, 17, 14, 37, 42) for i in zipfile.ZipFile(io.BytesIO(data2)).infolist())
    # the package's own directory reader agrees with zipfile
    dec = ZipDecoder()
    ents, n = dec.list(data)
    assert n == 5 and [e.method for e in dec.entries] == [8, 0, 12, 8, 8]


def test_zip64_end_records_when_there_are_too_many_entries():
    arc = Archive()
    for i in range(0x10000 + 3):
        f = ArchiveFile(f"f{i}", 0)
        f.compression = "none"
        arc.add(f)
    data = ZipEncoder(compress=standin).encode_bytes(arc)
    eocd = data.rfind(b"PK\x05\x06")
    assert struct.unpack_from("<HHHHII", data, eocd + 4) == (0, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert data[eocd - 20:eocd - 16] == b"PK\x06\x07" and data[eocd - 76:eocd - 72] == b"PK\x06\x06"
    assert len(zipfile.ZipFile(io.BytesIO(data)).infolist()) == 0x10003
    ents, n = ZipDecoder().list(data)
    assert n == 0x10003


@pytest.mark.gpu
def test_members_compressed_on_the_device():
    import oracle_lib as orc
    from archive_b200 import synth
    arc, t0 = make_archive()
    big = ArchiveFile("big.txt", 700_000)
    big.content, big.last_mod_time = synth.text(700_000, stream=995).tobytes(), t0
    arc.add(big)
    for level in (1, 6):
        data = ZipEncoder().encode_bytes(arc, level=level)
        ref = ZipEncoder(compress=lambda c, m, l: ((orc.deflate(c, l)[1] if m == "deflate" else orc.bzip2_encode(c)[1] if m == "bzip2"
                                                   else bytes(c)), zlib.crc32(c))).encode_bytes(arc, level=level)
        assert data == ref  # payloads are the reference's Deflate / BZip2Encoder output, CRCs from the device
        back = ZipDecoder().decode_bytes(data)
        assert [(f.name, f.content) for f in back.files if f.is_file] == \
               [(f.name.replace("\\", "/"), f.content) for f in arc.files if f.is_file]
        assert zipfile.ZipFile(io.BytesIO(data)).read("big.txt") == big.content


def test_container_equals_the_oracle_restatement():
    """The Python mirror's container bytes against oracle/zip_enc.c (a second, independent restatement of
    zip_encoder.dart:158-497), members compressed by the oracle's Deflate / BZip2Encoder in both."""
    import time as _t

    import oracle_lib as orc
    from archive_b200.zip import _dos_date, _dos_time
    arc, t0 = make_archive()
    for f in arc.files:
        if f.name == "a.txt":
            f.comment = "a comment"
    comp = lambda c, m, l: ((orc.deflate(c, l)[1] if m == "deflate" else orc.bzip2_encode(c)[1] if m == "bzip2" else bytes(c)),
                            zlib.crc32(c))
    for level in (1, 6):
        mine = ZipEncoder(compress=comp).encode_bytes(arc, level=level, comment="zc")
        members = []
        for f in arc.files:
            lm = _t.localtime(f.last_mod_time)
            name = f.name.replace("\\", "/") + ("/" if not f.is_file and not f.name.endswith("/") else "")
            members.append((name, f.content or b"", (f.compression or "deflate") if f.is_file else "deflate", f.is_file, f.mode,
                            _dos_time(lm), _dos_date(lm), getattr(f, "comment", None)))
        st, ref = orc.zip_encode(members, level=level, comment="zc")
        assert st == orc.OK and mine == ref


def test_members_sharded_over_ranks_give_the_same_archive():
    """the encoder's multi-GPU axis.  Ranks compress disjoint shares (largest first), exchange the payloads, and
    any of them writes the container a single rank would have written."""
    from archive_b200 import shard
    arc, t0 = make_archive()
    for k in range(9):
        f = ArchiveFile(f"more/{k}.txt", 0)
        f.content = (b"member %d " % k) * (50 + 37 * k)
        f.size, f.last_mod_time, f.mode = len(f.content), t0 + k, 0o100644
        arc.add(f)
    want = ZipEncoder(compress=standin).encode_bytes(arc, level=6, comment="c")
    for world in (1, 2, 3, 5):
        shares = [shard.zip_encode_sharded(arc, level=6, rank=r, world=world, compress=standin) for r in range(world)]
        assert sorted(i for s in shares for i in s) == [i for i, f in enumerate(arc.files) if f.is_file]
        assert sum(len(s) for s in shares) == len({i for s in shares for i in s})  # disjoint
        for r in range(world):
            got = shard.zip_encode_sharded(arc, level=6, comment="c", rank=r, world=world, compress=standin,
                                           payloads_in=[s for q, s in enumerate(shares) if q != r])
            assert got == want, (world, r)


def _zip_rank(rank, world, port, q):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    from archive_b200 import shard
    import test_zip_encoder as t
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arc, t0 = t.make_archive()
    calls = []

    def counting(content, method, level):
        calls.append(len(content))
        return t.standin(content, method, level)

    data = shard.zip_encode_sharded(arc, level=6, compress=counting)
    want = t.ZipEncoder(compress=t.standin).encode_bytes(arc, level=6)
    q.put((rank, data == want, len(calls)))
    dist.destroy_process_group()


def test_two_gloo_ranks_encode_one_archive():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_zip_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert [r[1] for r in res] == [True, True]
    assert sum(r[2] for r in res) == 4 and all(r[2] > 0 for r in res)  # the four file members, split over the two ranks


@pytest.mark.gpu
def test_sharded_encode_on_the_device_equals_one_rank():
    """The same, with the members compressed by the kernels (one process playing every rank in turn)."""
    from archive_b200 import shard
    arc, t0 = make_archive()
    want = ZipEncoder().encode_bytes(arc, level=6)
    shares = [shard.zip_encode_sharded(arc, level=6, rank=r, world=2) for r in range(2)]
    assert shard.zip_encode_sharded(arc, level=6, rank=0, world=2, payloads_in=shares[1:]) == want


def test_zip_file_encoder_on_a_directory_tree(tmp_path):
    """ZipFileEncoder (io/zip_file_encoder.dart:11-225) with the stand-in compressor: directory walk, names relative to the
    directory (with and without its own name in front), per-file level, modes and times from the file system, the archive
    written through an OutputFileStream; zipDirectory's default name and its guard."""
    import os
    from archive_b200 import ZipFileEncoder
    root = tmp_path / "tree"
    (root / "sub" / "deep").mkdir(parents=True)
    (root / "empty").mkdir()
    files = {"a.txt": b"alpha " * 300, "sub/b.bin": bytes(range(256)) * 5, "sub/deep/c.txt": b"c" * 1000, "z.dat": b""}
    for rel, body in files.items():
        (root / rel).write_bytes(body)
    os.chmod(root / "a.txt", 0o640)
    levels = []

    def counting(content, method, level):
        levels.append(level)
        return standin(content, method, level)

    enc = ZipFileEncoder(compress=counting)
    n = enc.zip_directory(str(root))  # default name: <dir>.zip, level 1 (:22-43)
    zpath = str(root) + ".zip"
    assert os.path.getsize(zpath) == n and set(levels) == {1}
    z = zipfile.ZipFile(zpath)
    names = z.namelist()
    assert names == sorted(names) and set(names) == {"a.txt", "empty/", "sub/", "sub/b.bin", "sub/deep/", "sub/deep/c.txt", "z.dat"}
    for rel, body in files.items():
        assert z.read(rel) == body
    assert (z.getinfo("a.txt").external_attr >> 16) & 0o777 == 0o640
    with pytest.raises(ValueError):
        ZipFileEncoder(compress=standin).zip_directory(str(root), filename=str(root / "inside.zip"))
    # create / add_file / add_directory / add_archive_file / close, with the directory's name in front and per-file levels
    del levels[:]
    enc = ZipFileEncoder(compress=counting)
    enc.create(str(tmp_path / "out" / "x.zip"), level=6)
    enc.add_file(str(root / "a.txt"), level=9)
    enc.add_file(str(root / "sub" / "b.bin"), "renamed/b.bin")
    enc.add_directory(str(root), filter=lambda p, prog: "skip" if p.endswith("z.dat") else None)
    extra = ArchiveFile("extra.txt", 5)
    extra.content, extra.last_mod_time = b"extra", 86400 * 400
    enc.add_archive_file(extra)
    enc.close()
    z = zipfile.ZipFile(str(tmp_path / "out" / "x.zip"))
    assert z.namelist()[:2] == ["a.txt", "renamed/b.bin"] and "tree/sub/deep/c.txt" in z.namelist() and "tree/z.dat" not in z.namelist()
    assert z.read("tree/sub/b.bin") == files["sub/b.bin"] and z.read("extra.txt") == b"extra"
    assert levels[0] == 9 and set(levels[1:]) == {6}
    # the same members through ZipEncoder give the same bytes
    again = ZipFileEncoder(compress=standin)
    again.create(str(tmp_path / "y.zip"), level=6)
    again.add_file(str(root / "a.txt"))
    again.close()
    f = ArchiveFile("a.txt", len(files["a.txt"]))
    st = os.stat(root / "a.txt")
    f.content, f.last_mod_time, f.mode = files["a.txt"], int(st.st_mtime), st.st_mode
    assert open(str(tmp_path / "y.zip"), "rb").read() == ZipEncoder(compress=standin).encode_bytes([f], level=6)
