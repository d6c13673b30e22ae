"""InputFileStream -> codec -> OutputFileStream on the sm_100a path (b200z_file_codec, csrc/b200z_file.cu):
the file must hold exactly what the memory entry points and the oracle produce for the same bytes -- segmented gzip decode
(segments cut at member boundaries, forced small here), members without size hints, a lying hint, bad data with its
partial output, stream positions either side.  Mirrors the reference's stream tests (test/io_test.dart:471-492 'stream gzip
encode / decode', test/zlib_test.dart:35-58 'encodeStream', io/extract_archive_to_disk.dart:183-202).
Also here, for the same reason (written after the round's last GPU run; the file sorts last so that the established parity
tests come first): b200z_deflate_batch / ZipEncoder(batch=True), the chunk pipeline of b200z_gzip_decode with
test-sized chunks, extract_file_to_disk."""
import bz2
import os
import struct
import zlib

import pytest

import oracle_lib as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]
G = os.path.join(os.path.dirname(__file__), "golden")


def rd(n):
    return open(os.path.join(G, n), "rb").read()


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


def _members(text: bytes, unit: int, hint=True):
    from archive_b200 import synth
    return [synth.gzip_member(text[i:i + unit], 6, hint) for i in range(0, len(text), unit)]


def _file_decode(a, dec, tmp_path, blob: bytes, name="in.gz", prefix=b"", lead=b"", **kw):
    """decodeStream(InputFileStream, OutputFileStream) -> (bool | 'throw', bytes in the file after the prefix)."""
    src, dst = str(tmp_path / name), str(tmp_path / (name + ".out"))
    open(src, "wb").write(lead + blob)
    inp = a.InputFileStream(src)
    inp.skip(len(lead))
    out = a.OutputFileStream(dst)
    out.write_bytes(prefix)
    try:
        ok = dec.decode_stream(inp, out, **kw)
    except a.DartRangeError:
        ok = "throw"
    assert inp.is_eos and inp.position == len(lead) + len(blob)
    n = out.length
    out.close_sync()
    got = open(dst, "rb").read()
    assert len(got) == n and got[:len(prefix)] == prefix
    return ok, got[len(prefix):]


def _stats():
    import ctypes as C
    from archive_b200 import _ffi
    seg, whole = C.c_uint32(0), C.c_uint32(0)
    _ffi.lib().b200z_file_last_stats(C.byref(seg), C.byref(whole))
    return seg.value, whole.value


def _mem_decode(a, dec, blob: bytes, **kw):
    out = a.OutputMemoryStream()
    try:
        ok = dec.decode_stream(a.InputMemoryStream(blob), out, **kw)
    except a.DartRangeError:
        ok = "throw"
    return ok, out.get_bytes()


def test_gzip_segments_equal_one_call(a, tmp_path, monkeypatch):
    from archive_b200 import synth
    text = synth.text(40 * 16384, stream=41).tobytes()
    ms = _members(text, 16384)  # ~6.5 KiB each: ten to a 64 KiB segment
    blob = b"".join(ms)
    assert len(blob) > 4 * 65536
    want = orc.gzip_decode(blob)
    assert want == (orc.OK, text)
    monkeypatch.setenv("B200Z_FILE_SEG_KB", "64")
    ok, got = _file_decode(a, a.GZipDecoder(), tmp_path, blob, prefix=b"already there:", lead=b"skipped header bytes")
    assert ok is True and got == text
    assert _stats()[0] >= 4 and _stats()[1] == 0  # every byte went through the segment pipeline
    monkeypatch.setenv("B200Z_FILE_THREADS", "3")
    monkeypatch.setenv("B200Z_FILE_SEG_KB", "100")  # segment ends fall inside members: the tail is read again
    ok, got = _file_decode(a, a.GZipDecoder(), tmp_path, blob, name="b.gz")
    assert ok is True and got == text
    monkeypatch.delenv("B200Z_FILE_SEG_KB")  # default segment (256 MiB): one call over the whole range
    ok, got = _file_decode(a, a.GZipDecoder(), tmp_path, blob, name="c.gz")
    assert ok is True and got == text
    assert _stats() == (0, 1)


def test_gzip_mixed_members_and_bad_data(a, tmp_path, monkeypatch):
    """Members without hints, a hint that lies, garbage after the last member, a truncated file: the verdict and the bytes
    are those of the memory entry point and of the oracle, segmented or not."""
    from archive_b200 import synth
    text = synth.text(30 * 16384, stream=42).tobytes()
    hinted = _members(text, 16384)
    plain = _members(text, 16384, hint=False)
    lying = bytearray(hinted[17])
    lying[-4:] = struct.pack("<I", 16000)  # ISIZE says less than the member holds
    short = bytearray(hinted[12])
    short[16:18] = struct.pack("<H", len(short) - 40 - 1)  # the 'BC' size ends inside the member
    huge = bytearray(hinted[9])
    huge[-4:] = struct.pack("<I", 0xFFFFFFF0)  # an ISIZE no DEFLATE stream of this size can reach: not a hint at all
    cases = {
        "huge_isize": b"".join(hinted[:9] + [bytes(huge)] + hinted[10:]),
        "nohint_middle": b"".join(hinted[:14] + plain[14:16] + hinted[16:]),
        "nohint_first": b"".join(plain[:1] + hinted[1:]),
        "all_plain": b"".join(plain[:6]),
        "lying_isize": b"".join(hinted[:17] + [bytes(lying)] + hinted[18:]),
        "short_bsize": b"".join(hinted[:12] + [bytes(short)] + hinted[13:]),
        "garbage_tail": b"".join(hinted) + b"\x1f\x8b\x08\x00" + bytes(range(40)),
        "zlib_fallback_tail": b"".join(hinted[:12]) + zlib.compress(text[:5000]),
        "truncated": b"".join(hinted)[:-3000],
        "truncated_in_header": b"".join(hinted[:13]) + hinted[13][:7],
        "empty": b"",
    }
    for seg in ("64", None):
        if seg:
            monkeypatch.setenv("B200Z_FILE_SEG_KB", seg)
        else:
            monkeypatch.delenv("B200Z_FILE_SEG_KB")
        for name, blob in cases.items():
            mem = _mem_decode(a, a.GZipDecoder(), blob)
            got = _file_decode(a, a.GZipDecoder(), tmp_path, blob, name=name + ".gz")
            assert got[0] == mem[0] and got[1] == mem[1], (name, seg, got[0], mem[0], len(got[1]), len(mem[1]))
            if seg and name in ("nohint_middle", "lying_isize", "short_bsize", "garbage_tail", "truncated"):
                assert _stats()[0] >= 1 and _stats()[1] == 1, (name, _stats())  # segments first, then the rest in one piece
            ost, oout = orc.gzip_decode(blob)
            if ost == orc.OK:  # (bad data: DESIGN.md "Divergences" -- the memory entry point is the yardstick above)
                assert got == (True, oout), (name, seg)


def test_reference_stream_round_trips(a, tmp_path):
    cat = rd("cat.jpg")
    # io_test.dart:471-492: GZipEncoder.encodeStream file -> file, then GZipDecoder.decodeStream file -> file
    enc_out = a.OutputFileStream(str(tmp_path / "cat.jpg.gz"))
    a.GZipEncoder().encode_stream(a.InputFileStream(os.path.join(G, "cat.jpg")), enc_out, mtime=0)
    enc_out.close_sync()
    z = open(str(tmp_path / "cat.jpg.gz"), "rb").read()
    assert z == orc.gzip_encode(cat, 6, 0)[1] == a.GZipEncoder().encode_bytes(cat, mtime=0)
    dec_out = a.OutputFileStream(str(tmp_path / "cat.jpg"))
    assert a.GZipDecoder().decode_stream(a.InputFileStream(str(tmp_path / "cat.jpg.gz")), dec_out) is True
    dec_out.close_sync()
    assert open(str(tmp_path / "cat.jpg"), "rb").read() == cat
    # the reference's own fixture, file -> file (extract_archive_to_disk.dart:183-190)
    out = a.OutputFileStream(str(tmp_path / "test2.tar"))
    assert a.GZipDecoder().decode_stream(a.InputFileStream(os.path.join(G, "test2.tar.gz")), out) is True
    out.close_sync()
    assert open(str(tmp_path / "test2.tar"), "rb").read() == rd("test2.tar")
    # zlib_test.dart:35-58: memory -> OutputFileStream, then InputFileStream -> memory
    buf = bytes((i * 7) & 0xFF for i in range(10000))
    zo = a.OutputFileStream(str(tmp_path / "zlib_stream.zlib"))
    a.ZLibEncoder().encode_stream(a.InputMemoryStream(buf), zo)
    zo.close_sync()
    assert open(str(tmp_path / "zlib_stream.zlib"), "rb").read() == orc.zlib_encode(buf)[1]
    mo = a.OutputMemoryStream()
    assert a.ZLibDecoder().decode_stream(a.InputFileStream(str(tmp_path / "zlib_stream.zlib")), mo) is True
    assert mo.get_bytes() == buf
    # zlib file -> file, both directions, with parameters; raw deflate
    for level, wbits, raw in ((6, 15, False), (9, 12, False), (1, 15, True), (0, 15, False)):
        src = str(tmp_path / "plain.bin")
        open(src, "wb").write(cat)
        zo = a.OutputFileStream(str(tmp_path / "p.z"))
        a.ZLibEncoder().encode_stream(a.InputFileStream(src), zo, level=level, window_bits=wbits, raw=raw)
        zo.close_sync()
        z = open(str(tmp_path / "p.z"), "rb").read()
        assert z == orc.zlib_encode(cat, level, wbits, raw)[1], (level, wbits, raw)
        ok, got = _file_decode(a, a.ZLibDecoder(), tmp_path, z, name="p2.z", verify=True, raw=raw)
        assert ok is True and got == cat
    # Inflate.stream on an InputFileStream leaves the stream where the reference does (inflate.dart:337-340)
    raw = zlib.compress(buf)[2:-4]
    open(str(tmp_path / "raw.bin"), "wb").write(raw + b"TRAILING")
    fs = a.InputFileStream(str(tmp_path / "raw.bin"))
    assert a.Inflate.stream(fs).get_bytes() == buf and fs.position == len(raw)


def test_bzip2_files(a, tmp_path):
    # extract_archive_to_disk.dart:191-202: .tar.bz2 -> temp.tar through file streams
    out = a.OutputFileStream(str(tmp_path / "t.tar"))
    assert a.BZip2Decoder().decode_stream(a.InputFileStream(os.path.join(G, "test2.tar.bz2")), out, verify=True) is True
    out.close_sync()
    assert open(str(tmp_path / "t.tar"), "rb").read() == rd("test2.tar")
    # encodeStream file -> file == the oracle's bytes; and back
    from archive_b200 import synth
    text = synth.text(150000, stream=43).tobytes()
    src = str(tmp_path / "text.bin")
    open(src, "wb").write(text)
    eo = a.OutputFileStream(str(tmp_path / "text.bz2"))
    assert a.BZip2Encoder().encode_stream(a.InputFileStream(src), eo) is True
    eo.close_sync()
    z = open(str(tmp_path / "text.bz2"), "rb").read()
    assert z == orc.bzip2_encode(text)[1]
    ok, got = _file_decode(a, a.BZip2Decoder(), tmp_path, z, name="t.bz2", verify=True)
    assert ok is True and got == text
    # two blocks, the second one damaged: the first block's bytes are in the file, decodeStream says false
    big = synth.text(230000, stream=44).tobytes()
    z1 = bytearray(bz2.compress(big, 1))  # 100 kB blocks -> 3 blocks
    z1[len(z1) * 2 // 3] ^= 0x10
    z1 = bytes(z1)
    for verify in (False, True):
        ost, oout = orc.bzip2_decode(z1, verify=verify)
        ok, got = _file_decode(a, a.BZip2Decoder(), tmp_path, z1, name="bad.bz2", verify=verify)
        want = {orc.OK: True, orc.FALSE: False, orc.THROW: "throw"}[ost]
        assert ok == want and (want == "throw" or got == oout), (verify, ok, want, len(got), len(oout))


def test_file_codec_argument_errors(a, tmp_path):
    import ctypes as C
    from archive_b200 import _ffi
    L = _ffi.ensure_init()
    used, got = C.c_uint64(1), C.c_uint64(1)
    missing = os.fsencode(str(tmp_path / "does_not_exist"))
    outp = os.fsencode(str(tmp_path / "o"))
    assert L.b200z_file_codec(_ffi.FILE_GZIP_DECODE, missing, 0, 10, outp, 0, 0, 0, 0, C.byref(used), C.byref(got)) == _ffi.E_ARG
    assert got.value == 0 and b"cannot open" in L.b200z_last_error()
    assert L.b200z_file_codec(99, missing, 0, 10, outp, 0, 0, 0, 0, C.byref(used), C.byref(got)) == _ffi.E_ARG
    src = str(tmp_path / "in.gz")
    open(src, "wb").write(rd("a.txt.gz"))
    # a range past the end of the file is clamped (input_file_stream.dart:196-207): nothing to decode, nothing written
    assert L.b200z_file_codec(_ffi.FILE_GZIP_DECODE, os.fsencode(src), 10**9, 5, outp, 0, 0, 0, 0, C.byref(used), C.byref(got)) == 0
    assert (used.value, got.value) == (0, 0)
    assert L.b200z_file_codec(_ffi.FILE_GZIP_DECODE, os.fsencode(src), 0, 2**64 - 1, outp, 3, 0, 0, 0, C.byref(used), C.byref(got)) == 0
    assert used.value == os.path.getsize(src)
    assert open(str(tmp_path / "o"), "rb").read()[3:] == orc.gzip_decode(rd("a.txt.gz"))[1]


# ---------------------------------------------------------------------------------------------
# b200z_deflate_batch: ZipEncoder's members in one call, several in flight
# ---------------------------------------------------------------------------------------------
def test_deflate_batch_equals_one_by_one(a, monkeypatch):
    import random
    from archive_b200 import synth
    from archive_b200.zip import deflate_batch
    rng = random.Random(77)
    text = synth.text(300000, stream=45).tobytes()
    items = [b"", b"x", text[:70000], bytes(rng.randrange(256) for _ in range(3000)), text[70000:70000 + 33333],
             b"ab" * 5000, text[110000:300000], bytes(1000), text[5:4000]]
    for lanes in ("1", "8"):  # (on the CPU emulation the lanes' host threads run, their kernels one at a time)
        monkeypatch.setenv("B200Z_DEFLATE_LANES", lanes)
        for level in (6, 1, 0, 9):
            got = deflate_batch(items, level)
            for it, (payload, crc) in zip(items, got):
                d = a.Deflate(it, level=level)
                assert payload == d.get_bytes() and crc == d.crc32 == zlib.crc32(it), (lanes, level, len(it))
            if level in (6, 1):
                for it, (payload, _) in zip(items, got):
                    assert payload == orc.deflate(it, level)[1]
    # levels 1-3: the tokens of a whole group of members come from one k_defl_fast_batch launch; a small token store
    # splits the batch into several groups
    monkeypatch.setenv("B200Z_DEFLATE_TOK_MB", "1")
    for level in (1, 2, 3):
        got = deflate_batch(items, level)
        for it, (payload, crc) in zip(items, got):
            assert payload == orc.deflate(it, level)[1] and crc == zlib.crc32(it), (level, len(it))
    monkeypatch.delenv("B200Z_DEFLATE_TOK_MB")
    assert deflate_batch([], 6) == []
    with pytest.raises(a.B200ZError):
        deflate_batch([b"abc"], 11)


def test_zip_encoder_batch_mode(a):
    import time
    from archive_b200.zip import Archive, ArchiveFile, ZipDecoder, ZipEncoder
    from archive_b200 import synth
    arc = Archive()
    t0 = int(time.mktime((2024, 5, 17, 13, 37, 42, 0, 0, -1)))
    text = synth.text(200000, stream=46).tobytes()
    for i, (name, body, comp) in enumerate([("a.txt", text[:90000], None), ("b.bin", bytes(range(256)) * 9, "none"),
                                            ("c.src", b"bzip me " * 500, "bzip2"), ("empty", b"", None),
                                            ("d.txt", text[90000:], "deflate")]):
        f = ArchiveFile(name, len(body))
        f.content, f.compression, f.last_mod_time, f.mode = body, comp, t0 + 2 * i, 0o100644
        arc.add(f)
    for level in (1, 6):
        one = ZipEncoder().encode_bytes(arc, level=level)
        assert ZipEncoder(batch=True).encode_bytes(arc, level=level) == one
    back = ZipDecoder().decode_bytes(one)
    assert [f.content for f in back.files] == [f.content for f in arc.files]


def test_gzip_chunk_pipeline_many_chunks(a, monkeypatch):
    """The end-to-end chunk pipeline of b200z_gzip_decode (copy in / kernels / copy out on three streams, first chunks ramped)
    with chunks small enough that a test-sized input makes a dozen of them: same bytes as the oracle, ramp on or off."""
    from archive_b200 import synth
    text = synth.text(64 * 16384, stream=47).tobytes()
    blob = b"".join(_members(text, 16384))
    assert orc.gzip_decode(blob) == (orc.OK, text)
    monkeypatch.setenv("B200Z_GZIP_CHUNK_KB", "16")
    assert a.GZipDecoder().decode_bytes(blob) == text
    monkeypatch.setenv("B200Z_GZIP_RAMP", "0")
    assert a.GZipDecoder().decode_bytes(blob) == text
    monkeypatch.delenv("B200Z_GZIP_RAMP")
    monkeypatch.setenv("B200Z_GZIP_PIPED_WALK", "1")  # the optional form that walks the members between the launches
    assert a.GZipDecoder().decode_bytes(blob) == text
    cut_piped = _mem_decode(a, a.GZipDecoder(), blob[:-9])  # the last member ends short: same verdict, same bytes
    monkeypatch.delenv("B200Z_GZIP_PIPED_WALK")
    assert cut_piped == _mem_decode(a, a.GZipDecoder(), blob[:-9])
    monkeypatch.setenv("B200Z_GZIP_CHUNK_KB", "200")
    tail = blob + b"\x1f\x8b\x08\x00" + bytes(30)  # garbage behind the last member: the verdict comes from the slow path
    got = _mem_decode(a, a.GZipDecoder(), tail)
    monkeypatch.delenv("B200Z_GZIP_CHUNK_KB")
    assert got == _mem_decode(a, a.GZipDecoder(), tail) and got[1][:len(text)] == text


def test_gzip_file_fuzz_vs_memory(a, tmp_path, monkeypatch):
    """Seeded mixes of hinted / hint-free / lying / damaged members and segment sizes: the file always holds what the memory
    entry point returns, with the same verdict."""
    import random
    from archive_b200 import synth
    rng = random.Random(0xF11E)
    text = synth.text(48 * 8192, stream=48).tobytes()
    hinted = _members(text, 8192)
    plain = _members(text, 8192, hint=False)
    for trial in range(10):
        ms = []
        for i in range(rng.randrange(20, 48)):
            r = rng.random()
            if r < 0.80:
                ms.append(hinted[i])
            elif r < 0.90:
                ms.append(plain[i])
            elif r < 0.95:
                m = bytearray(hinted[i])
                m[-4:] = struct.pack("<I", rng.randrange(1, 9000))  # ISIZE lies
                ms.append(bytes(m))
            else:
                m = bytearray(hinted[i])
                m[rng.randrange(30, len(m) - 8)] ^= 1 << rng.randrange(8)  # damaged payload
                ms.append(bytes(m))
        blob = b"".join(ms)
        if rng.random() < 0.3:
            blob = blob[:rng.randrange(len(blob) // 2, len(blob))]
        monkeypatch.setenv("B200Z_FILE_SEG_KB", str(rng.choice([64, 64, 96, 128])))
        monkeypatch.setenv("B200Z_FILE_THREADS", str(rng.choice([1, 2, 8])))
        mem = _mem_decode(a, a.GZipDecoder(), blob)
        got = _file_decode(a, a.GZipDecoder(), tmp_path, blob, name=f"fuzz{trial}.gz", prefix=b"P" * rng.randrange(0, 5))
        assert got[0] == mem[0] and got[1] == mem[1], (trial, got[0], mem[0], len(got[1]), len(mem[1]), _stats())


def test_extract_file_to_disk(a, tmp_path):
    """extractFileToDisk (io/extract_archive_to_disk.dart:160-267) for what this package decodes: .zip archives unpacked to
    files, and the GZip / BZip2 stage of a compressed tar done file -> file by the library."""
    import shutil
    Z = os.path.join(G, "zip")
    out = str(tmp_path / "unz")
    written = a.extract_file_to_disk(os.path.join(Z, "test.zip"), out)
    arc = a.ZipDecoder().decode_bytes(open(os.path.join(Z, "test.zip"), "rb").read())
    files = [f for f in arc.files if f.is_file]
    assert sorted(os.path.relpath(p, out) for p in written) == sorted(os.path.normpath(f.name) for f in files)
    for f in files:
        assert open(os.path.join(out, os.path.normpath(f.name)), "rb").read() == f.content
    # symbolic links: the reference's fixture points OUTSIDE the output directory ("../target") and is skipped
    # (_isValidSymLink :24-38); a link that stays inside becomes a link
    out2 = str(tmp_path / "sym")
    assert a.extract_file_to_disk(os.path.join(Z, "symlink.zip"), out2) == []
    assert not os.path.lexists(os.path.join(out2, "symlink"))
    import io
    import zipfile
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.writestr("data/real.txt", b"real " * 100)
        zi = zipfile.ZipInfo("data/alias")
        zi.create_system, zi.external_attr = 3, 0o120777 << 16
        zf.writestr(zi, "real.txt")
    p = str(tmp_path / "links.zip")
    open(p, "wb").write(buf.getvalue())
    out2b = str(tmp_path / "sym2")
    a.extract_file_to_disk(p, out2b)
    assert os.readlink(os.path.join(out2b, "data/alias")) == "real.txt"
    assert open(os.path.join(out2b, "data/alias"), "rb").read() == b"real " * 100
    # bzip2 members inside a zip
    out3 = str(tmp_path / "bz")
    a.extract_file_to_disk(os.path.join(G, "zip_bzip2.zip"), out3)
    for f in a.ZipDecoder().decode_bytes(rd("zip_bzip2.zip")).files:
        if f.is_file:
            assert open(os.path.join(out3, os.path.normpath(f.name)), "rb").read() == f.content
    # the decompression stage of .tar.gz / .tgz / .tar.bz2 / .tbz, file -> file
    for src, name in (("test2.tar.gz", "test2.tar.gz"), ("test2.tar.gz", "other.TGZ"), ("test2.tar.bz2", "test2.tar.bz2"),
                      ("test2.tar.bz2", "x.tbz")):
        p = str(tmp_path / name)
        shutil.copy(os.path.join(G, src), p)
        d = str(tmp_path / ("out_" + name))
        (tar_path,) = a.extract_file_to_disk(p, d)
        assert os.path.dirname(tar_path) == d and tar_path.endswith(".tar")
        assert open(tar_path, "rb").read() == rd("test2.tar")


def test_zip_file_encoder_on_the_device(tmp_path):
    import os
    from archive_b200 import ZipFileEncoder, extract_file_to_disk
    from archive_b200 import synth
    root = tmp_path / "tree"
    (root / "d").mkdir(parents=True)
    text = synth.text(120000, stream=64).tobytes()
    files = {"one.txt": text[:50000], "d/two.txt": text[50000:], "d/three.bin": bytes(range(256)) * 20}
    for rel, body in files.items():
        (root / rel).write_bytes(body)
    sizes = {}
    for batch in (False, True):
        zp = str(tmp_path / f"t{int(batch)}.zip")
        ZipFileEncoder(batch=batch).zip_directory(str(root), filename=zp, level=6, modified=86400 * 365 * 30)
        sizes[batch] = open(zp, "rb").read()
    assert sizes[False] == sizes[True]
    out = str(tmp_path / "back")
    extract_file_to_disk(str(tmp_path / "t1.zip"), out)
    for rel, body in files.items():
        assert open(os.path.join(out, rel), "rb").read() == body
