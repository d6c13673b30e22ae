"""ZipDecoder on the B200 path: every member of an archive is decoded by ONE b200z_zip_extract call.
Checked against the reference's fixtures (test/zip_test.dart:1-211,731-775; expectations from CPython's zipfile), against
the oracle restatement of ZipFile.getStream member by member, and on synthetic archives (stored / deflate / bzip2 / zip64 /
data descriptors / directory entries / size fields that lie)."""
import hashlib
import io
import json
import os
import struct
import zipfile
import zlib

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
Z = os.path.join(os.path.dirname(__file__), "golden", "zip")
MAN = json.load(open(os.path.join(Z, "manifest.json")))


@pytest.fixture(scope="module")
def a():
    import archive_b200
    return archive_b200


@pytest.mark.parametrize("name", sorted(MAN))
def test_fixture_contents(a, name):
    data = open(os.path.join(Z, name), "rb").read()
    arc = a.ZipDecoder().decode_bytes(data)
    st, ents = orc.zip_list(data)
    assert st == orc.OK
    by_name = {}
    for e in ents:  # what the reference's Archive ends up holding: first entry of a name wins (zip_decoder.dart:43-51)
        nm = data[e.name_off:e.name_off + e.name_len].decode("utf-8", "replace") if e.has_data else ""
        by_name.setdefault(nm, e)
    assert [f.name for f in arc.files] == list(by_name)
    for f in arc.files:
        e = by_name[f.name]
        if not f.is_file or e.flags & 1:
            continue
        ost, want = orc.zip_member(data, e)
        assert f.content == want, (name, f.name)
        assert f.crc32 == e.crc32 and f.last_mod_time == (e.mod_date << 16 | e.mod_time)
    want = MAN[name].get("entries")
    if want and name != "readme.notzip":
        sha = {w["name"]: w["sha256"] for w in want}
        for f in arc.files:
            if f.is_file and sha.get(f.name):
                assert hashlib.sha256(f.content).hexdigest() == sha[f.name], (name, f.name)


def test_symlink_and_modes(a):
    arc = a.ZipDecoder().decode_bytes(open(os.path.join(Z, "symlink.zip"), "rb").read())
    assert arc.files[0].is_symbolic_link and arc.files[0].symbolic_link == "../target"  # zip_decoder.dart:58-70
    arc = a.ZipDecoder().decode_bytes(open(os.path.join(Z, "unix.zip"), "rb").read())
    assert [f.mode & 0o777 for f in arc.files] == [w["mode"] & 0o777 for w in MAN["unix.zip"]["entries"]]


def _synthetic(n_members=40, seed=3, force64=False):
    from archive_b200 import synth
    rng = np.random.default_rng(seed)
    txt = synth.text(n_members * 300_000, stream=900 + seed).tobytes()
    buf = io.BytesIO()
    want = {}
    with zipfile.ZipFile(buf, "w", allowZip64=True) as z:
        for i in range(n_members):
            size = int(rng.choice([0, 1, 100, 5000, 70_000, 300_000]))
            body = txt[i * 300_000:i * 300_000 + size]
            kind = i % 5
            name = f"dir{i % 3}/member{i}.txt"
            if kind == 0:
                zi, kw = zipfile.ZipInfo(name), dict(compress_type=zipfile.ZIP_STORED)
            elif kind == 4:
                zi, kw = zipfile.ZipInfo(name), dict(compress_type=zipfile.ZIP_BZIP2)
            else:
                zi, kw = zipfile.ZipInfo(name), dict(compress_type=zipfile.ZIP_DEFLATED, compresslevel=int(rng.choice([1, 6, 9])))
            zi.external_attr = (0o100644 | (i & 0o111)) << 16
            if i % 7 == 3:  # streamed member: sizes and CRC in a data descriptor after the data
                zi.compress_type = kw["compress_type"]
                with z.open(zi, "w", force_zip64=force64) as f:
                    f.write(body)
            else:
                zi.compress_type = kw["compress_type"]
                z.writestr(zi, body, **{k: v for k, v in kw.items() if k == "compresslevel"})
            want[name] = body
        z.writestr("emptydir/", b"")
    return buf.getvalue(), want


@pytest.mark.parametrize("force64", [False, True])
def test_synthetic_archive_equals_zipfile(a, force64):
    data, want = _synthetic(force64=force64)
    arc = a.ZipDecoder().decode_bytes(data)
    got = {f.name: f for f in arc.files}
    assert set(got) == set(want) | {"emptydir/"}
    assert not got["emptydir/"].is_file
    for name, body in want.items():
        assert got[name].content == body, name
        assert got[name].size == len(body) and got[name].crc32 == zlib.crc32(body)


def test_web_end_of_stream_quirk_is_optional(a):
    """A member's deflate stream ends with the member: the pure-Dart Inflate then wants maxCodeLength more bits and drops
    the last symbols; dart:io's zlib -- what ZipDecoder uses on the VM -- does not."""
    data, want = _synthetic(n_members=25, seed=5)
    st, ents = orc.zip_list(data)
    vm = {f.name: f.content for f in a.ZipDecoder().decode_bytes(data).files}
    web = {f.name: f.content for f in a.ZipDecoder(web_eos=True).decode_bytes(data).files}
    for e in ents:
        nm = data[e.name_off:e.name_off + e.name_len].decode()
        if not e.has_data or nm.endswith("/"):
            continue
        assert vm[nm] == orc.zip_member(data, e, web_eos=False)[1] == want[nm]
        assert web[nm] == orc.zip_member(data, e, web_eos=True)[1]
    # (with zlib-made streams the two readings rarely differ: the end-of-block code is as long as the longest code, so
    # enough bits are left when the last literal is read; hand-made streams that do differ: tests/test_inflate_gpu.py)


def test_size_fields_that_lie(a):
    """The central directory's sizes are hints: content comes from the data (zip_file.dart:201-248)."""
    data, want = _synthetic(n_members=12, seed=7)
    d = bytearray(data)
    pos = 0
    while True:  # zero every uncompressed-size field of the central directory
        pos = d.find(b"PK\x01\x02", pos)
        if pos < 0:
            break
        struct.pack_into("<I", d, pos + 24, 0)
        pos += 46
    arc = a.ZipDecoder().decode_bytes(bytes(d))
    for f in arc.files:
        if f.is_file:
            assert f.content == want[f.name]


def test_large_members_one_batch(a):
    from archive_b200 import synth
    n, size = 48, 4 << 20
    txt = synth.text(n * size, stream=950)
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w") as z:
        for i in range(n):
            z.writestr(f"m{i}", txt[i * size:(i + 1) * size].tobytes(), compress_type=zipfile.ZIP_DEFLATED, compresslevel=6)
    arc = a.ZipDecoder().decode_bytes(buf.getvalue())
    for i, f in enumerate(arc.files):
        assert zlib.crc32(f.content) == zlib.crc32(txt[i * size:(i + 1) * size].tobytes())


def test_extraction_in_chunks_equals_one_batch(a, monkeypatch):
    """b200z_zip_extract decodes a large archive in chunks of units and copies a finished chunk's bytes (with the stored
    members that lie between its units) to the host while the next chunk is decoded (B200Z_ZIP_CHUNKS; default: 8 chunks
    from 512 MiB of output on).  Forced on a small archive with deflated, stored and empty members in between: the same
    contents as one batch."""
    from archive_b200 import synth
    txt = synth.text(14 * 300_000, stream=955).tobytes()
    buf = io.BytesIO()
    want = {}
    with zipfile.ZipFile(buf, "w") as z:
        for i in range(14):
            body = txt[i * 300_000:(i + 1) * 300_000] if i % 5 != 4 else b""
            kind = zipfile.ZIP_STORED if i % 3 == 1 else zipfile.ZIP_DEFLATED
            z.writestr(f"m{i}", body, compress_type=kind, compresslevel=6)
            want[f"m{i}"] = body
    for chunks in ("1", "3", "5", "64"):
        monkeypatch.setenv("B200Z_ZIP_CHUNKS", chunks)
        arc = a.ZipDecoder().decode_bytes(buf.getvalue())
        assert {f.name: f.content for f in arc.files} == want, chunks


def test_flush_points_split_members(a):
    """Members written with Z_FULL_FLUSH points are decoded piece by piece (proven by a sizing pass); Z_SYNC_FLUSH points
    look the same but keep the window, so those members must come out right as well (decoded whole), and a flush-point
    pattern inside stored data must not fool the splitter."""
    import zipfile
    from archive_b200 import synth
    txt = synth.text(6 * (3 << 20), stream=960).tobytes()
    parts, members = [], []
    for i in range(6):
        body = txt[i * (3 << 20):(i + 1) * (3 << 20)]
        if i == 4:
            body = (b"\x00\x00\xff\xff" * 1000 + body[:200000]) * 3  # marker bytes in the data itself
        z = synth.deflate_raw_flushed(body, every=65536 if i % 2 == 0 else 100_000,
                                      flush=zlib.Z_SYNC_FLUSH if i == 3 else zlib.Z_FULL_FLUSH, level=0 if i == 4 else 6)
        members.append((f"m{i}", z, zlib.crc32(body), len(body)))
        parts.append(body)
    data = synth.zip_from_deflated(members)
    assert zipfile.ZipFile(io.BytesIO(data)).read("m1") == parts[1]
    for split in (True, False):
        arc = a.ZipDecoder(split_flush_points=split).decode_bytes(data)
        assert [f.content for f in arc.files] == parts, split


@pytest.mark.parametrize("name", [k for k in json.load(open(os.path.join(Z, "reference_table.json"))) if not k.startswith("_")])
def test_reference_table_contents(a, name):
    """The expectations of the reference's own 'unzip' tests (test/zip_test.dart:10-211 checked by :731-775): content,
    verifyCrc32, isFile, symbolic links."""
    want = json.load(open(os.path.join(Z, "reference_table.json")))[name]
    data = open(os.path.join(Z, name), "rb").read()
    arc = a.ZipDecoder().decode_bytes(data)
    for h in want.get("File", []):
        f = arc.find(h["Name"])
        assert f is not None, h["Name"]
        if "Content" in h and f.is_file:
            assert f.content == h["Content"].encode("latin-1")
        if "File" in h:
            assert f.content == open(os.path.join(Z, h["File"]), "rb").read()
        if h.get("VerifyChecksum"):
            assert zlib.crc32(f.content) == f.crc32  # ZipFile.verifyCrc32 (zip_file.dart:151-155)
        if "isFile" in h:
            assert f.is_file == h["isFile"]
        if h.get("isSymbolicLink"):
            assert f.is_symbolic_link and f.symbolic_link == h["Content"]
