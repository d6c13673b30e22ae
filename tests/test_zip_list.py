"""b200z_zip_list (host-side ZipDirectory / ZipFileHeader / ZipFile.read of the C ABI; needs no GPU) against the oracle
restatement (oracle/zip.c) and CPython's zipfile, on the reference's own fixtures (test/zip_test.dart:1-211) and on
damaged copies of them."""
import ctypes as C
import json
import os
import random

import pytest

import oracle_lib as orc
from archive_b200 import _ffi

Z = os.path.join(os.path.dirname(__file__), "golden", "zip")
MAN = json.load(open(os.path.join(Z, "manifest.json")))
OK, E_THROW = 0, -5


def b200_list(data):
    L = _ffi.lib()
    cnt = C.c_size_t(0)
    cap = 4096
    ents = (_ffi.ZipEntry * cap)()
    rc = L.b200z_zip_list(data, len(data), ents, cap, C.byref(cnt))
    return rc, [tuple(getattr(ents[i], f) for f, _ in _ffi.ZipEntry._fields_) for i in range(min(cnt.value, cap))]


def same_as_oracle(data, tag):
    rc, mine = b200_list(data)
    st, ref = orc.zip_list(data)
    if st == orc.THROW:
        assert rc == E_THROW, tag
        return None
    assert rc == OK and st == orc.OK, (tag, rc, st)
    assert mine == [e.astuple() for e in ref], tag
    return ref


@pytest.mark.parametrize("name", sorted(MAN))
def test_fixture_directory(name):
    data = open(os.path.join(Z, name), "rb").read()
    ref = same_as_oracle(data, name)
    want = MAN[name].get("entries")
    if want is None or name == "readme.notzip":
        return
    assert len(ref) == len(want)
    for e, w in zip(ref, want):
        assert data[e.cd_name_off:e.cd_name_off + e.cd_name_len].decode("utf-8", "replace") == w["name"] or e.flags & 0x800 == 0
        assert e.hint_uncomp_size == w["size"] and e.method == w["method"] and (e.ext_attr >> 16) == w["mode"]
        if not e.flags & 8:
            assert e.crc32 == w["crc32"]


def test_damaged_archives_agree_with_the_oracle():
    rng = random.Random(4)
    for name in sorted(MAN):
        data = open(os.path.join(Z, name), "rb").read()
        if len(data) > 60000:
            continue
        for it in range(60):
            d = bytearray(data)
            k = rng.choice(["flip", "trunc", "tail", "zero"])
            if k == "flip":
                for _ in range(rng.randint(1, 4)):
                    d[rng.randrange(len(d))] ^= 1 << rng.randrange(8)
            elif k == "trunc":
                d = d[:rng.randrange(len(d))]
            elif k == "tail":
                d += bytes(rng.randrange(256) for _ in range(rng.randint(1, 2100)))
            else:
                i = rng.randrange(len(d))
                d[i:i + rng.randint(1, 8)] = bytes(rng.randint(1, 8))
            same_as_oracle(bytes(d), f"{name}:{k}:{it}")


def test_eocd_search_quirks():
    """_findSignature (zip_directory.dart:139-182): 1024-byte chunks from the end; a signature across a chunk boundary or
    in the last 4 bytes is not seen; the LAST signature inside the scanned area wins."""
    base = open(os.path.join(Z, "test.zip"), "rb").read()
    for pad in (0, 1, 3, 4, 1000, 1018, 1019, 1020, 1021, 1022, 1023, 1024, 1025, 2047, 2048, 5000):
        same_as_oracle(base + bytes(pad), f"pad{pad}")
        same_as_oracle(base + b"PK\x05\x06" * 3 + bytes(pad), f"fake-eocd pad{pad}")
    same_as_oracle(b"", "empty")
    same_as_oracle(b"PK\x05\x06", "only-sig")
    same_as_oracle(b"PK\x05\x06" + bytes(18), "empty-archive")


TABLE = json.load(open(os.path.join(Z, "reference_table.json")))


@pytest.mark.parametrize("name", [k for k in TABLE if not k.startswith("_")])
def test_reference_table_directory(name):
    """What the reference's own 'unzip' tests expect of the directory (test/zip_test.dart:731-775): the archive comment, the
    number of file headers and their names."""
    from archive_b200.zip import ZipDecoder
    data = open(os.path.join(Z, name), "rb").read()
    dec = ZipDecoder()
    ents, n = dec.list(data)
    want = TABLE[name]
    if "Comment" in want:
        assert dec.zip_file_comment == want["Comment"]
    if "File" in want:
        assert n == len(want["File"])
        for e, h in zip(dec.entries, want["File"]):
            assert data[e.name_off:e.name_off + e.name_len].decode() == h["Name"]


def _zip64_archive(comp64=None, cd_off64=None, cd_size64=None, lho64=None):
    """One stored member 'a' holding b'hello', with a zip64 extra field in its central header and a zip64 end-of-central-
    directory record + locator, so that every 64-bit field of the format can be set to anything."""
    import struct
    name, data = b"a", b"hello"
    crc = orc.crc32(data)
    local = struct.pack("<IHHHHHIIIHH", 0x04034b50, 45, 0, 0, 0, 0x21, crc, len(data), len(data), len(name), 0) + name + data
    extra = b""
    comp32, lho32 = len(data), 0
    if comp64 is not None:
        comp32 = 0xffffffff
    if lho64 is not None:
        lho32 = 0xffffffff
    body = b""
    if comp64 is not None:
        body += struct.pack("<Q", comp64)
    if lho64 is not None:
        body += struct.pack("<Q", lho64)
    if body:
        extra = struct.pack("<HH", 1, len(body)) + body
    cd = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014b50, 0x031e, 45, 0, 0, 0, 0x21, crc, comp32, len(data), len(name), len(extra),
                     0, 0, 0, 0o100644 << 16, lho32) + name + extra
    cd_off = len(local)
    z64 = struct.pack("<IQHHIIQQQQ", 0x06064b50, 44, 45, 45, 0, 0, 1, 1, len(cd) if cd_size64 is None else cd_size64,
                      cd_off if cd_off64 is None else cd_off64)
    z64_off = cd_off + len(cd)
    loc = struct.pack("<IIQI", 0x07064b50, 0, z64_off, 1)
    eocd = struct.pack("<IHHHHIIH", 0x06054b50, 0, 0, 1, 1, len(cd), cd_off, 0)
    return local + cd + z64 + loc + eocd


def test_zip64_fields_that_wrap_or_point_outside_the_archive():
    """64-bit sizes and offsets are the archive's to choose: sums with them must not wrap, sizes are cut to
    what the archive holds exactly as readBytes / subset do, and what makes Uint8List.view throw is B200Z_E_THROW."""
    ok = _zip64_archive()
    ref = same_as_oracle(ok, "plain zip64")
    assert len(ref) == 1 and ref[0].comp_size == 5
    arch = _zip64_archive(comp64=1 << 40)
    big = same_as_oracle(arch, "compressed size beyond the archive")
    assert big[0].comp_size == len(arch) - big[0].data_off  # clipped to what is there
    for tag, kw in (("compressed size that wraps the sum", dict(comp64=0xffffffffffffffe2)),
                    ("compressed size 2^63", dict(comp64=1 << 63)),
                    ("directory offset near 2^64", dict(cd_off64=(1 << 64) - 4, cd_size64=2)),
                    ("directory offset beyond the archive", dict(cd_off64=1 << 33)),
                    ("directory size 2^64-1", dict(cd_size64=(1 << 64) - 1)),
                    ("local header offset near 2^64", dict(lho64=(1 << 64) - 2))):
        data = _zip64_archive(**kw)
        assert same_as_oracle(data, tag) is None, tag  # the reference throws (RangeError)
    # a directory size that merely reaches beyond the archive is cut to its end, and reading on from there throws
    same_as_oracle(_zip64_archive(cd_size64=1 << 40), "directory size beyond the archive")
