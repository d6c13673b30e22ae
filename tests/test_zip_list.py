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
