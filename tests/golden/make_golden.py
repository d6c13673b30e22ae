#!/usr/bin/env python
"""Builds tests/golden/ from the reference's own test fixtures (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

Copies the small binary fixtures that pin the decode results of the hot path,
extracts the (disabled but present) git zlib vector from test/inflate_test.dart:64-179, and writes
manifest.json with the expected sizes / sha256 of the decoded bytes, cross-checked with CPython's
zlib / bz2 so that a wrong expectation cannot be committed silently.
"""
import bz2
import gzip
import hashlib
import json
import os
import re
import shutil
import zlib

REF = "/root/reference/test"
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def dart_int_list(src, name):
    m = re.search(r"final %s = Uint8List\.fromList\(<int>\[(.*?)\]\);" % name, src, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    return bytes(int(x) for x in re.findall(r"\d+", body))


def main():
    man = {}
    copies = {
        "_data/inflate/data.bin": "inflate_data.bin",
        "_data/cat.jpg.gz": "cat.jpg.gz",
        "_data/cat.jpg": "cat.jpg",
        "_data/test2.tar.gz": "test2.tar.gz",
        "_data/a.txt.gz": "a.txt.gz",
        "_data/bzip2/test.bz2": "test.bz2",
        "_data/test2.tar.bz2": "test2.tar.bz2",
        "_data/test2.tar": "test2.tar",
        "_data/zip/zip_bzip2.zip": "zip_bzip2.zip",
    }
    for src, dst in copies.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)
    rd = lambda n: open(os.path.join(HERE, n), "rb").read()

    # test/inflate_test.dart:14-20: raw deflate -> 5259 UTF-8 chars
    d = zlib.decompressobj(-15)
    out = d.decompress(rd("inflate_data.bin"))
    assert len(out.decode("utf8")) == 5259 and d.eof and not d.unused_data
    man["inflate_data.bin"] = {"kind": "raw", "size": len(out), "sha256": sha(out), "ref": "test/inflate_test.dart:14-20"}
    # test/gzip_test.dart:63-93
    for gz, plain in (("cat.jpg.gz", "cat.jpg"), ("test2.tar.gz", "test2.tar"), ("a.txt.gz", None)):
        out = gzip.decompress(rd(gz))
        if plain:
            assert out == rd(plain)
        man[gz] = {"kind": "gzip", "size": len(out), "sha256": sha(out), "ref": "test/gzip_test.dart:63-93"}
    a_txt = ("this is a test\nof the\nzip archive\nformat.\n" * 3).encode()  # test/_test_util.dart:8-20
    assert gzip.decompress(rd("a.txt.gz")) == a_txt
    # test/bzip2_test.dart:8-12 and io_test (test2.tar.bz2)
    for bz, plain in (("test.bz2", None), ("test2.tar.bz2", "test2.tar")):
        out = bz2.decompress(rd(bz))
        if plain:
            assert out == rd(plain)
        man[bz] = {"kind": "bzip2", "size": len(out), "sha256": sha(out), "ref": "test/bzip2_test.dart:8-12"}
    # disabled git vector, test/inflate_test.dart:57-60,64-179: zlib stream, "only 148 bytes consumed"
    src = open(os.path.join(REF, "inflate_test.dart")).read()
    gin, gout = dart_int_list(src, "gitInflateInput"), dart_int_list(src, "gitExpectedOutput")
    d = zlib.decompressobj()
    got = d.decompress(gin)
    assert got == gout, "git vector expectation"
    open(os.path.join(HERE, "git_inflate_input.bin"), "wb").write(gin)
    open(os.path.join(HERE, "git_expected_output.bin"), "wb").write(gout)
    man["git_inflate_input.bin"] = {"kind": "zlib-first-stream", "size": len(gout), "sha256": sha(gout),
                                    "consumed": len(gin) - len(d.unused_data), "ref": "test/inflate_test.dart:57-179"}
    # checksum known-answer tests: test/adler32_test.dart:6-24, test/crc32_test.dart:6-24
    man["kat"] = {
        "adler32": {"empty": 1, "one": 0x20002, "ten": 0xDC002E, "hundred_k": 0x96C8DE2B},
        "crc32": {"empty": 0, "one": 0xA505DF1B, "ten": 0xC5F5BE65, "hundred_k": 0x3AC67C2B},
        "inputs": {"one": "[1]", "ten": "[1,2,3,4,5,6,7,8,9,0]", "hundred_k": "ten repeated 10000 times, fed incrementally"},
    }
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(man, indent=1))


if __name__ == "__main__":
    main()
