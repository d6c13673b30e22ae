"""InputFileStream / OutputFileStream of the host-side mirror (archive_b200/streams.py) against the expectations of the
reference's own stream tests: test/io_test.dart:200-330 (group 'InputFileStream', 'InputFileStream/OutputFileStream
(files)') and test/input_file_stream_test.dart.  CPU tier: no codec runs here."""
import os

import pytest

from archive_b200 import BIG_ENDIAN, InputFileStream, OutputFileStream

G = os.path.join(os.path.dirname(__file__), "golden")
DATA = bytes(range(120))  # io_test.dart:20-33 writes bytes 0..119 to test.bin


@pytest.fixture()
def path(tmp_path):
    p = tmp_path / "test.bin"
    p.write_bytes(DATA)
    return str(p)


def test_length_and_read_byte(path):  # io_test.dart:202-213
    fs = InputFileStream(path, buffer_size=2)
    assert fs.length == len(DATA)
    assert [fs.read_byte() for _ in DATA] == list(DATA)
    assert fs.is_eos and fs.read_byte() == 0  # input_file_stream.dart:144-148: 0 at the end, no error


def test_read_bytes(path):  # io_test.dart:215-229
    fs = InputFileStream(path)
    assert fs.length == 120
    ai = 0
    while not fs.is_eos:
        bs = fs.read_bytes(40)
        assert bs.length == 40
        b = bs.to_uint8_list()
        assert b == DATA[ai:ai + 40]
        ai += len(b)
    assert ai == 120
    assert fs.read_bytes(10).length == 0


def test_position_skip_rewind_peek(path):  # io_test.dart:231-285
    fs = InputFileStream(path, buffer_size=2)
    fs.position = 50
    assert fs.read_bytes(50).to_uint8_list() == DATA[50:100]
    fs = InputFileStream(path, buffer_size=2)
    fs.skip(50)
    assert fs.read_bytes(50).to_uint8_list() == DATA[50:100]
    fs = InputFileStream(path, buffer_size=2)
    fs.skip(50)
    fs.rewind(10)
    assert fs.read_bytes(50).to_uint8_list() == DATA[40:90]
    fs = InputFileStream(path, buffer_size=2)
    b = fs.read_bytes(50).to_uint8_list()
    fs.rewind(50)
    assert [fs.read_byte() for _ in range(50)] == list(b)
    fs = InputFileStream(path, buffer_size=2)
    assert fs.peek_bytes(10).to_uint8_list() == DATA[:10] and fs.position == 0
    fs.rewind(5)
    assert fs.position == 0  # clamps (input_file_stream.dart:131-136)


def test_clone_and_words(path):  # io_test.dart:287-297, input_file_stream_test.dart:45-105
    fs = InputFileStream(path)
    sub = InputFileStream.from_file_stream(fs, position=6, length=5)
    assert sub.read_bytes(5).to_uint8_list() == DATA[6:11]
    assert fs.file_range() == (path, 0, 120)
    fs.skip(7)
    assert fs.subset(position=3, length=4).file_range() == (path, 3, 4)
    fs = InputFileStream(path, buffer_size=2)
    assert fs.read_uint16() == 0x0100 and fs.read_uint24() == 0x040302 and fs.read_uint32() == 0x08070605
    assert fs.read_uint64() == 0x100f0e0d0c0b0a09
    be = InputFileStream(path, byte_order=BIG_ENDIAN)
    assert be.read_uint16() == 0x0001 and be.read_uint32() == 0x02030405
    fs.close_sync()
    assert fs.length == 0


def test_copy_file_through_streams(tmp_path):  # io_test.dart:300-326
    src = os.path.join(G, "cat.jpg")
    inp = InputFileStream(src)
    out = OutputFileStream(str(tmp_path / "sub" / "cat2.jpg"))  # createSync(recursive: true)
    total = inp.length
    off = 0
    while not inp.is_eos:
        bs = inp.read_bytes(50)
        if off + 50 > total:
            assert bs.length == total - off
        off += bs.length
        out.write_stream(bs)
    inp.close_sync()
    assert out.length == total
    out.close_sync()
    assert open(str(tmp_path / "sub" / "cat2.jpg"), "rb").read() == open(src, "rb").read()


def test_output_buffering_words_subset(tmp_path):  # output_file_stream.dart:96-234
    p = str(tmp_path / "o.bin")
    open(p, "wb").write(b"stale content that the stream truncates")
    out = OutputFileStream(p, buffer_size=8)
    assert os.path.getsize(p) == 0
    for v in b"abc":
        out.write_byte(v)
    out.write_bytes(b"0123456789ABCDEF")  # larger than the buffer: flush, then straight to the file
    out.write_uint16(0x1234)
    out.write_uint32(0xA1B2C3D4)
    out.write_uint64(0x8000000000000001)
    want = b"abc0123456789ABCDEF" + bytes.fromhex("3412") + bytes.fromhex("d4c3b2a1") + bytes.fromhex("0100000000000080")
    assert out.length == len(want)
    assert out.subset(3, 13) == b"0123456789" and out.subset(-8) == want[-8:]
    tail = out.file_tail()
    assert tail == (p, len(want)) and os.path.getsize(p) == len(want)  # everything flushed before the library writes
    with open(p, "r+b") as f:  # what b200z_file_codec does: bytes appear behind the stream's back ...
        f.seek(len(want))
        f.write(b"XYZ")
    out.advanced(3)  # ... and the stream is told
    out.write_bytes(b"!")
    out.close_sync()
    assert open(p, "rb").read() == want + b"XYZ!"
    be = OutputFileStream(p, byte_order=BIG_ENDIAN)
    be.write_uint16(0x1234)
    be.write_uint32(0xA1B2C3D4)
    be.close_sync()
    assert open(p, "rb").read() == bytes.fromhex("1234a1b2c3d4")


# ---------------------------------------------------------------------------------------------
# extract_archive_to_disk (lib/src/io/extract_archive_to_disk.dart:19-103): host logic, no device
# ---------------------------------------------------------------------------------------------
def test_extract_archive_to_disk_layout_and_safety(tmp_path):
    from archive_b200 import extract_archive_to_disk, get_input_extension
    from archive_b200.zip import Archive, ArchiveFile
    arc = Archive()

    def add(name, body=None, link=None, mode=0o100644, is_file=True):
        f = ArchiveFile(name, len(body or b""), is_file=is_file)
        f.content, f.mode, f.symbolic_link = body, mode, link
        arc.add(f)

    add("a.txt", b"alpha")
    add("dir/sub/b.bin", bytes(range(200)), mode=0o100600)
    add("emptydir/", is_file=False, mode=0o40755)
    add("dir/link", b"sub/b.bin", link="sub/b.bin", mode=0o120777)
    add("../evil.txt", b"outside")                                # leaves the output directory: skipped (:19-22)
    add("dir/../../evil2.txt", b"outside")
    add("dir/badlink", b"/etc/passwd", link="/etc/passwd", mode=0o120777)       # absolute target: skipped (:27-30)
    add("dir/badlink2", b"../../x", link="../../x", mode=0o120777)              # target outside: skipped (:32-35)
    out = str(tmp_path / "out")
    written = extract_archive_to_disk(arc, out, buffer_size=7)
    assert sorted(os.path.relpath(p, out) for p in written) == ["a.txt", "dir/link", "dir/sub/b.bin"]
    assert open(os.path.join(out, "a.txt"), "rb").read() == b"alpha"
    assert open(os.path.join(out, "dir/sub/b.bin"), "rb").read() == bytes(range(200))
    assert os.stat(os.path.join(out, "dir/sub/b.bin")).st_mode & 0o777 == 0o600
    assert os.path.isdir(os.path.join(out, "emptydir"))
    assert os.readlink(os.path.join(out, "dir/link")) == "sub/b.bin"
    assert open(os.path.join(out, "dir/link"), "rb").read() == bytes(range(200))
    assert not os.path.exists(str(tmp_path / "evil.txt")) and not os.path.exists(str(tmp_path / "evil2.txt"))
    assert not os.path.lexists(os.path.join(out, "dir/badlink")) and not os.path.lexists(os.path.join(out, "dir/badlink2"))
    for name, ext in (("x.tar.gz", ".tar.gz"), ("X.TAR.BZ2", ".tar.bz2"), ("a.b.tgz", ".tgz"), ("q.zip", ".zip"), ("noext", "")):
        assert get_input_extension(name) == ext  # (:146-157)
    from archive_b200 import extract_file_to_disk
    with pytest.raises(ValueError):
        extract_file_to_disk(str(tmp_path / "noext"), out)
    with pytest.raises(ValueError):
        extract_file_to_disk(str(tmp_path / "thing.rar"), out)
