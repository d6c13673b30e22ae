"""Host mirror of the reference's ZIP reader for the B200 path: `ZipDecoder().decode_bytes(data)` ->
`Archive` of `ArchiveFile`s, as lib/src/codecs/zip_decoder.dart:18-81 builds it.  The directory is parsed by
b200z_zip_list (ZipDirectory / ZipFileHeader / ZipFile.read), and -- this is the point of the batching -- ALL members are
decompressed by ONE b200z_zip_extract call (every deflate member is a unit of the same inflate batch) instead of one
Inflate per member on first access (zip_file.dart:201-248)."""
from __future__ import annotations

import ctypes as C

from . import _ffi

U_DONE, U_EOS, U_STOP, U_NOSPC = 0, 1, -1, -2
ZIP_ENCRYPTED, ZIP_TOO_LARGE = -20, -21
COMPRESSION = {0: "none", 8: "deflate", 12: "bzip2"}  # zip_file.dart:36-40; anything else is read as "none" (:83)


def _name(raw: bytes) -> str:
    try:  # InputStream.readString: UTF-8, falling back to one char per byte (input_stream.dart:140-149)
        return raw.decode("utf-8")
    except UnicodeDecodeError:
        return raw.decode("latin-1")


class ArchiveFile:
    """archive_file.dart:14-130 (the fields ZipDecoder fills)."""

    def __init__(self, name: str, size: int, is_file: bool = True):
        self.name, self.size, self.is_file = name, size, is_file
        self.mode = 0o644
        self.crc32 = None
        self.last_mod_time = 0
        self.compression = None
        self.symbolic_link = None
        self.content = b"" if is_file else None
        self.status = U_DONE  # unit status of the member's decode (include/b200z.h)

    @property
    def is_symbolic_link(self):
        return bool(self.symbolic_link)

    def read_bytes(self):
        return self.content


class Archive:
    """archive.dart:6-60: files in directory order; a later entry with a name already present replaces the earlier."""

    def __init__(self):
        self.files, self._index = [], {}

    def find(self, name):
        i = self._index.get(name)
        return None if i is None else self.files[i]

    def add(self, f: ArchiveFile):
        i = self._index.get(f.name)
        if i is not None:
            self.files[i] = f
            return
        self._index[f.name] = len(self.files)
        self.files.append(f)

    def __iter__(self):
        return iter(self.files)

    def __len__(self):
        return len(self.files)


class ZipDecoder:
    def __init__(self, web_eos: bool = False, split_flush_points: bool = True):
        # web_eos: the pure-Dart Inflate's end-of-stream behaviour; default is what the Dart VM's ZipDecoder
        # gives (dart:io zlib): a member's last symbols are always decoded
        self.flags = (1 if web_eos else 0) | (0 if split_flush_points else 2)
        self.entries = []
        self.zip_file_comment = ""

    def list(self, data):
        L = _ffi.lib()
        addr, n, keep = _ffi.as_buffer(data)
        cnt = C.c_size_t(0)
        _ffi.check(L.b200z_zip_list(addr, n, None, 0, C.byref(cnt)))
        ents = (_ffi.ZipEntry * max(1, cnt.value))()
        _ffi.check(L.b200z_zip_list(addr, n, ents, cnt.value, C.byref(cnt)))
        self.entries = [ents[i] for i in range(cnt.value)]
        off, clen = C.c_uint64(0), C.c_uint32(0)
        _ffi.check(L.b200z_zip_comment(addr, n, C.byref(off), C.byref(clen)))
        raw = bytes(memoryview(data)[off.value:off.value + clen.value]) if clen.value else b""
        self.zip_file_comment = raw.decode("latin-1")  # readString(utf8: false) (zip_directory.dart:43)
        return ents, cnt.value

    def decode_stream(self, input, verify: bool = False, password=None) -> Archive:
        """ZipDecoder().decodeStream(input) (zip_decoder.dart:29-81): the rest of an InputMemoryStream or InputFileStream."""
        from .streams import InputFileStream
        if isinstance(input, InputFileStream):
            data = input.to_uint8_list()
            input.skip(len(data))
        else:
            data = bytes(input.buffer[input.position:])
            input.position = len(input.buffer)
        return self.decode_bytes(data, verify=verify, password=password)

    def decode_bytes(self, data, verify: bool = False, password=None) -> Archive:
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        ents, n = self.list(data)
        contents, statuses = self._extract(data, ents, n)
        archive = Archive()
        for i in range(n):
            e = ents[i]
            name = _name(data[e.name_off:e.name_off + e.name_len]) if e.has_data else ""
            is_dir = name.endswith("/") or name.endswith("\\")
            entry = archive.find(name)
            if entry is None:
                entry = ArchiveFile(name, 0, is_file=False) if is_dir else ArchiveFile(name, e.uncomp_size if e.has_data else 0)
                entry.compression = COMPRESSION.get(e.method, "none") if e.has_data else "none"
                if not is_dir:
                    entry.content, entry.status = contents[i], statuses[i]
                archive.add(entry)
            entry.mode = e.ext_attr >> 16
            if (e.version_made_by >> 8) == 3 and (entry.mode & 0xF000) == 0xA000:  # unix symlink (:58-70)
                try:
                    entry.symbolic_link = contents[i].decode("utf-8")
                except UnicodeDecodeError:
                    pass
            entry.crc32 = e.crc32
            entry.last_mod_time = (e.mod_date << 16) | e.mod_time
        return archive

    def _extract(self, data, ents, n):
        if n == 0:
            return [], []
        L = _ffi.ensure_init()
        addr, zlen, keep = _ffi.as_buffer(data)
        room = [max(int(ents[i].hint_uncomp_size), int(ents[i].uncomp_size), 1) if ents[i].has_data else 0 for i in range(n)]
        contents, statuses = [b""] * n, [U_DONE] * n
        todo = list(range(n))
        while todo:
            m = len(todo)
            sub = (_ffi.ZipEntry * m)(*[ents[i] for i in todo])
            off, tot = [], 0
            for i in todo:
                off.append(tot)
                tot += (room[i] + 63) & ~63
            out = (C.c_uint8 * max(tot, 1))()
            a64 = lambda l: (C.c_uint64 * m)(*l)
            out_len, st = (C.c_uint64 * m)(), (C.c_int32 * m)()
            _ffi.check(L.b200z_zip_extract(addr, zlen, sub, m, C.addressof(out), max(tot, 1), a64(off),
                                           a64([room[i] for i in todo]), out_len, st, self.flags))
            again = []
            for k, i in enumerate(todo):
                statuses[i] = st[k]
                if st[k] == U_NOSPC and room[i] < (1 << 32) - 64:  # the size fields lied: the data decide (grow and retry)
                    room[i] = min(max(room[i] * 4, int(out_len[k]), int(ents[i].comp_size) * 4), (1 << 32) - 64)
                    again.append(i)
                    continue
                contents[i] = C.string_at(C.addressof(out) + off[k], min(int(out_len[k]), room[i]))
            todo = again
        return contents, statuses


# ---------------------------------------------------------------------------------------------
# ZipEncoder (lib/src/codecs/zip_encoder.dart:66-583)
# ---------------------------------------------------------------------------------------------
def _dos_time(t):  # _getTime :33-40
    t1 = ((t.tm_min & 0x7) << 5) | (t.tm_sec // 2)
    t2 = (t.tm_hour << 3) | (t.tm_min >> 3)
    return ((t2 & 0xFF) << 8) | (t1 & 0xFF)


def _dos_date(t):  # _getDate :42-49
    d1 = ((t.tm_mon & 0x7) << 5) | t.tm_mday
    d2 = (((t.tm_year - 1980) & 0x7F) << 1) | (t.tm_mon >> 3)
    return ((d2 & 0xFF) << 8) | (d1 & 0xFF)


def _b200_compress(content: bytes, method: str, level: int):
    """-> (payload, crc32 of content): the member's data as the reference produces it -- raw DEFLATE through
    platformZLibEncoder.encodeStream(raw: true) (:244-249), BZip2Encoder (:250-255) or the bytes themselves -- on the device."""
    L = _ffi.ensure_init()
    addr, n, keep = _ffi.as_buffer(content)
    crc = C.c_uint32(0)
    if method == "deflate":
        cap = L.b200z_deflate_bound(n)
        out = (C.c_uint8 * cap)()
        out_len = C.c_size_t(0)
        _ffi.check(L.b200z_deflate_raw(addr, n, level, 15, C.addressof(out), cap, C.byref(out_len), C.byref(crc)))
        return C.string_at(out, out_len.value), crc.value
    _ffi.check(L.b200z_crc32(addr, n, C.byref(crc)))
    if method == "bzip2":
        cap = L.b200z_bzip2_bound(n)
        out = (C.c_uint8 * cap)()
        out_len = C.c_size_t(0)
        _ffi.check(L.b200z_bzip2_encode(addr, n, C.addressof(out), cap, C.byref(out_len)))
        return C.string_at(out, out_len.value), crc.value
    return bytes(content), crc.value


def deflate_batch(contents, level: int = 6, window_bits: int = 15):
    """Raw DEFLATE of every item of `contents` with ONE b200z_deflate_batch call (all inputs staged at once, several members in
    flight on the device) -> list of (payload, crc32).  Each payload equals Deflate(item, level:, windowBits:).getBytes()."""
    import numpy as np
    L = _ffi.ensure_init()
    n = len(contents)
    if n == 0:
        return []
    in_len = np.array([len(c) for c in contents], dtype=np.uint64)
    in_off = np.zeros(n, dtype=np.uint64)
    in_off[1:] = np.cumsum(in_len)[:-1]
    blob = b"".join(bytes(c) for c in contents)
    addr, nb, keep = _ffi.as_buffer(blob if blob else b"\0")
    out_cap = np.array([L.b200z_deflate_bound(int(x)) for x in in_len], dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(out_cap)[:-1]
    out = np.empty(int(out_cap.sum()), dtype=np.uint8)
    out_len = np.zeros(n, dtype=np.uint64)
    crc = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data
    _ffi.check(L.b200z_deflate_batch(addr, p(in_off), p(in_len), n, level, window_bits, p(out), p(out_off), p(out_cap),
                                     p(out_len), p(crc), p(status)))
    assert not status.any(), "b200z_deflate_bound is an upper bound"
    return [(out[int(out_off[i]):int(out_off[i] + out_len[i])].tobytes(), int(crc[i])) for i in range(n)]


class ZipEncoder:
    """`ZipEncoder().encode_bytes(archive, level: 1, modified:)` (zip_encoder.dart:66-121): local headers + data, central
    directory, (zip64) end records, written field by field as `_writeFile` :309-372 and `_writeCentralDirectory` :391-497 do.
    Members are compressed on the device (`compress` exists so that the CPU test tier can check the container logic with a
    stand-in).  Not mirrored: encryption, and passing already-compressed members through (`file.isCompressed`, :214-235) --
    the ArchiveFile of this package holds content, not the source archive's bytes."""

    VERSION = 20

    def __init__(self, compress=None, batch: bool = False):
        """batch=True: all deflate members go to the device in one b200z_deflate_batch call (several members in flight)
        instead of one b200z_deflate_raw call each; the archive bytes are the same."""
        self._compress = compress or _b200_compress
        self._batch = batch and compress is None

    def encode_bytes(self, archive, level: int = 1, modified=None, comment: str = "") -> bytes:
        import struct
        import time
        out = bytearray()
        files = []
        archive = list(archive)
        compress = self._compress
        def level_of(e):  # add(file, level: ...) overrides the encoder's level for that member (zip_encoder.dart:137-183)
            own = getattr(e, "compress_level", None)
            return own if own is not None else (level if level is not None else 6)

        if self._batch:
            idx = [i for i, e in enumerate(archive) if e.is_file and (e.compression or "deflate") == "deflate"]
            table = {}
            for lv in sorted({level_of(archive[i]) for i in idx}):  # one device batch per level in use
                grp = [i for i in idx if level_of(archive[i]) == lv]
                table.update(zip(grp, deflate_batch([archive[i].content or b"" for i in grp], lv)))
            at = [None]

            def compress(content, method, level_):
                return table[at[0]] if method == "deflate" else self._compress(content, method, level_)
        for pos_in_archive, entry in enumerate(archive):
            if self._batch:
                at[0] = pos_in_archive
            lm = time.localtime(modified if modified is not None else entry.last_mod_time)  # DateTime.fromMillisecondsSinceEpoch
            name = entry.name.replace("\\", "/")
            if not entry.is_file and not name.endswith("/"):
                name += "/"
            method = (entry.compression or "deflate") if entry.is_file else "deflate"
            payload, crc = b"", 0
            if entry.is_file:
                payload, crc = compress(entry.content or b"", method, level_of(entry))
            fd = dict(name=name, time=_dos_time(lm), date=_dos_date(lm), crc=crc, csize=len(payload),
                      usize=entry.size if entry.is_file else 0, method=method, mode=entry.mode, pos=len(out),
                      comment=getattr(entry, "comment", None) or "")
            files.append(fd)
            # _writeFile
            z64 = fd["csize"] > 0xFFFFFFFF or fd["usize"] > 0xFFFFFFFF
            extra = struct.pack("<BBBBQQ", 1, 0, 0x10, 0, fd["usize"], fd["csize"]) if z64 else b""
            m = {"deflate": 8, "bzip2": 12}.get(method, 0)
            nb = name.encode("utf-8")
            out += struct.pack("<IHHHHHIIIHH", 0x04034B50, self.VERSION, 0x800, m, fd["time"], fd["date"], crc,
                               0xFFFFFFFF if z64 else fd["csize"], 0xFFFFFFFF if z64 else fd["usize"], len(nb), len(extra))
            out += nb + extra + payload
        # _writeCentralDirectory
        cd_pos = len(out)
        any64 = False
        for fd in files:
            z64 = fd["csize"] > 0xFFFFFFFF or fd["usize"] > 0xFFFFFFFF or fd["pos"] > 0xFFFFFFFF
            any64 |= z64
            extra = struct.pack("<BBBBQQQ", 1, 0, 0x18, 0, fd["usize"], fd["csize"], fd["pos"]) if z64 else b""
            m = {"deflate": 8, "bzip2": 12}.get(fd["method"], 0)
            nb, cb = fd["name"].encode("utf-8"), fd["comment"].encode("utf-8")
            out += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, (0 << 8) | self.VERSION, self.VERSION, 0x800, m, fd["time"],
                               fd["date"], fd["crc"], 0xFFFFFFFF if z64 else fd["csize"], 0xFFFFFFFF if z64 else fd["usize"],
                               len(nb), len(extra), len(cb), 0, 0, (fd["mode"] << 16) & 0xFFFFFFFF,
                               0xFFFFFFFF if z64 else fd["pos"])
            out += nb + extra + cb
        cd_size = len(out) - cd_pos
        n = len(files)
        need64 = any64 or n > 0xFFFF or cd_size > 0xFFFFFFFF or cd_pos > 0xFFFFFFFF
        if need64:
            eocd64 = len(out)
            out += struct.pack("<IQHHIIQQQQ", 0x06064B50, 0x2C, 0x2D, 0x2D, 0, 0, n, n, cd_size, cd_pos)
            out += struct.pack("<IIQI", 0x07064B50, 0, eocd64, 1)
        cb = (comment or "").encode("utf-8")
        out += struct.pack("<IHHHHIIH", 0x06054B50, 0, 0xFFFF if need64 else 0, 0xFFFF if need64 else n,
                           0xFFFF if need64 else n, 0xFFFFFFFF if need64 else cd_size, 0xFFFFFFFF if need64 else cd_pos, len(cb))
        out += cb
        return bytes(out)

    encode = encode_bytes
