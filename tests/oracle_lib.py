"""ctypes access to oracle/liboracle.so -- TEST INFRASTRUCTURE (checker only, never the product path)."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OK, FALSE, THROW, RUNAWAY = 0, 1, 2, 3

_L = None


def L():
    global _L
    if _L is None:
        _L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _L.orc_crc32.restype = C.c_uint32
        _L.orc_adler32.restype = C.c_uint32
        _L.orc_set_runaway_limit(C.c_int64(1 << 24))
    return _L


def _take(out, n):
    r = C.string_at(out, n.value)
    L().orc_free(out)
    return r


def inflate(data: bytes):
    """-> (status, output, consumed)"""
    out, n, c = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_size_t()
    st = L().orc_inflate_bytes(data, C.c_size_t(len(data)), C.byref(out), C.byref(n), C.byref(c))
    return st, _take(out, n), c.value


def gzip_decode(data: bytes, verify=False, raw=False):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_gzip_decode_bytes_raw(data, C.c_size_t(len(data)), int(verify), int(raw), C.byref(out), C.byref(n))
    return st, _take(out, n)


def zlib_decode(data: bytes, verify=False, raw=False):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_zlib_decode_bytes(data, C.c_size_t(len(data)), int(verify), int(raw), C.byref(out), C.byref(n))
    return st, _take(out, n)


def crc32(data: bytes, crc=0):
    return L().orc_crc32(data, C.c_size_t(len(data)), C.c_uint32(crc))


def adler32(data: bytes, adler=1):
    return L().orc_adler32(data, C.c_size_t(len(data)), C.c_uint32(adler))


_E = None


def emul_inflate(data: bytes, cap: int = 1 << 20):
    """The CUDA kernel's per-stream decode logic compiled for the host (tests/host_emul).
    -> (unit_status, output, in_used, ntok)"""
    global _E
    if _E is None:
        _E = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libemul.so"))
    out = (C.c_uint8 * max(cap, 1))()
    ol, iu, st, nt = C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_uint32()
    _E.emul_inflate(data, len(data), out, cap, C.byref(ol), C.byref(iu), C.byref(st), C.byref(nt))
    return st.value, bytes(out[:ol.value]), iu.value, nt.value


def deflate(data: bytes, level=6, window_bits=15):
    """-> (status, compressed, crc32_of_input)"""
    out, n, crc = C.POINTER(C.c_uint8)(), C.c_size_t(), C.c_uint32()
    st = L().orc_deflate_bytes(data, C.c_size_t(len(data)), level, window_bits, C.byref(out), C.byref(n), C.byref(crc))
    return st, _take(out, n), crc.value


def zlib_encode(data: bytes, level=6, window_bits=15, raw=False):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_zlib_encode_bytes(data, C.c_size_t(len(data)), level, window_bits, int(raw), C.byref(out), C.byref(n))
    return st, _take(out, n)


def gzip_encode(data: bytes, level=6, mtime=0):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_gzip_encode_bytes(data, C.c_size_t(len(data)), level, C.c_uint32(mtime), C.byref(out), C.byref(n))
    return st, _take(out, n)


def bzip2_decode(data: bytes, verify=True):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_bzip2_decode_bytes(data, C.c_size_t(len(data)), int(verify), C.byref(out), C.byref(n))
    return st, _take(out, n)


def set_truncate_heuristic(on: bool):
    L().orc_deflate_set_truncate_heuristic(int(on))


def bzip2_encode(data: bytes):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_bzip2_encode_bytes(data, C.c_size_t(len(data)), C.byref(out), C.byref(n))
    return st, _take(out, n)


_BE = None


def emul_bzip2_encode(data: bytes):
    """The device BZip2 encoder's kernels run on the CUDA emulation (tests/host_emul/bz2enc_emul.cpp).
    -> (rc, output, stats[n_blocks, n_serial_blocks, rounds, _])"""
    global _BE
    if _BE is None:
        _BE = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libbz2enc_emul.so"))
    cap = len(data) + len(data) // 32 + 8192
    out = (C.c_uint8 * (cap + 16))()
    n = C.c_size_t()
    st = (C.c_uint32 * 4)()
    rc = _BE.emu_bzip2_encode(data, C.c_size_t(len(data)), out, C.c_size_t(cap), C.byref(n), st)
    return rc, bytes(out[:n.value]), list(st)


class ZipEntry(C.Structure):
    _fields_ = [("local_header_off", C.c_uint64), ("data_off", C.c_uint64), ("comp_size", C.c_uint64),
                ("uncomp_size", C.c_uint64), ("hint_uncomp_size", C.c_uint64), ("name_off", C.c_uint64),
                ("cd_name_off", C.c_uint64), ("name_len", C.c_uint32), ("cd_name_len", C.c_uint32), ("crc32", C.c_uint32),
                ("method", C.c_uint32), ("flags", C.c_uint32), ("mod_time", C.c_uint32), ("mod_date", C.c_uint32),
                ("ext_attr", C.c_uint32), ("version_made_by", C.c_uint32), ("has_data", C.c_uint32)]

    def astuple(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


def zip_list(data: bytes):
    """-> (status, [ZipEntry])"""
    cap = 4096
    ents = (ZipEntry * cap)()
    n = C.c_size_t()
    st = L().orc_zip_list(data, C.c_size_t(len(data)), ents, C.c_size_t(cap), C.byref(n))
    return st, [ents[i] for i in range(min(n.value, cap))]


def zip_member(data: bytes, entry, web_eos=False):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_zip_member(data, C.c_size_t(len(data)), C.byref(entry), int(web_eos), C.byref(out), C.byref(n))
    return st, _take(out, n)


class ZipMemberIn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("content", C.c_char_p), ("content_len", C.c_size_t), ("method", C.c_int), ("is_file", C.c_int),
                ("mode", C.c_uint32), ("dos_time", C.c_uint32), ("dos_date", C.c_uint32), ("comment", C.c_char_p)]


def zip_encode(members, level=1, comment=""):
    """members: [(name, content, method 'none'|'deflate'|'bzip2', is_file, mode, dos_time, dos_date, comment)] -> (status, bytes)"""
    arr = (ZipMemberIn * max(1, len(members)))()
    keep = []
    for i, (name, content, method, is_file, mode, t, d, cm) in enumerate(members):
        nb, cb = name.encode(), (cm.encode() if cm else None)
        keep += [nb, cb, content]
        arr[i] = ZipMemberIn(nb, content, len(content), {"none": 0, "deflate": 1, "bzip2": 2}[method], int(is_file), mode, t, d, cb)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    st = L().orc_zip_encode(arr, C.c_size_t(len(members)), level, comment.encode(), C.byref(out), C.byref(n))
    return st, _take(out, n)


_BD = None


def emul_bzip2_decode(data: bytes, verify: bool = True, cap: int | None = None):
    """The device BZip2 decoder's kernels (magic scan, entropy decode, inverse BWT, RLE, CRC) run on the CUDA emulation
    (tests/host_emul/bz2dec_emul.cpp); the block chain is walked by the product's host logic (archive_b200/shard.py).
    -> (status, output, n_block_reports): status OK, FALSE (decodeStream returns false) or THROW (RangeError)."""
    global _BD
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from archive_b200._ffi import Bz2Block
    from archive_b200.shard import bz2_walk_chain
    if _BD is None:
        _BD = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libbz2dec_emul.so"))
    n = len(data)
    cap = cap or max(1 << 20, n * 8)
    cap_blocks = n // 8 + 64
    blocks = (Bz2Block * cap_blocks)()
    while True:
        out = (C.c_uint8 * cap)()
        out_len, nb, empty = C.c_size_t(), C.c_size_t(), C.c_int()
        rc = _BD.emu_bzip2_blocks(data, C.c_size_t(n), out, C.c_size_t(cap), C.byref(out_len), blocks,
                                  C.c_size_t(cap_blocks), C.byref(nb), C.byref(empty))
        if rc == -3 and out_len.value > cap:
            cap = out_len.value + 64
            continue
        break
    if rc:
        return {-1: FALSE, -2: THROW}[rc], b"", 0
    if empty.value:
        return 0, b"", 0
    reports, off = [], 0
    for i in range(nb.value):
        b = blocks[i]
        reports.append((b.start_bit, b.end_bit, b.out_bytes, b.crc_calc, b.crc_stored, b.status, b.flags, 0, off))
        off += b.out_bytes
    kind, chain, n_out = bz2_walk_chain(reports, n, verify, data=data)
    body = b"".join(bytes(out[r[8]:r[8] + r[2]]) for r in chain)
    assert len(body) == n_out
    return {"ok": OK, "data": FALSE, "throw": THROW}[kind], body, nb.value


def emul_bzip2_last_quirk() -> int:
    """Blocks of the last emul_bzip2_decode call that took the literal path (k_bz2_entropy_literal)."""
    return int(_BD.emu_bzip2_last_quirk()) if _BD is not None else 0


def emul_bzip2_last_fast() -> int:
    """Blocks of the last emul_bzip2_decode call that k_bz2_entropy_fast finished (the rest went to the exact kernel)."""
    return int(_BD.emu_bzip2_last_fast()) if _BD is not None else 0


_DE = None


def emul_deflate_raw(data: bytes, level: int = 6, window_bits: int = 15):
    """The device Deflate encoder (kernels + host driver of deflate_kernels.cu, levels 1..9) run on the CUDA emulation
    (tests/host_emul/deflate_emul.cpp).  -> (rc, raw stream, stats[tokens, blocks, re-speculated chunks])"""
    global _DE
    if _DE is None:
        _DE = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libdeflate_emul.so"))
        _DE.emu_deflate_bound.restype = C.c_size_t
    cap = _DE.emu_deflate_bound(C.c_size_t(len(data)))
    out = (C.c_uint8 * cap)()
    n = C.c_size_t()
    st = (C.c_uint32 * 3)()
    rc = _DE.emu_deflate_raw(data, C.c_size_t(len(data)), level, window_bits, out, C.c_size_t(cap), C.byref(n), st)
    return rc, bytes(out[:n.value]), list(st)
