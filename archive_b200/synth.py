"""Synthetic "enwik-style" workloads of SURVEY.md section 8(d) / BASELINE.md section 3.

Data generation only (bench / tests / smoke input); nothing here is on the codec path.
"""
from __future__ import annotations

import os
import struct
import zlib
from concurrent.futures import ProcessPoolExecutor

import numpy as np

SEED = 0xB200
_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_MARKUP = [b"[[", b"]]", b"''", b"<ref>", b"</ref>", b"{{", b"}}", b"==", b"\n", b". ", b", ", b"|", b"&quot;"]


class _Vocab:
    def __init__(self, rng):
        n = 50_000
        self.wlen = rng.integers(2, 11, size=n).astype(np.int64)
        w = 1.0 / np.arange(1, 27) ** 0.8
        w /= w.sum()
        chars = _LETTERS[rng.choice(26, p=w, size=int(self.wlen.sum()))]
        seps = [b" "] + _MARKUP
        self.src = np.concatenate([chars, np.frombuffer(b"".join(seps), dtype=np.uint8)])
        self.woff = np.concatenate([[0], np.cumsum(self.wlen)[:-1]])
        sl = np.array([len(s) for s in seps], dtype=np.int64)
        self.sep_len = sl
        self.sep_off = len(chars) + np.concatenate([[0], np.cumsum(sl)[:-1]])
        z = 1.0 / np.arange(1, n + 1) ** 1.07
        self.cdf = np.cumsum(z / z.sum())


def text(n_bytes: int, seed: int = SEED, stream: int = 0) -> np.ndarray:
    """n_bytes of synthetic wiki-like text (uint8 array).  `stream` re-draws (not tiles) new text."""
    vocab = _Vocab(np.random.Generator(np.random.PCG64(seed)))
    rng = np.random.Generator(np.random.PCG64([seed, stream + 1]))
    out = np.empty(n_bytes, dtype=np.uint8)
    pos = 0
    chunk = 8 << 20
    while pos < n_bytes:
        want = min(chunk, n_bytes - pos)
        k = want // 5 + 16  # mean piece is > 6 bytes: always enough
        idx = np.searchsorted(vocab.cdf, rng.random(k)).clip(0, len(vocab.wlen) - 1)
        sep = np.where(rng.random(k) < 0.88, 0, 1 + rng.integers(0, 13, size=k))
        seg_len = np.empty(2 * k, dtype=np.int64)
        seg_src = np.empty(2 * k, dtype=np.int64)
        seg_len[0::2] = vocab.wlen[idx]
        seg_src[0::2] = vocab.woff[idx]
        seg_len[1::2] = vocab.sep_len[sep]
        seg_src[1::2] = vocab.sep_off[sep]
        ends = np.cumsum(seg_len)
        total = int(ends[-1])
        gather = np.repeat(seg_src - (ends - seg_len), seg_len) + np.arange(total, dtype=np.int64)
        piece = vocab.src[gather]
        take = min(want, total)
        out[pos:pos + take] = piece[:take]
        pos += take
    return out


def deflate_raw(chunk: bytes, level: int = 6, mem_level: int = 9, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    return co.compress(chunk) + co.flush()


def gzip_member(chunk: bytes, level: int = 6, hint: bool = True) -> bytes:
    """One gzip member holding `chunk` as ONE dynamic block (memLevel 9), MTIME 0, with the BGZF
    'BC' FEXTRA subfield carrying the member size.  The reference skips FEXTRA blindly
    (_gzip_decoder_web.dart:119-122), so the same bytes decode identically there."""
    body = deflate_raw(chunk, level)
    trailer = struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk) & 0xffffffff)
    if hint:
        total = 10 + 2 + 6 + len(body) + 8
        if total <= 65536:
            hdr = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1)
            return hdr + body + trailer
    return b"\x1f\x8b\x08\x00" + b"\0\0\0\0" + b"\x00\xff" + body + trailer


def _members_job(args):
    buf, unit, level, hint = args
    mv = memoryview(buf)
    return [gzip_member(bytes(mv[i:i + unit]), level, hint) for i in range(0, len(buf), unit)]


def gzip_members(data: np.ndarray, unit: int = 65536, level: int = 6, hint: bool = True, workers: int | None = None):
    """Split `data` into `unit`-byte chunks, one gzip member each.  Returns the list of members."""
    raw = data.tobytes()
    n = len(raw)
    workers = workers or min(32, os.cpu_count() or 1)
    per = max(unit, (n // (workers * 4) // unit + 1) * unit)
    jobs = [(raw[i:i + per], unit, level, hint) for i in range(0, n, per)]
    if workers == 1 or len(jobs) == 1:
        parts = [_members_job(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            parts = list(ex.map(_members_job, jobs))
    return [m for p in parts for m in p]


def gzip_header_len(member: bytes) -> int:
    """Length of the gzip header of a member produced by gzip_member()."""
    return 18 if member[3] & 4 else 10


# ---------------------------------------------------------------------------------------------
# BASELINE config 2 workload: n_units gzip members x `unit` bytes, one dynamic block each
# ---------------------------------------------------------------------------------------------
_BLOCK_UNITS = 256  # 16 MiB of text per generation job (its own re-drawn stream)


def _workload_job(args):
    seed, stream, n_units, unit, level, keep_text = args
    t = text(n_units * unit, seed, stream)
    raw = t.tobytes()
    members = [gzip_member(raw[i:i + unit], level) for i in range(0, len(raw), unit)]
    sizes = np.array([len(m) for m in members], dtype=np.int64)
    return b"".join(members), sizes, (raw if keep_text else None)


def gzip_workload(n_units: int, unit: int = 65536, seed: int = SEED, stream0: int = 0, level: int = 6,
                  workers: int | None = None, keep_text: bool = False, cache_dir: str | None = None):
    """-> dict(blob=uint8[C], member_off=int64[n+1], unit=unit, n_units=n, text=uint8[U] | None)

    Text is re-drawn (not tiled) per 16 MiB job with stream ids stream0, stream0+1, ... so the compressed
    input holds no L2-resident duplicates.  Results are cached under cache_dir when given."""
    key = f"gzwl_s{seed:x}_{stream0}_{n_units}x{unit}_l{level}"
    if cache_dir and not keep_text:
        path = os.path.join(cache_dir, key + ".npz")
        if os.path.exists(path):
            try:
                z = np.load(path)
                return dict(blob=z["blob"], member_off=z["member_off"], unit=unit, n_units=n_units, text=None)
            except Exception:
                pass
    jobs = []
    left, s = n_units, stream0
    while left > 0:
        k = min(_BLOCK_UNITS, left)
        jobs.append((seed, s, k, unit, level, keep_text))
        left -= k
        s += 1
    workers = workers or min(32, os.cpu_count() or 1)
    if workers == 1 or len(jobs) == 1:
        parts = [_workload_job(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=min(workers, len(jobs))) as ex:
            parts = list(ex.map(_workload_job, jobs))
    blob = np.frombuffer(b"".join(p[0] for p in parts), dtype=np.uint8)
    sizes = np.concatenate([p[1] for p in parts])
    member_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    txt = np.frombuffer(b"".join(p[2] for p in parts), dtype=np.uint8) if keep_text else None
    if cache_dir and not keep_text:
        try:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = os.path.join(cache_dir, key + f".{os.getpid()}.tmp.npz")
            np.savez(tmp, blob=blob, member_off=member_off)
            os.replace(tmp, os.path.join(cache_dir, key + ".npz"))
        except Exception:
            pass
    return dict(blob=blob, member_off=member_off, unit=unit, n_units=n_units, text=txt)


def deflate_raw_flushed(chunk: bytes, every: int = 65536, level: int = 6, flush=zlib.Z_FULL_FLUSH) -> bytes:
    """Raw DEFLATE with a flush point every `every` input bytes (still one valid stream, it decodes
    identically in the reference, but exposes sub-member parallelism)."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9)
    out = []
    for i in range(0, len(chunk), every):
        out.append(co.compress(chunk[i:i + every]))
        if i + every < len(chunk):
            out.append(co.flush(flush))
    out.append(co.flush())
    return b"".join(out)


def zip_from_deflated(members) -> bytes:
    """A .zip whose members are already-compressed raw DEFLATE streams: members = [(name, deflated, crc32, size)]."""
    import struct
    out, cd = bytearray(), bytearray()
    for name, z, crc, size in members:
        nb = name.encode()
        off = len(out)
        out += struct.pack("<IHHHHHIIIHH", 0x04034b50, 20, 0, 8, 0, 0x21, crc, len(z), size, len(nb), 0) + nb + z
        cd += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014b50, 0x031e, 20, 0, 8, 0, 0x21, crc, len(z), size, len(nb), 0, 0, 0, 0,
                          0o100644 << 16, off) + nb
    cd_off = len(out)
    out += cd
    out += struct.pack("<IHHHHIIH", 0x06054b50, 0, 0, len(members), len(members), len(cd), cd_off, 0)
    return bytes(out)
