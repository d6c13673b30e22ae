"""Parity survey on damaged input: seeded truncations and single-bit flips of gzip / zlib streams through the C ABI against the
oracle -- how often verdict (true / false / RangeError) and bytes agree, and which documented divergence the rest falls under.
Runs on a B200, or without one on the emulation tier:
  B200Z_EMU_TESTS=1 B200Z_LIB=tests/host_emul/libb200z_emu.so python scripts/parity_survey.py"""
import json, os, random, struct, sys, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from collections import Counter
import oracle_lib as orc
import archive_b200 as a
from archive_b200 import synth

NAMES = {orc.OK: "ok", orc.FALSE: "false", orc.THROW: "throw", orc.RUNAWAY: "runaway"}
text = synth.text(3 * 8192, stream=49).tobytes()


def run(dec, z, **kw):
    out = a.OutputMemoryStream()
    try:
        ok = dec.decode_stream(a.InputMemoryStream(z), out, **kw)
        return (orc.OK if ok else orc.FALSE), out.get_bytes()
    except a.DartRangeError:
        return orc.THROW, out.get_bytes()


def survey(name, dec, ofn, streams, **kw):
    c = Counter()
    for z in streams:
        ost, oout = ofn(z, **kw)
        st, got = run(dec, z, **kw)
        agree = st == ost and (st == orc.THROW or got == oout)
        c["agree" if agree else f"oracle {NAMES[ost]} / here {NAMES[st]}"] += 1
    return {"case": name, "n": len(streams), **c}


gz = b"".join(synth.gzip_member(text[i:i + 8192], 6, hint=(i == 0)) for i in range(0, len(text), 8192))
zb = zlib.compress(text[:9000], 6) + zlib.compress(text[9000:20000], 9) + zlib.compress(text[20000:], 1)
rng = random.Random(5)


def flips(z, n):
    out = []
    for _ in range(n):
        b = bytearray(z)
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        out.append(bytes(b))
    return out


rows = [
    survey("gzip, 3 members, truncated at every 37th byte + the last 11", a.GZipDecoder(), orc.gzip_decode,
           [gz[:c] for c in list(range(0, len(gz), 37)) + [len(gz) - k for k in range(1, 12)]]),
    survey("gzip, single-bit flips", a.GZipDecoder(), orc.gzip_decode, flips(gz, 150)),
    survey("zlib, 3 streams, truncated at every 41st byte, verify off", a.ZLibDecoder(), orc.zlib_decode,
           [zb[:c] for c in range(0, len(zb), 41)], verify=False),
    survey("zlib, 3 streams, truncated at every 41st byte, verify on", a.ZLibDecoder(), orc.zlib_decode,
           [zb[:c] for c in range(0, len(zb), 41)], verify=True),
    survey("zlib, single-bit flips, verify on", a.ZLibDecoder(), orc.zlib_decode, flips(zb, 200), verify=True),
]
for r in rows:
    print(json.dumps(r))
