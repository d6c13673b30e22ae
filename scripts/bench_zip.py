"""Side measurement, config 5 style: a .zip of N members x 4 MiB text, method 8, every member decoded by one
b200z_zip_extract call (host buffers in, host buffers out)."""
import ctypes as C, io, json, os, sys, time, zipfile, zlib
sys.path.insert(0, '.')
import numpy as np
from archive_b200 import _ffi, synth
from concurrent.futures import ThreadPoolExecutor
L = _ffi.ensure_init()
n, size = int(os.environ.get('ZIP_MEMBERS', 256)), int(os.environ.get('ZIP_MEMBER_MIB', 4)) << 20
txt = synth.text(n * size, stream=700)
flush = os.environ.get('ZIP_FLUSH', '1') == '1'
def comp(i):
    b = txt[i * size:(i + 1) * size].tobytes()
    z = synth.deflate_raw_flushed(b, 65536) if flush else synth.deflate_raw(b)
    return z, zlib.crc32(b)
with ThreadPoolExecutor(32) as ex: parts = list(ex.map(comp, range(n)))
data = synth.zip_from_deflated([(f"member{i:04d}.txt", z, crc, size) for i, (z, crc) in enumerate(parts)])
assert zipfile.ZipFile(io.BytesIO(data)).read("member0003.txt") == txt[3 * size:4 * size].tobytes()
zl = len(data); h_in = L.b200z_host_alloc(zl); C.memmove(h_in, data, zl)
cnt = C.c_size_t(0); ents = (_ffi.ZipEntry * n)()
assert L.b200z_zip_list(h_in, zl, ents, n, C.byref(cnt)) == 0 and cnt.value == n
tot = n * size; h_out = L.b200z_host_alloc(tot)
off = (C.c_uint64 * n)(*[i * size for i in range(n)]); room = (C.c_uint64 * n)(*[size] * n); ol = (C.c_uint64 * n)(); st = (C.c_int32 * n)()
times = []
for it in range(4):
    t0 = time.perf_counter(); rc = L.b200z_zip_extract(h_in, zl, ents, n, h_out, tot, off, room, ol, st, 0); times.append(time.perf_counter() - t0)
    assert rc == 0, _ffi.last_error()
assert all(s == 0 for s in st) and all(o == size for o in ol)
out = np.ctypeslib.as_array((C.c_uint8 * tot).from_address(h_out))
assert all(zlib.crc32(out[i * size:(i + 1) * size].tobytes()) == parts[i][1] for i in range(0, n, 17))
print(json.dumps({"workload": f"zip {n} x {size >> 20} MiB deflate-6 members", "zip_bytes": zl, "out_bytes": tot, "times_s": [round(t, 4) for t in times],
                  "full_flush_every_64KiB": flush, "GBps_out_e2e": round(tot / min(times[1:]) / 1e9, 2)}))
