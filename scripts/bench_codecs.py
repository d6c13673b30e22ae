"""Side measurements for the other SURVEY section-8(d) configs (not the headline metric; bench.py is): Deflate level 6 on
256 MiB (config 3) and BZip2 decode of 512 MiB (config 4), wall clock through the C ABI with pinned host buffers."""
import bz2, ctypes as C, json, os, sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from archive_b200 import _ffi, synth
L = _ffi.ensure_init()
res = {}
# ---- config 3 ----
n = int(os.environ.get('DEFL_MIB', 256)) << 20
t0 = time.time(); text = synth.text(n, stream=100); res['gen_s'] = round(time.time() - t0, 1)
h_in = L.b200z_host_alloc(n); C.memmove(h_in, text.ctypes.data, n)
cap = L.b200z_deflate_bound(n); h_out = L.b200z_host_alloc(cap)
out_len, crc = C.c_size_t(0), C.c_uint32(0)
times = []
for i in range(4):
    t0 = time.perf_counter()
    rc = L.b200z_deflate_raw(h_in, n, 6, 15, h_out, cap, C.byref(out_len), C.byref(crc))
    times.append(time.perf_counter() - t0); assert rc == 0, _ffi.last_error()
z = C.string_at(h_out, out_len.value)
assert zlib.decompress(z, -15) == text.tobytes() and crc.value == zlib.crc32(text.tobytes())
res['deflate6'] = {'in_MiB': n >> 20, 'out_bytes': out_len.value, 'ratio': round(n / out_len.value, 3),
                   'best_s': round(min(times[1:]), 4), 'GBps_in': round(n / min(times[1:]) / 1e9, 2)}
if os.environ.get('DEFL_CHECK_ORACLE', '1') == '1':
    import oracle_lib as orc
    k = min(n, 64 << 20)
    rc = L.b200z_deflate_raw(h_in, k, 6, 15, h_out, cap, C.byref(out_len), C.byref(crc))
    t0 = time.time(); oz = orc.deflate(text[:k].tobytes(), 6)[1]; res['oracle_deflate6_MBps_1core'] = round(k / (time.time() - t0) / 1e6, 1)
    res['deflate6']['identical_to_oracle_first_MiB'] = (C.string_at(h_out, out_len.value) == oz, k >> 20)
# ---- config 4 ----
m = int(os.environ.get('BZ_MIB', 512)) << 20
if m:
    t0 = time.time(); src = synth.text(m, stream=200).tobytes()
    from concurrent.futures import ThreadPoolExecutor
    # one BZh9 stream: libbz2 single-threaded (~13 MB/s)
    zb = bz2.compress(src, 9); res['bz2_compress_s'] = round(time.time() - t0, 1)
    hb = L.b200z_host_alloc(len(zb)); C.memmove(hb, zb, len(zb))
    ho = L.b200z_host_alloc(m + 1024); ol = C.c_size_t(0)
    times = []
    for i in range(4):
        t0 = time.perf_counter(); rc = L.b200z_bzip2_decode(hb, len(zb), 1, ho, m + 1024, C.byref(ol)); times.append(time.perf_counter() - t0)
        assert rc == 0, _ffi.last_error()
    assert ol.value == m and C.string_at(ho, m) == src
    res['bzip2_decode'] = {'out_MiB': m >> 20, 'in_bytes': len(zb), 'blocks': (m + 899980) // 899981, 'best_s': round(min(times[1:]), 4),
                           'GBps_out': round(m / min(times[1:]) / 1e9, 2)}
print(json.dumps(res))
