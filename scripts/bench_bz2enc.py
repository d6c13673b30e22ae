"""Side measurement: BZip2 encode through the C ABI with pinned host buffers."""
import bz2, ctypes as C, json, os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from archive_b200 import _ffi, synth
L = _ffi.ensure_init()
m = int(os.environ.get('BZE_MIB', 256)) << 20
src = synth.text(m, stream=500)
h_in = L.b200z_host_alloc(m); C.memmove(h_in, src.ctypes.data, m)
cap = L.b200z_bzip2_bound(m); h_out = L.b200z_host_alloc(cap); ol = C.c_size_t(0)
times = []
for i in range(3):
    t0 = time.perf_counter(); rc = L.b200z_bzip2_encode(h_in, m, h_out, cap, C.byref(ol)); times.append(time.perf_counter() - t0)
    assert rc == 0, _ffi.last_error()
z = C.string_at(h_out, ol.value)
res = {'in_MiB': m >> 20, 'out_bytes': ol.value, 'ratio': round(m / ol.value, 3), 'times_s': [round(t, 3) for t in times],
       'GBps_in': round(m / min(times[1:]) / 1e9, 3), 'launches': L.b200z_launch_count()}
if os.environ.get('BZE_CHECK', '1') == '1':
    k = min(m, 32 << 20)
    t0 = time.time(); assert bz2.decompress(z) == src.tobytes(); res['libbz2_decompress_s'] = round(time.time() - t0, 1)
    import oracle_lib as orc
    rc = L.b200z_bzip2_encode(h_in, k, h_out, cap, C.byref(ol))
    t0 = time.time(); oz = orc.bzip2_encode(src[:k].tobytes())[1]; res['oracle_MBps_1core'] = round(k / (time.time() - t0) / 1e6, 2)
    res['identical_to_oracle_MiB'] = (C.string_at(h_out, ol.value) == oz, k >> 20)
print(json.dumps(res))
