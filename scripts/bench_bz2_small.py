"""One BZip2Decoder call on a single 900 kB block and on 64 MiB of text, timed (host clock around the C-ABI call)."""
import ctypes as C, sys, time, zlib
sys.path.insert(0, '.')
import numpy as np
from archive_b200 import _ffi, synth
L = _ffi.ensure_init()
sizes = [int(a) for a in sys.argv[1:]] or [900000, 64 << 20]
for m in sizes:
    text = synth.text(m, stream=200)
    cap = L.b200z_bzip2_bound(m); z = np.empty(cap, dtype=np.uint8); zl = C.c_size_t(0)
    assert L.b200z_bzip2_encode(text.ctypes.data, m, z.ctypes.data, cap, C.byref(zl)) == 0
    out = np.empty(m + 1024, dtype=np.uint8); ol = C.c_size_t(0)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        assert L.b200z_bzip2_decode(z.ctypes.data, zl.value, 1, out.ctypes.data, m + 1024, C.byref(ol)) == 0
        best = min(best, time.perf_counter() - t0)
    assert ol.value == m and zlib.crc32(out[:m].tobytes()) == zlib.crc32(text.tobytes())
    st = (C.c_ulonglong * 2)()
    L.b200z_debug_bz2_blocks(st)
    print("blocks so far: fast kernel %d, exact kernel %d" % (st[0], st[1]))
    print("bzip2 decode %d bytes (%d compressed): best %.2f ms, %.3f GB/s" % (m, zl.value, best * 1e3, m / best / 1e9))
