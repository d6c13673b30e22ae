"""One BZip2Decoder call on 64 MiB of text (for an ncu launch list of the decode kernels)."""
import ctypes as C, sys, zlib
sys.path.insert(0, '.')
import numpy as np
from archive_b200 import _ffi, synth
L = _ffi.ensure_init()
m = 64 << 20
text = synth.text(m, stream=200)
cap = L.b200z_bzip2_bound(m); z = np.empty(cap, dtype=np.uint8); zl = C.c_size_t(0)
assert L.b200z_bzip2_encode(text.ctypes.data, m, z.ctypes.data, cap, C.byref(zl)) == 0
out = np.empty(m + 1024, dtype=np.uint8); ol = C.c_size_t(0)
for _ in range(2):
    assert L.b200z_bzip2_decode(z.ctypes.data, zl.value, 1, out.ctypes.data, m + 1024, C.byref(ol)) == 0
assert ol.value == m and zlib.crc32(out[:m].tobytes()) == zlib.crc32(text.tobytes())
print("ok", zl.value)
