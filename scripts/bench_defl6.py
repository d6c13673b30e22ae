"""One Deflate(level 6) call on N MiB of text (for an ncu launch list / capture of the encoder kernels).
python scripts/bench_defl6.py [MiB] [level]"""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import numpy as np
from archive_b200 import _ffi, synth
L = _ffi.ensure_init()
m = (int(sys.argv[1]) if len(sys.argv) > 1 else 64) << 20
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
text = synth.text(m, stream=400)
out = np.empty(L.b200z_deflate_bound(m), dtype=np.uint8); ol = C.c_size_t(0)
best = 1e9
for _ in range(2):
    t0 = time.perf_counter()
    assert L.b200z_deflate_raw(text.ctypes.data, m, level, 15, out.ctypes.data, out.size, C.byref(ol), None) == 0
    best = min(best, time.perf_counter() - t0)
print("deflate level %d, %d MiB: best %.1f ms, %.3f GB/s in, ratio %.3f" % (level, m >> 20, best * 1e3, m / best / 1e9, ol.value / m))
