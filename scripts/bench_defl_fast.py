"""Deflate levels 1-3 (k_defl_fast_batch): b200z_deflate_batch over M members x S MiB of text (ZipEncoder's default level 1,
zip_encoder.dart:87) and b200z_deflate_raw on one stream, timed around the C-ABI call (host buffers), parity of a sample
against the oracle.   python scripts/bench_defl_fast.py [members] [MiB per member] [level]"""
import ctypes as C, sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from archive_b200 import _ffi, synth
import oracle_lib as orc
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = (int(sys.argv[2]) if len(sys.argv) > 2 else 4) << 20
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L = _ffi.ensure_init()
base = synth.text(min(M, 64) * S, stream=300)
blob = np.concatenate([base] * ((M + 63) // 64))[: M * S] if M > 64 else base
in_len = np.full(M, S, dtype=np.uint64)
in_off = (np.arange(M, dtype=np.uint64) * np.uint64(S))
cap1 = L.b200z_deflate_bound(S)
out_cap = np.full(M, cap1, dtype=np.uint64)
out_off = (np.arange(M, dtype=np.uint64) * np.uint64(cap1))
out = np.empty(M * cap1, dtype=np.uint8)
out_len = np.zeros(M, dtype=np.uint64); crc = np.zeros(M, dtype=np.uint32); status = np.zeros(M, dtype=np.int32)
p = lambda a: a.ctypes.data
best = 1e9
for it in range(3):
    t0 = time.perf_counter()
    rc = L.b200z_deflate_batch(p(blob), p(in_off), p(in_len), M, level, 15, p(out), p(out_off), p(out_cap), p(out_len), p(crc), p(status))
    best = min(best, time.perf_counter() - t0)
    assert rc == 0 and not status.any(), (rc, status[:4])
for i in (0, M - 1):
    want = orc.deflate(blob[i * S:(i + 1) * S].tobytes(), level)[1]
    got = out[int(out_off[i]):int(out_off[i] + out_len[i])].tobytes()
    assert got == want and int(crc[i]) == zlib.crc32(blob[i * S:(i + 1) * S].tobytes()), i
print("deflate_batch level %d: %d members x %d MiB: best %.1f ms, %.2f GB/s in (ratio %.3f), parity of members 0 and %d vs the oracle ok"
      % (level, M, S >> 20, best * 1e3, M * S / best / 1e9, float(out_len.sum()) / (M * S), M - 1))
one = blob[: 16 * S] if M >= 16 else blob
n1 = one.size
o1 = np.empty(L.b200z_deflate_bound(n1), dtype=np.uint8); ol = C.c_size_t(0)
best = 1e9
for it in range(2):
    t0 = time.perf_counter()
    rc = L.b200z_deflate_raw(p(one), n1, level, 15, p(o1), o1.size, C.byref(ol), None)
    best = min(best, time.perf_counter() - t0)
    assert rc == 0
assert o1[:ol.value].tobytes() == orc.deflate(one.tobytes(), level)[1]
print("deflate_raw level %d: one stream of %d MiB: best %.1f ms, %.3f GB/s in, parity vs the oracle ok" % (level, n1 >> 20, best * 1e3, n1 / best / 1e9))
