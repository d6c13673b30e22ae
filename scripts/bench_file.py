"""Side measurement of the file-stream path: BASELINE config 2's gzip members written to a file, decoded
file -> file by b200z_file_codec (pinned segments, threaded pread/pwrite, read / decode / write overlapped), next to the
memory entry point on the same bytes.  FILE_DIR picks the file system (default: /dev/shm, i.e. page-cache speed);
FILE_UNITS the number of 64 KiB members; B200Z_FILE_SEG_KB / B200Z_FILE_THREADS are the knobs of the path itself.
Also times ZipEncoder's members one by one against b200z_deflate_batch."""
import ctypes as C, hashlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from archive_b200 import _ffi, synth
from archive_b200.zip import deflate_batch
import archive_b200 as a

L = _ffi.ensure_init()
units = int(os.environ.get('FILE_UNITS', 16384))
d = os.environ.get('FILE_DIR', '/dev/shm')
wl = synth.gzip_workload(units)
blob = wl["blob"] if isinstance(wl, dict) else wl[0]
blob = bytes(blob) if not isinstance(blob, (bytes, bytearray)) else blob
src, dst = os.path.join(d, 'b200z_bench.gz'), os.path.join(d, 'b200z_bench.out')
open(src, 'wb').write(blob)
out_bytes = units * 65536
res = {"workload": f"{units} gzip members x 64 KiB, file -> file in {d}", "in_bytes": len(blob), "out_bytes": out_bytes}
# memory entry point with pinned buffers (what bench.py's e2e measures)
h_in, h_out = L.b200z_host_alloc(len(blob)), L.b200z_host_alloc(out_bytes + 64)
C.memmove(h_in, blob, len(blob))
n = C.c_size_t(0)
ts = []
for _ in range(4):
    t0 = time.perf_counter(); rc = L.b200z_gzip_decode(h_in, len(blob), 0, h_out, out_bytes + 64, C.byref(n)); ts.append(time.perf_counter() - t0)
    assert rc == 0 and n.value == out_bytes
want = hashlib.sha256(C.string_at(h_out, out_bytes)).hexdigest()
res["memory_GBps_out"] = round(out_bytes / min(ts[1:]) / 1e9, 2)
for seg_kb in os.environ.get('SEG_SWEEP_KB', '65536,131072,262144,1048576').split(','):
    os.environ['B200Z_FILE_SEG_KB'] = seg_kb
    ts = []
    for _ in range(4):
        if os.path.exists(dst):
            os.unlink(dst)
        used, got = C.c_uint64(0), C.c_uint64(0)
        t0 = time.perf_counter()
        rc = L.b200z_file_codec(_ffi.FILE_GZIP_DECODE, os.fsencode(src), 0, 2**64 - 1, os.fsencode(dst), 0, 0, 0, 0, C.byref(used), C.byref(got))
        ts.append(time.perf_counter() - t0)
        assert rc == 0 and got.value == out_bytes, (rc, got.value, _ffi.last_error())
    seg, whole = C.c_uint32(0), C.c_uint32(0)
    L.b200z_file_last_stats(C.byref(seg), C.byref(whole))
    assert hashlib.sha256(open(dst, 'rb').read()).hexdigest() == want
    res[f"file_seg_{int(seg_kb) >> 10}MiB"] = {"GBps_out": round(out_bytes / min(ts[1:]) / 1e9, 2), "times_s": [round(t, 4) for t in ts],
                                              "segments": seg.value, "whole": whole.value}
# plain Python file IO around the memory entry point, for scale (read(), decode, write())
t0 = time.perf_counter(); raw = open(src, 'rb').read(); C.memmove(h_in, raw, len(raw))
L.b200z_gzip_decode(h_in, len(raw), 0, h_out, out_bytes + 64, C.byref(n)); open(dst, 'wb').write(C.string_at(h_out, n.value))
res["read_decode_write_GBps_out"] = round(out_bytes / (time.perf_counter() - t0) / 1e9, 2)
os.unlink(src); os.unlink(dst)
# ZipEncoder members: one b200z_deflate_raw call each vs one b200z_deflate_batch call
members, msize = int(os.environ.get('ENC_MEMBERS', 64)), int(os.environ.get('ENC_MEMBER_KIB', 1024)) << 10
txt = synth.text(members * msize, stream=701).tobytes()
items = [txt[i * msize:(i + 1) * msize] for i in range(members)]
enc = {}
for level in (1, 6):
    t0 = time.perf_counter(); one = [a.Deflate(it, level=level).get_bytes() for it in items]; t1 = time.perf_counter() - t0
    row = {"one_by_one_GBps_in": round(members * msize / t1 / 1e9, 3)}
    for lanes in ('1', '4', '8'):
        os.environ['B200Z_DEFLATE_LANES'] = lanes
        deflate_batch(items[:2], level)
        t0 = time.perf_counter(); got = deflate_batch(items, level); t2 = time.perf_counter() - t0
        assert [g[0] for g in got] == one
        row[f"batch_{lanes}_lanes_GBps_in"] = round(members * msize / t2 / 1e9, 3)
    enc[f"level{level}"] = row
res["deflate_members"] = {"members": members, "member_KiB": msize >> 10, **enc}
print(json.dumps(res))
