"""Config 4 across GPUs: one BZh9 stream, blocks sharded over the ranks of a torchrun job, per-block
reports exchanged over gloo, decoded bytes left on the rank that produced them (strong scaling: the stream is fixed).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_bz2_multi.py"""
import ctypes as C, json, os, sys, time, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, torch.distributed as dist
from archive_b200 import _ffi, shard, synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")
else:
    gloo = None
L = _ffi.ensure_init(local)
m = int(os.environ.get('BZ_MIB', 512)) << 20
src = synth.text(m, stream=200)
cap = L.b200z_bzip2_bound(m); zbuf = (C.c_uint8 * cap)(); zl = C.c_size_t(0)
rc = L.b200z_bzip2_encode(src.ctypes.data, m, C.addressof(zbuf), cap, C.byref(zl)); assert rc == 0, _ffi.last_error()
z = bytes(zbuf[:zl.value])
h_z = L.b200z_host_alloc(len(z)); C.memmove(h_z, z, len(z)); zv = (C.c_uint8 * len(z)).from_address(h_z)
ocap = m // world + (64 << 20); h_o = L.b200z_host_alloc(ocap)
def run():
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = shard.bzip2_decode_sharded(zv, verify=True, group=gloo, out_buf=(h_o, ocap))
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    return time.perf_counter() - t0, r
times = []
for i in range(4):
    t, r = run(); times.append(t)
ok = r["kind"] == "ok" and r["total"] == m
sb = src.tobytes()
for off, v in r["pieces"]:
    ok = ok and zlib.crc32(v) == zlib.crc32(sb[off:off + len(v)])
best = min(times[1:])
if world > 1:
    tt = torch.tensor([best, 1.0 if ok else 0.0], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX if True else None)
    t2 = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(t2, op=dist.ReduceOp.MIN)
    best, ok = float(tt[0]), bool(t2[0] > 0.5)
if rank == 0:
    print(json.dumps({"metric": "bzip2_decode_uncompressed_GBps", "n_gpus": world, "value": round(m / best / 1e9, 3), "out_MiB": m >> 20,
                      "in_bytes": len(z), "n_chain": r["n_chain"], "best_s": round(best, 4), "ok": ok, "scaling": "strong",
                      "my_pieces": len(r["pieces"])}))
if world > 1: dist.destroy_process_group()
