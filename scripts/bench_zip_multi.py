"""Config 5 across GPUs: a .zip of ZIP_MEMBERS x ZIP_MEMBER_MIB text members (method 8, a full-flush point every
64 KiB), members packed largest-first over the ranks of a torchrun job (shard.pack_members: deterministic, no rank talks),
each rank decoding its share with ONE b200z_zip_extract call into its own pinned buffer.  Strong scaling: the archive is fixed.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_zip_multi.py
ZIP_BACKEND=gloo + B200Z_EMU_TESTS=1 + B200Z_LIB=tests/host_emul/libb200z_emu.so dry-runs the script without a GPU."""
import ctypes as C, io, json, os, sys, time, zipfile, zlib
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from archive_b200 import _ffi, shard, synth
from concurrent.futures import ThreadPoolExecutor
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
backend = os.environ.get("ZIP_BACKEND", "nccl")
cuda = backend == "nccl"
if cuda:
    torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group(backend, **({"device_id": torch.device("cuda", local)} if cuda else {}))
L = _ffi.ensure_init(local if cuda else 0)
n, size = int(os.environ.get('ZIP_MEMBERS', 1024)), int(float(os.environ.get('ZIP_MEMBER_MIB', 4)) * (1 << 20))
txt = synth.text(n * size, stream=700)
def comp(i):
    b = txt[i * size:(i + 1) * size].tobytes()
    return synth.deflate_raw_flushed(b, 65536), zlib.crc32(b)
with ThreadPoolExecutor(32) as ex:
    parts = list(ex.map(comp, range(n)))
data = synth.zip_from_deflated([(f"member{i:04d}.txt", z, crc, size) for i, (z, crc) in enumerate(parts)])
zl = len(data); h_in = L.b200z_host_alloc(zl); C.memmove(h_in, data, zl)
cnt = C.c_size_t(0); ents = (_ffi.ZipEntry * n)()
assert L.b200z_zip_list(h_in, zl, ents, n, C.byref(cnt)) == 0 and cnt.value == n  # host work, every rank
mine = shard.pack_members([e.comp_size for e in ents], world)[rank]
k = len(mine)
sub = (_ffi.ZipEntry * max(k, 1))(*[ents[i] for i in mine])
tot = k * size; h_out = L.b200z_host_alloc(max(tot, 1))
off = (C.c_uint64 * max(k, 1))(*[j * size for j in range(k)]); room = (C.c_uint64 * max(k, 1))(*[size] * k)
ol = (C.c_uint64 * max(k, 1))(); st = (C.c_int32 * max(k, 1))()
def sync():
    if cuda: torch.cuda.synchronize()
    if world > 1: dist.barrier()
times = []
for it in range(4):
    sync(); t0 = time.perf_counter()
    rc = L.b200z_zip_extract(h_in, zl, sub, k, h_out, tot, off, room, ol, st, 0) if k else 0
    sync(); times.append(time.perf_counter() - t0)
    assert rc == 0, _ffi.last_error()
ok = all(st[j] == 0 and ol[j] == size for j in range(k))
for j in range(0, k, 7):
    ok = ok and zlib.crc32(C.string_at(h_out + j * size, size)) == parts[mine[j]][1]
best = min(times[1:])
if world > 1:
    dev = "cuda" if cuda else "cpu"
    tt = torch.tensor([best], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t2 = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(t2, op=dist.ReduceOp.MIN)
    best, ok = float(tt[0]), bool(t2[0] > 0.5)
if rank == 0:
    print(json.dumps({"metric": "zip_extract_uncompressed_GBps_e2e", "n_gpus": world, "value": round(n * size / best / 1e9, 3),
                      "members": n, "member_MiB": size / (1 << 20), "zip_bytes": zl, "best_s": round(best, 4), "ok": ok,
                      "scaling": "strong", "members_on_rank0": k}))
if world > 1: dist.destroy_process_group()
