"""world_size-2 gloo test of the multi-GPU host logic (no GPU): members are split by rank, every rank decodes its
range (here with the CPU oracle standing in for the kernels -- tests may use it as the checker), writes into its
slice of the full buffer, and ONE all-gather reassembles the stream in member order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import oracle_lib as orc
    from archive_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    unit = 4096
    text = synth.text(64 * unit, stream=21)
    members = synth.gzip_members(text, unit=unit, workers=1)
    sizes = [len(m) for m in members]
    ranges = shard.split_units(sizes, world)
    lo, hi = ranges[rank]
    # equal-size units -> equal slices only if the unit counts match; use the ragged (all-gather-v) form
    out_len = [unit] * len(members)
    slices = shard.out_slices(out_len, ranges)
    st, mine = orc.gzip_decode(b"".join(members[lo:hi]))
    assert st == orc.OK and len(mine) == slices[rank][1] - slices[rank][0]
    # all-gather-v: sizes first, then padded payloads
    max_len = max(b - a for a, b in slices)
    buf = torch.zeros(max_len, dtype=torch.uint8)
    buf[:len(mine)] = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    gathered = [torch.zeros(max_len, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, buf)
    full = b"".join(bytes(gathered[r][:slices[r][1] - slices[r][0]].numpy()) for r in range(world))
    ok = full == text.tobytes() and sum(hi - lo for lo, hi in ranges) == len(members)
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


def test_two_rank_member_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


def test_split_is_contiguous_and_balanced():
    from archive_b200 import shard
    rng = np.random.default_rng(1)
    sizes = rng.integers(1000, 40000, size=1000)
    for world in (1, 2, 4, 8):
        rs = shard.split_units(sizes, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(sizes)
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        loads = [int(sizes[a:b].sum()) for a, b in rs]
        assert max(loads) - min(loads) <= 2 * int(sizes.max())
    assert shard.split_units([], 4) == [(0, 0)] * 4
