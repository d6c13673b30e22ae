#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: Inflate GB/s (uncompressed) on 1 GiB synthetic DEFLATE.

Workload (BASELINE config 2): 16 384 gzip members x 64 KiB of synthetic wiki-like text,
one dynamic-Huffman block per member, ~394 MiB compressed -> 1 GiB, per GPU.

  value     device-resident: compressed members already in HBM -> decoded bytes in HBM
            (b200z_inflate_batch_device: k_inflate_decode + k_inflate_expand), CUDA events, max over ranks.
            Multi-GPU = one process per GPU, members sharded by rank ("weak": every rank owns its own 1 GiB),
            no data-path collective; the north-star "reassemble with one all-gather" variant is timed
            separately and reported under "with_allgather".
  e2e       the same metric through the reference-facing call GZipDecoderWeb.decodeBytes ==
            b200z_gzip_decode(host in, host out): pinned host buffers, H2D + framing walk + kernels + D2H
            inside the timed region.
  roofline  HBM bound; algorithmic bytes = C + U per pass (compressed read once + output written once).
  cpu_baseline / --impl reference
            the C oracle (a restatement of the pure-Dart Inflate; the reference itself is Dart and there is
            no Dart SDK here) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = 65536
N_UNITS = int(os.environ.get("B200Z_BENCH_UNITS", 16384))
CACHE = os.environ.get("B200Z_CACHE", "/tmp/b200z_cache")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=False)
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    return L


def oracle_gzip_throughput(blob: np.ndarray, member_off: np.ndarray, n_sample_units: int, threads: int, repeats: int = 1):
    """Times the oracle's gzip member loop over `n_sample_units` members split across `threads` host threads
    (ctypes drops the GIL).  -> (GB/s uncompressed, seconds, bytes_out)"""
    from concurrent.futures import ThreadPoolExecutor
    L = load_oracle()
    n_sample_units = min(n_sample_units, len(member_off) - 1)
    per = max(1, n_sample_units // threads)
    ranges = [(i, min(i + per, n_sample_units)) for i in range(0, n_sample_units, per)]
    raw = blob.tobytes() if not isinstance(blob, bytes) else blob

    def job(rg):
        a, b = rg
        lo, hi = int(member_off[a]), int(member_off[b])
        seg = raw[lo:hi]
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        st = L.orc_gzip_decode_bytes(seg, C.c_size_t(len(seg)), 0, C.byref(out), C.byref(n))
        got = n.value
        L.orc_free(out)
        assert st == 0 and got == (b - a) * UNIT, (st, got)
        return got

    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            total = sum(ex.map(job, ranges))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return total / best / 1e9, best, total


def gpu_local_cpus(local_rank: int):
    """CPUs on the NUMA node the GPU's PCIe link hangs off (/sys/bus/pci/devices/<bdf>/local_cpulist), restricted to the
    CPUs this process may use; None when the box does not say.  One process per GPU bound to its GPU's node is how the
    end-to-end path is meant to be deployed: the pinned staging buffers are then allocated on the memory the DMA reaches
    without crossing the socket link."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.  The
    reference is pure Dart and cannot run here (no Dart SDK, no network): the arm times the C oracle, a
    restatement of the same algorithm (kind "port"), with all host threads."""
    if rank != 0:
        return
    from archive_b200 import synth
    cores = os.cpu_count() or 1
    sample_units = int(os.environ.get("B200Z_REF_UNITS", min(N_UNITS, 256 * max(1, cores // 2))))
    w = synth.gzip_workload(sample_units, UNIT, stream0=0, cache_dir=CACHE)
    for _ in range(max(1, args.warmup if args.warmup < 2 else 1)):
        oracle_gzip_throughput(w["blob"], w["member_off"], sample_units, cores)
    times = []
    total = 0
    for _ in range(args.steps):
        g, dt, total = oracle_gzip_throughput(w["blob"], w["member_off"], sample_units, cores)
        times.append(dt)
    t = sum(times)
    val = total * len(times) / t / 1e9
    line = {
        "impl": "reference", "metric": "inflate_uncompressed_GBps", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "gzip-multimember-64KiB-dynamic (BASELINE config 2)", "unit_bytes": UNIT,
                   "sample_units_per_step": sample_units, "bytes_per_step": total},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_units} members ({total >> 20} MiB out) per step, C oracle restating "
                                   "the pure-Dart GZipDecoderWeb/Inflate, one member range per host thread"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200z")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from archive_b200 import _ffi, synth
    L = _ffi.ensure_init(local_rank)

    # ---------------- workload: every rank owns its own re-drawn 1 GiB (weak scaling) ----------------
    t_gen = time.time()
    w = synth.gzip_workload(N_UNITS, UNIT, stream0=rank * 4096, cache_dir=CACHE)
    blob, moff = w["blob"], w["member_off"]
    n = N_UNITS
    hdr = 18  # synth.gzip_member with the BC hint: 10 + 2 + 6
    in_off = (moff[:-1] + hdr).astype(np.uint64)
    in_len = (moff[1:] - moff[:-1] - hdr).astype(np.uint32)  # member payload incl. the 8-byte trailer
    out_off = (np.arange(n, dtype=np.uint64) * UNIT)
    out_cap = np.full(n, UNIT, dtype=np.uint32)
    C_bytes, U_bytes = int(len(blob)), n * UNIT
    gen_s = time.time() - t_gen

    d_in = torch.empty(C_bytes + 64, dtype=torch.uint8, device=dev)
    d_in[:C_bytes].copy_(torch.from_numpy(blob.copy()))
    d_full = torch.empty(world * U_bytes, dtype=torch.uint8, device=dev)  # rank r decodes into slice r
    d_out = d_full[rank * U_bytes:(rank + 1) * U_bytes]
    d_in_off = torch.from_numpy(in_off.view(np.int64)).to(dev)
    d_in_len = torch.from_numpy(in_len.view(np.int32)).to(dev)
    d_out_off = torch.from_numpy(out_off.view(np.int64)).to(dev)
    d_out_cap = torch.from_numpy(out_cap.view(np.int32)).to(dev)
    d_out_len = torch.zeros(n, dtype=torch.int32, device=dev)
    d_status = torch.full((n,), -99, dtype=torch.int32, device=dev)
    d_used = torch.zeros(n, dtype=torch.int32, device=dev)
    ws_bytes = L.b200z_inflate_workspace_bytes(n, C_bytes, U_bytes)
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: the library launches on it, the events time it
    torch.cuda.set_stream(stream)

    def step():
        rc = L.b200z_inflate_batch_device(d_in.data_ptr(), d_in_off.data_ptr(), d_in_len.data_ptr(), d_out.data_ptr(),
                                          d_out_off.data_ptr(), d_out_cap.data_ptr(), d_out_len.data_ptr(),
                                          d_status.data_ptr(), d_used.data_ptr(), n, d_ws.data_ptr(), ws_bytes,
                                          stream.cuda_stream)
        if rc:
            raise RuntimeError(_ffi.last_error())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.b200z_profile_enable(0)
    launches0 = L.b200z_launch_count()
    ms_total = timed(step, args.steps, args.warmup)
    launches = (L.b200z_launch_count() - launches0) * args.steps // (args.steps + args.warmup)  # timed steps only
    ms_step = ms_total / args.steps
    value = world * U_bytes / (ms_step * 1e-3) / 1e9

    # correctness gate on what the timed region produced: status, lengths, and CRC-32 against the trailers
    st = d_status.cpu().numpy()
    ol = d_out_len.cpu().numpy()
    us = d_used.cpu().numpy()
    assert (st == 0).all(), f"unit status {np.unique(st)}"
    assert (ol == UNIT).all()
    assert (us.astype(np.int64) + 8 == in_len.astype(np.int64)).all()
    host_out = d_out.cpu().numpy()
    check_idx = np.linspace(0, n - 1, 512).astype(int)
    for i in check_idx:
        m_end = int(moff[i + 1])
        crc = int.from_bytes(blob[m_end - 8:m_end - 4].tobytes(), "little")
        assert zlib.crc32(host_out[i * UNIT:(i + 1) * UNIT].tobytes()) == crc, f"unit {i} CRC mismatch"

    # per-kernel breakdown (separate pass so event records do not sit inside the headline region)
    L.b200z_profile_enable(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    fms, dms, ems, nb = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
    L.b200z_profile_read(C.byref(fms), C.byref(dms), C.byref(ems), C.byref(nb))
    L.b200z_profile_enable(0)
    k_fast, k_dec, k_exp = (v.value / max(1, nb.value) for v in (fms, dms, ems))

    # ---------------- north-star variant: decode + ONE in-place all-gather ----------------
    with_gather = None
    if world > 1:
        def step_gather():
            step()
            dist.all_gather_into_tensor(d_full, d_out)
        ms_g = timed(step_gather, args.steps, args.warmup) / args.steps
        with_gather = {"value": world * U_bytes / (ms_g * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_g,
                       "collective": "ncclAllGather in place, %d MiB per rank" % (U_bytes >> 20)}

    # ---------------- end to end through the reference-facing call, host buffers ----------------
    e2e = None
    if not args.no_e2e:
        all_cpus = os.sched_getaffinity(0)
        near = gpu_local_cpus(local_rank)
        if near and near != all_cpus:
            os.sched_setaffinity(0, near)  # for the staging buffers' placement and the calling thread; undone below
        h_in = torch.empty(C_bytes, dtype=torch.uint8).pin_memory()
        h_in.numpy()[:] = blob
        h_out = torch.empty(U_bytes, dtype=torch.uint8).pin_memory()
        out_len = C.c_size_t(0)

        def e2e_step():
            rc = L.b200z_gzip_decode(h_in.data_ptr(), C_bytes, 0, h_out.data_ptr(), U_bytes, C.byref(out_len))
            if rc:
                raise RuntimeError(_ffi.last_error())

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        k = max(3, min(args.steps, 10))
        for _ in range(k):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert out_len.value == U_bytes
        ho = h_out.numpy()
        for i in check_idx[::16]:
            m_end = int(moff[i + 1])
            crc = int.from_bytes(blob[m_end - 8:m_end - 4].tobytes(), "little")
            assert zlib.crc32(ho[i * UNIT:(i + 1) * UNIT].tobytes()) == crc
        e2e = {"value": world * U_bytes / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": C_bytes,
               "d2h_bytes_per_step": U_bytes, "ms_per_step": dt * 1e3,
               "call": "b200z_gzip_decode(host in, host out) == GZipDecoderWeb.decodeBytes, pinned host buffers",
               "cpu_binding": ("%d CPUs of the GPU's NUMA node" % len(near)) if near and near != all_cpus else "none"}
        if near and near != all_cpus:
            os.sched_setaffinity(0, all_cpus)

    clocks = sampler.stop() if rank == 0 else None

    # ---------------- CPU baseline beside it (rank 0, N=1 only; bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(n, 128 * cores)
        g, dt, tot = oracle_gzip_throughput(blob, moff, sample, cores)
        cpu = {"value": g, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} members ({tot >> 20} MiB out), {dt:.2f} s, C oracle restating the pure-Dart "
                         "GZipDecoderWeb/Inflate (no Dart SDK on the box), one member range per host thread"}

    if rank == 0:
        peak, peak_src = peaks()
        kernel_s = (k_fast + k_dec + k_exp) * 1e-3
        achieved = (C_bytes + U_bytes) / kernel_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_pass")
            except Exception:
                traffic = None
        line = {
            "metric": "inflate_uncompressed_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "gzip-multimember-64KiB-dynamic (BASELINE config 2)", "units_per_gpu": n,
                       "unit_bytes": UNIT, "compressed_bytes_per_gpu": C_bytes, "uncompressed_bytes_per_gpu": U_bytes,
                       "sharding": "members by rank, no data-path collective", "l2": "inputs larger than L2 "
                       "(394 MiB in + 1 GiB out per pass vs 126 MB L2)", "generate_s": round(gen_s, 1)},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_pass": C_bytes + U_bytes,
                         "kernels": {"k_inflate_fast_ms": k_fast, "k_inflate_decode_ms": k_dec, "k_inflate_expand_ms": k_exp},
                         "note": "one pass = k_inflate_fast (all clean units, in shared memory) + the exact pair over what it "
                                 "left (nothing on this workload); achieved = (C+U) / their CUDA-event time"},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        if with_gather:
            line["with_allgather"] = with_gather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
