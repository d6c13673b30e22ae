"""TEST INFRASTRUCTURE: build tests/host_emul/libb200z_emu.so -- the WHOLE product library (C ABI, host orchestration and
every kernel) compiled for the host against the CUDA execution-model emulation (cuda_emu.h).  The sources are generated
copies of archive_b200/csrc/*.cu (gen_emul.py: launch syntax and shared-memory declarations only); "device" memory is host
memory.  With B200Z_LIB pointing at it (tests/conftest.py does that when B200Z_EMU_TESTS=1) the `-m gpu` parity tests run in
a container without a GPU: functional coverage of the product code, not of the hardware (no timing, no memory-model
effects) -- the GPU tier stays the parity gate.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "archive_b200", "csrc")
sys.path.insert(0, HERE)
import gen_emul  # noqa: E402

SO = os.path.join(HERE, "libb200z_emu.so")


def build(force: bool = False, asan: bool = False) -> str:
    """asan=True: the same sources with -fsanitize=address as libb200z_emu_asan.so -- out-of-bounds accesses of "device"
    memory (what a GPU reports as an illegal address, or silently survives) stop the run.  Use:
      B200Z_EMU_TESTS=1 B200Z_LIB=tests/host_emul/libb200z_emu_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
      ASAN_OPTIONS=detect_leaks=0 python -m pytest tests -m gpu"""
    if asan:
        return _build(force, os.path.join(HERE, "libb200z_emu_asan.so"), ["-O1", "-fsanitize=address", "-fno-omit-frame-pointer"],
                      ".asan.o")
    return _build(force, SO, ["-O2"], ".emu.o")


def _build(force: bool, so: str, opt: list, suffix: str) -> str:
    units = [gen_emul.generate(ROOT, n) for n in ("b200z_api.cu", "b200z_file.cu", "b200z_multi.cu", "inflate_kernels.cu", "bzip2_kernels.cu", "deflate_kernels.cu")]
    units.append(os.path.join(CSRC, "bzip2_enc_kernels.cu"))  # carries its own B200Z_EMU switch
    deps = units + [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "b200z.h")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inl"))]
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(os.path.getmtime(d) for d in deps):
        return so
    flags = [*opt, "-g", "-fPIC", "-std=c++17", "-DB200Z_EMU=1", "-w", "-I", os.path.join(HERE, "shim"), "-I", HERE, "-I", CSRC]
    objs = [os.path.join(HERE, "_gen", os.path.basename(u).split(".")[0] + suffix) for u in units]

    def cc(job):
        src, obj = job
        subprocess.run(["g++", *flags, "-x", "c++", "-c", src, "-o", obj], check=True)

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        list(ex.map(cc, zip(units, objs)))
    subprocess.run(["g++", "-shared", *([f for f in opt if f.startswith("-fsanitize")]), "-o", so, *objs, "-lpthread"], check=True)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
