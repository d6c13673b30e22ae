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


def build(force: bool = False) -> str:
    units = [gen_emul.generate(ROOT, n) for n in ("b200z_api.cu", "b200z_file.cu", "inflate_kernels.cu", "bzip2_kernels.cu", "deflate_kernels.cu")]
    units.append(os.path.join(CSRC, "bzip2_enc_kernels.cu"))  # carries its own B200Z_EMU switch
    deps = units + [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "b200z.h")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inl"))]
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(d) for d in deps):
        return SO
    flags = ["-O2", "-g", "-fPIC", "-std=c++17", "-DB200Z_EMU=1", "-w", "-I", os.path.join(HERE, "shim"), "-I", HERE, "-I", CSRC]
    objs = [os.path.join(HERE, "_gen", os.path.basename(u).split(".")[0] + ".emu.o") for u in units]

    def cc(job):
        src, obj = job
        subprocess.run(["g++", *flags, "-x", "c++", "-c", src, "-o", obj], check=True)

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        list(ex.map(cc, zip(units, objs)))
    subprocess.run(["g++", "-shared", "-o", SO, *objs, "-lpthread"], check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
