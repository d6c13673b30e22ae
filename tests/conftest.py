"""Test tiers (mirrors the reference's single `dart test` tier, SURVEY.md section 4, split by device):
  -m "not gpu": oracle vs the reference's golden vectors, host logic, C-ABI surface, kernel-logic emulation
  -m gpu      : parity tests proper -- CUDA path through the C ABI vs the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_device: a gpu test that cannot run on the emulated library (device tensors, timing)")
    # checker libraries (test infrastructure): the oracle and the host emulation of the decode logic
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    emul = os.path.join(ROOT, "tests", "host_emul")
    src = os.path.join(emul, "emul_decode.cpp")
    so = os.path.join(emul, "libemul.so")
    hdr = os.path.join(ROOT, "archive_b200", "csrc", "inflate_decode.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-g", "-fPIC", "-shared", "-std=c++17", "-x", "c++", src, "-o", so], check=True)
    # the inflate kernels (several lanes per stream) on the CUDA execution-model emulation
    src = os.path.join(emul, "inflate_emul.cpp")
    so = os.path.join(emul, "libinflate_emul.so")
    deps = [src, os.path.join(emul, "cuda_emu.h"), hdr, os.path.join(ROOT, "archive_b200", "csrc", "inflate_kernels.cu"),
            os.path.join(ROOT, "archive_b200", "csrc", "inflate_fast.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-I", emul, "-I",
                        os.path.join(ROOT, "archive_b200", "csrc"), src, "-o", so], check=True)
    # ... and the same kernels with k_inflate_fast's LZ77 pass by blocks (a build option of inflate_fast.cuh, -DFP_LZBLK=1)
    so = os.path.join(emul, "libinflate_emul_lzblk.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-DFP_LZBLK=1", "-I", emul, "-I",
                        os.path.join(ROOT, "archive_b200", "csrc"), src, "-o", so], check=True)
    # the BZip2 encoder kernels, compiled against the CUDA execution-model emulation (tests/host_emul/cuda_emu.h)
    csrc = os.path.join(ROOT, "archive_b200", "csrc")
    src = os.path.join(emul, "bz2enc_emul.cpp")
    so = os.path.join(emul, "libbz2enc_emul.so")
    deps = [src, os.path.join(emul, "cuda_emu.h")] + [
        os.path.join(csrc, f) for f in os.listdir(csrc) if f.startswith("bzip2_enc")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-I", emul, "-I", csrc, src, "-o", so],
                       check=True)


    # the BZip2 decode kernels on the emulation, from a generated copy of the product source (gen_emul.py)
    sys.path.insert(0, emul)
    import gen_emul
    inc = gen_emul.generate(ROOT, "bzip2_kernels.cu")
    src = os.path.join(emul, "bz2dec_emul.cpp")
    so = os.path.join(emul, "libbz2dec_emul.so")
    deps = [src, inc, os.path.join(emul, "cuda_emu.h"), os.path.join(csrc, "b200z_internal.h"), os.path.join(csrc, "bz2_rnums.h"),
            os.path.join(ROOT, "include", "b200z.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-w", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(emul, "shim"), "-I", emul, "-I",
                        csrc, src, "-o", so], check=True)

    # the Deflate encoder kernels + their host driver on the emulation (gen_emul.py)
    inc = gen_emul.generate(ROOT, "deflate_kernels.cu")
    src = os.path.join(emul, "deflate_emul.cpp")
    so = os.path.join(emul, "libdeflate_emul.so")
    deps = [src, inc, os.path.join(emul, "cuda_emu.h"), os.path.join(csrc, "b200z_internal.h"), os.path.join(ROOT, "include", "b200z.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-g", "-w", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(emul, "shim"), "-I", emul, "-I",
                        csrc, src, "-o", so], check=True)


EMU_TESTS = os.environ.get("B200Z_EMU_TESTS") == "1"
if EMU_TESTS:
    # Run the `-m gpu` parity tests without a GPU: the whole product library compiled against the CUDA execution-model
    # emulation (tests/host_emul/build_emu_lib.py).  Functional coverage of the product code on the build container; the
    # B200 run stays the gate.  Tests that need real device memory / torch.cuda mark themselves `needs_device`.
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_emul"))
    import build_emu_lib
    _emu_lib = build_emu_lib.build()
    os.environ.setdefault("B200Z_LIB", _emu_lib)  # (a sanitizer build of the same sources may be named from outside)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    if EMU_TESTS:
        skip = pytest.mark.skip(reason="needs a real CUDA device (emulated library run)")
        for it in items:
            if "needs_device" in it.keywords:
                it.add_marker(skip)
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
