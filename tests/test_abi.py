"""The C-ABI library loads without a GPU and exports every symbol include/b200z.h declares; compute
entry points fail loudly (no CPU fallback) when no device is present.  CPU only."""
import ctypes as C
import os
import re

import pytest

from archive_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200z.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200z_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_ffi.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, f"libb200z.so lacks {missing}"


def test_python_binding_covers_header():
    assert sorted(_ffi.declared_symbols()) == header_symbols()


def test_version_and_host_only_calls():
    L = _ffi.lib()
    assert L.b200z_version().startswith(b"b200z")
    assert L.b200z_device_count() >= 0
    assert isinstance(L.b200z_last_error(), bytes)


def test_gzip_bound_is_host_framing_only():
    """b200z_gzip_bound walks BGZF 'BC' hints + ISIZE: pure header arithmetic, no device needed."""
    from archive_b200 import synth
    import numpy as np
    data = np.frombuffer(bytes(range(256)) * 1024, dtype=np.uint8)
    members = synth.gzip_members(data, unit=65536, workers=1)
    blob = b"".join(members)
    L = _ffi.lib()
    addr, n, keep = _ffi.as_buffer(blob)
    assert L.b200z_gzip_bound(addr, n) == len(data)
    plain = synth.gzip_member(bytes(1000), hint=False)
    addr, n, keep = _ffi.as_buffer(plain)
    assert L.b200z_gzip_bound(addr, n) == 0  # unknown without hints


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    import archive_b200 as a
    with pytest.raises(a.B200ZError) as ei:
        a.Inflate(b"\x03\x00")
    assert ei.value.code == _ffi.E_NODEVICE
    with pytest.raises(a.B200ZError):
        a.GZipDecoder().decode_bytes(b"\x1f\x8b\x08\x00" + bytes(20))
