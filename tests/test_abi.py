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


def test_late_entry_points_fail_without_device(tmp_path):
    """The file entry point and the member batch have no CPU path either: B200Z_E_NODEVICE before a file is touched."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    L = _ffi.lib()
    src, dst = tmp_path / "in.gz", tmp_path / "out.bin"
    src.write_bytes(b"\x1f\x8b\x08\x00" + bytes(20))
    used, got = C.c_uint64(7), C.c_uint64(7)
    rc = L.b200z_file_codec(_ffi.FILE_GZIP_DECODE, str(src).encode(), 0, 2**64 - 1, str(dst).encode(), 0, 0, 0, 0,
                            C.byref(used), C.byref(got))
    assert rc == _ffi.E_NODEVICE and (used.value, got.value) == (0, 0) and not dst.exists()
    z = (C.c_uint64 * 1)(0)
    n = (C.c_uint64 * 1)(3)
    cap = (C.c_uint64 * 1)(64)
    out = (C.c_uint8 * 64)()
    ol, st = (C.c_uint64 * 1)(), (C.c_int32 * 1)()
    data = (C.c_uint8 * 3)(1, 2, 3)
    assert L.b200z_deflate_batch(data, z, n, 1, 6, 15, out, z, cap, ol, None, st) == _ffi.E_NODEVICE
    import archive_b200 as a
    tgz = tmp_path / "in.tgz"
    tgz.write_bytes(src.read_bytes())
    with pytest.raises(a.B200ZError) as ei:
        a.extract_file_to_disk(str(tgz), str(tmp_path / "o"))
    assert ei.value.code == _ffi.E_NODEVICE


def test_dart_binding_names_every_symbol():
    """dart/lib/src/b200z_ffi.dart cannot be compiled here (no Dart SDK); at least it must look up exactly the symbols the
    header declares -- no more, no fewer."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dart = open(os.path.join(root, "dart", "lib", "src", "b200z_ffi.dart")).read()
    bound = set(re.findall(r"lookupFunction<[^>]+>\(\s*'(b200z_[a-z0-9_]+)'\)", dart))
    assert bound == set(_ffi.declared_symbols())
