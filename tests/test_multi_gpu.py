"""Several GPUs driven by ONE process through the C ABI (include/b200z.h "several GPUs of one box"):
b200z_gzip_decode_multi / b200z_inflate_batch_multi deal the members (units) to the devices of b200z_multi_init and must
give byte for byte what the oracle gives -- with one device, with every device of the box, and with B200Z_MULTI_GATHER
(the shards exchanged over NVLink so that every device holds the whole stream)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu
EMU = os.environ.get("B200Z_EMU_TESTS") == "1"


@pytest.fixture(scope="module")
def L():
    from archive_b200 import _ffi
    return _ffi.ensure_init()


def masks(L):
    n = L.b200z_device_count()
    out = [1]
    if n >= 2:
        out.append(3)
    if n > 2:
        out.append((1 << n) - 1)
    return out


def decode_multi(L, blob: bytes, cap: int, flags: int):
    from archive_b200 import _ffi
    addr, n, keep = _ffi.as_buffer(blob)
    out = (C.c_uint8 * max(cap, 1))()
    got = C.c_size_t(0)
    rc = L.b200z_gzip_decode_multi(addr, n, 0, out, cap, C.byref(got), flags)
    return rc, bytes(out[:min(got.value, cap)]), got.value


def test_gzip_members_dealt_to_the_devices(L):
    from archive_b200 import synth
    text = synth.text(48 * 65536 + 777, stream=71)
    ms = synth.gzip_members(text, workers=1)
    blob = b"".join(ms)
    st, want = orc.gzip_decode(blob)
    assert st == orc.OK and want == text.tobytes()
    for mask in masks(L):
        assert L.b200z_multi_init(mask, 0) == 0, L.b200z_last_error()
        try:
            assert L.b200z_multi_device_count() == bin(mask).count("1")
            rc, got, n = decode_multi(L, blob, len(want) + 100, 0)
            assert rc == 0 and got == want, (mask, rc, n)
            # members without hints, a member whose hint lies, junk behind the run: the careful path, same bytes as the oracle
            odd = ms[0] + synth.gzip_member(text[:5000].tobytes(), hint=False) + ms[1]
            st2, want2 = orc.gzip_decode(odd)
            rc, got, n = decode_multi(L, odd, len(want2) + 100, 0)
            assert (rc == 0) == (st2 == orc.OK) and got == want2
            liar = bytearray(ms[0] + ms[1])
            liar[len(ms[0]) - 4] ^= 1  # ISIZE of the first member
            st3, want3 = orc.gzip_decode(bytes(liar))
            rc, got, n = decode_multi(L, bytes(liar), len(want3) + 100000, 0)
            assert (rc == 0) == (st3 == orc.OK) and got == want3
            # too little room is reported, not overrun
            rc, got, n = decode_multi(L, blob, len(want) - 1, 0)
            assert rc == -3
        finally:
            L.b200z_multi_shutdown()


def test_batch_units_and_statuses(L):
    from archive_b200 import synth
    text = synth.text(24 * 65536, stream=72).tobytes()
    plain = [text[i * 65536:(i + 1) * 65536] for i in range(24)]
    units = [synth.deflate_raw(p) + bytes(8) for p in plain]
    units[5] = units[5][:1000]                     # ends inside a block
    units[9] = b"\x07" + bytes(200)                # reserved block type
    caps = [65536] * 24
    caps[11] = 1000                                # output beyond out_cap
    blob = b"".join(units)
    in_off = np.cumsum([0] + [len(u) for u in units[:-1]]).astype(np.uint64)
    in_len = np.array([len(u) for u in units], dtype=np.uint32)
    out_off = (np.arange(24, dtype=np.uint64) * 65536)
    out_cap = np.array(caps, dtype=np.uint32)
    ref = [orc.emul_inflate(u, c) for u, c in zip(units, caps)]
    for mask in masks(L):
        assert L.b200z_multi_init(mask, 0) == 0, L.b200z_last_error()
        try:
            out = np.zeros(24 * 65536, dtype=np.uint8)
            ol, st, iu = np.zeros(24, np.uint32), np.zeros(24, np.int32), np.zeros(24, np.uint32)
            rc = L.b200z_inflate_batch_multi(blob, len(blob), in_off.ctypes.data, in_len.ctypes.data, out.ctypes.data, out.size,
                                             out_off.ctypes.data, out_cap.ctypes.data, ol.ctypes.data, st.ctypes.data,
                                             iu.ctypes.data, 24, 0)
            assert rc == 0, L.b200z_last_error()
            for i, (rst, rout, rused, _) in enumerate(ref):
                assert st[i] == rst and ol[i] == len(rout), (mask, i, st[i], rst)
                assert out[i * 65536:i * 65536 + min(len(rout), caps[i])].tobytes() == rout[:caps[i]], (mask, i)
                if rst not in (-2, -3):
                    assert iu[i] == rused, (mask, i)
        finally:
            L.b200z_multi_shutdown()


@pytest.mark.needs_device
def test_gather_leaves_the_whole_stream_on_every_device(L):
    """B200Z_MULTI_GATHER: after the call every device holds the stream in block order (checked through torch views of the
    devices' buffers), and the host copy is the same bytes."""
    import torch  # noqa: F401  (loads the CUDA runtime this test copies with)
    from archive_b200 import synth
    text = synth.text(64 * 65536, stream=73)
    blob = b"".join(synth.gzip_members(text, workers=1))
    want = text.tobytes()
    n = L.b200z_device_count()
    mask = (1 << n) - 1
    assert L.b200z_multi_init(mask, 0) == 0, L.b200z_last_error()
    try:
        rc, got, k = decode_multi(L, blob, len(want), 1)
        if rc == -1 and b"nccl" in L.b200z_last_error().lower():
            pytest.skip("no NCCL on this box: " + L.b200z_last_error().decode())
        assert rc == 0 and got == want, (rc, L.b200z_last_error())
        for slot in range(n):
            nb = C.c_size_t(0)
            p = L.b200z_multi_device_output(slot, C.byref(nb))
            assert p and nb.value == len(want)
            host = (C.c_uint8 * nb.value)()
            rt = C.CDLL("libcudart.so.12")  # (the runtime torch has already loaded: a raw device pointer has no torch owner)
            assert rt.cudaSetDevice(slot) == 0
            assert rt.cudaMemcpy(host, C.c_void_p(p), C.c_size_t(nb.value), 2) == 0  # cudaMemcpyDeviceToHost
            assert bytes(host) == want, slot
    finally:
        L.b200z_multi_shutdown()
