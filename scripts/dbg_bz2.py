import bz2, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ['B200Z_DEBUG'] = '1'
import archive_b200 as a, oracle_lib as orc
from archive_b200 import synth
d = synth.text(300000, stream=8).tobytes()
z = bz2.compress(d, 1)
bad = bytearray(z); bad[len(z) * 5 // 8] ^= 0x10
ost, oout = orc.bzip2_decode(bytes(bad), verify=True)
print('oracle', ost, len(oout), 'prefix ok', d.startswith(oout[:100000]))
out = a.OutputMemoryStream()
try:
    ok = a.BZip2Decoder().decode_stream(a.InputMemoryStream(bytes(bad)), out, verify=True)
    print('gpu ok', ok)
except Exception as e:
    print('gpu exc', e)
got = out.get_bytes()
n = next((i for i in range(min(len(got), len(d))) if got[i] != d[i]), min(len(got), len(d)))
print('gpu len', len(got), 'first mismatch at', n, 'oracle==gpu', got == oout)
