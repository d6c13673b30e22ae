// bz2enc_emul.cpp -- TEST INFRASTRUCTURE: the device BZip2 encoder compiled for the CPU emulation (cuda_emu.h).
#define B200Z_EMU 1
#include "../../archive_b200/csrc/bzip2_enc_kernels.cu"

extern "C" int emu_bzip2_encode(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *stats4) {
  using namespace b200z::bz2e;
  Plan p = plan(n, (size_t)3 << 30);
  void *ws = calloc(p.ws_bytes, 1);
  uint8_t *obuf = (uint8_t *)calloc(out_cap + 16, 1);
  Stats st{0, 0, 0, 0};
  int rc = encode_device(in, n, obuf, out_cap, ws, p, out_len, &st, nullptr);
  if (rc == 0) memcpy(out, obuf, *out_len);
  if (stats4) memcpy(stats4, &st, 16);
  free(ws);
  free(obuf);
  return rc;
}
