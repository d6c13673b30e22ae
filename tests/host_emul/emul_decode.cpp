// tests/host_emul/emul_decode.cpp -- TEST INFRASTRUCTURE.
// Compiles the per-stream decode logic of the CUDA kernel (archive_b200/csrc/inflate_decode.cuh) as
// plain C++ so the CPU-only test tier can check the kernel's LOGIC (bit reader, table build, quirks,
// token stream) against the oracle without a GPU.  It is never loaded by the product.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../archive_b200/csrc/inflate_decode.cuh"

using namespace b200z;

// sequential token expansion (the reference semantics of writeByte / writeBackReference)
static void expand(const uint32_t *tok, uint32_t nt, const uint8_t *in, uint8_t *out) {
  size_t o = 0;
  for (uint32_t i = 0; i < nt; ++i) {
    uint32_t t = tok[i];
    if (t & TOK_LIT) {
      out[o++] = (uint8_t)t;
    } else if (t & TOK_STORED) {
      uint32_t len = t & 0xffff;
      size_t pos = ((size_t)((t >> 16) & 3) << 30) | tok[++i];
      memcpy(out + o, in + pos, len);
      o += len;
    } else if (t) {
      uint32_t len = t >> 16, dist = t & 0xffff;
      for (uint32_t k = 0; k < len; ++k, ++o) out[o] = out[o - dist];
    }
  }
}

extern "C" int emul_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t cap, uint32_t *out_len,
                            uint32_t *in_used, int32_t *status, uint32_t *ntok_out) {
  // the kernel reads aligned 32-bit words: give the stream a padded, deliberately misaligned home
  std::vector<uint8_t> home(in_len + 64, 0xA5);
  uint8_t *p = home.data() + 4 + 1;
  while (((uintptr_t)p & 3) != 1) ++p;
  memcpy(p, in, in_len);
  std::vector<uint32_t> tok((size_t)cap + 64);
  std::vector<uint16_t> lut(LUT_HALFWORDS + 8);
  uint32_t xtab[64];
  for (int i = 0; i < 32; ++i) {
    xtab[i] = c_len_tab[i];
    xtab[32 + i] = c_dist_tab[i];
  }
  SpecCtx sc{};  // one lane per stream: no speculative helpers in this single-threaded build
  sc.G = 1;
  sc.spec = false;
  sc.count_only = false;
  alignas(16) static uint32_t stage[4];
  sc.stage = stage;
  UnitResult r = inflate_decode_unit(true, p, in_len, cap, tok.data(), lut.data(), lut.data() + (1 << B200Z_LBITS),
                                     c_len_tab, c_dist_tab, xtab, sc);
  expand(tok.data(), r.ntok, p, out);
  *out_len = r.out_len;
  *in_used = r.in_used;
  *status = r.status;
  if (ntok_out) *ntok_out = r.ntok;
  return 0;
}
