// deflate_emul.cpp -- TEST INFRASTRUCTURE: the Deflate ENCODER kernels and their host driver (deflate_slow_device /
// deflate_stored_device of archive_b200/csrc/deflate_kernels.cu) executed on the CUDA execution-model emulation, so that
// the CPU test tier checks the kernels themselves -- not only the model of their reformulation (deflate_model.cpp) --
// against the oracle's line-by-line Deflate.  Compiled from a generated copy of the product file (gen_emul.py: only the
// <<<>>> syntax and the dynamic shared memory declarations differ).
#include "cuda_emu.h"
#include "b200z_internal.h"
namespace b200z {
void count_launch() {}
}  // namespace b200z
#include "_gen/deflate_kernels_emu.inc"

#include <vector>

using namespace b200z;

// Deflate(bytes, level:, windowBits:) raw stream (deflate.dart:25-100), levels 1..9 -> 0, or a negative code.
// stats[3]: tokens, blocks, re-speculated chunks.
extern "C" int emu_deflate_raw(const uint8_t *in, size_t n, int level, int window_bits, uint8_t *out, size_t out_cap,
                               size_t *out_len, uint32_t *stats) {
  if (level < 1 || level > 9) return -1;
  std::vector<uint8_t> d_in(n + 64, 0);
  if (n) memcpy(d_in.data(), in, n);
  const size_t cap = deflate_bound(n) + 8;
  std::vector<uint32_t> d_out((cap + 64) / 4 + 1);
  const size_t wsb = deflate_workspace_bytes(n);
  std::vector<unsigned long long> ws(wsb / 8 + 64);
  size_t got = 0;
  cudaError_t e = deflate_slow_device(d_in.data(), n, level, window_bits, (uint8_t *)d_out.data(), cap, ws.data(), wsb + 256, &got,
                                      stats, nullptr);
  if (e != cudaSuccess) return -100 - (int)e;
  *out_len = got;
  if (got > out_cap) return -3;
  memcpy(out, d_out.data(), got);
  return 0;
}

extern "C" size_t emu_deflate_bound(size_t n) { return deflate_bound(n) + 8; }
