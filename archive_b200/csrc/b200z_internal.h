// b200z_internal.h -- shared declarations between the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200z.h"

#ifndef B200Z_LBITS
#define B200Z_LBITS 9  // literal/length primary LUT bits (2^9 x u16 per stream)
#endif
#ifndef B200Z_DBITS
#define B200Z_DBITS 9  // distance primary LUT bits
#endif
#define B200Z_DECODE_THREADS 32   // one warp per block: finest block-scheduler granularity
#define B200Z_EXPAND_THREADS 256

namespace b200z {

struct InflateBatch {
  const uint8_t *in_base;
  const uint64_t *in_off;
  const uint32_t *in_len;
  uint8_t *out_base;
  const uint64_t *out_off;
  const uint32_t *out_cap;
  uint32_t *out_len;
  int32_t *status;
  uint32_t *in_used;
  size_t n_units;
  void *workspace;   // [tok_bytes of tokens][n_units x u32 token counts]
  size_t tok_bytes;  // 4 * extent of the output layout, rounded up to 256
  int share = 1;     // how many batches run concurrently on the device (sizes the streams-per-warp choice)
};

cudaError_t launch_inflate(const InflateBatch &b, cudaStream_t stream);
void count_launch();
void profile_enable(bool on);
int profile_read(double *decode_ms, double *expand_ms, uint64_t *n);

}  // namespace b200z
