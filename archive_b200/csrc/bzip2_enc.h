// bzip2_enc.h -- plan / entry point of the device BZip2 encoder (bzip2_enc_kernels.cu).  Not installed.
#pragma once
#include <stddef.h>
#include <stdint.h>

#define BZ2E_BSTRIDE 901120u          // element stride of one block in every per-block array (440 tiles of 2048)
#define BZ2E_BLKBYTES (BZ2E_BSTRIDE)  // byte stride of the RLE1 block buffer

namespace b200z {
namespace bz2e {

struct BlkInfo {
  uint32_t start, end;  // input range [start, end) consumed by the block
  uint32_t e0;          // end of the block's first run (chopped from `start`)
  uint32_t c;           // end of the last closed run; in[c] (if c < n) is the byte that closed it
  uint32_t A;           // RLE1 bytes produced by [start, e0)
  uint32_t nblock;      // RLE1 bytes of the whole block
  unsigned long long gx0;  // global emitted-byte prefix at e0
};

struct Plan {
  size_t ws_bytes;      // device workspace
  uint32_t n_tiles;     // input tiles
  uint32_t max_blocks;  // capacity of the block table
  uint32_t batch;       // blocks sorted / coded together
};
Plan plan(size_t n, size_t mem_budget);
size_t bound(size_t n);

struct Stats {
  uint32_t n_blocks, n_serial_blocks, rounds, reserved;
};

// status: 0 ok, -3 out_cap too small (*out_len = bytes needed), -6 internal
int encode_device(const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, void *ws, const Plan &p, size_t *out_len,
                  Stats *stats, void *stream);

}  // namespace bz2e
}  // namespace b200z
