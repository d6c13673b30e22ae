// inflate_emul.cpp -- TEST INFRASTRUCTURE: k_inflate_decode / k_inflate_expand of archive_b200/csrc/inflate_kernels.cu
// executed on the CUDA execution-model emulation (cuda_emu.h), so that the warp-level logic -- several lanes per stream,
// speculative helpers, piece stitching, the expand kernel -- is covered by the CPU test tier.
#define B200Z_EMU 1
#include "../../archive_b200/csrc/inflate_kernels.cu"

using namespace b200z;
static std::vector<uint32_t> g_last_pieces;

// n units; returns 0.  upw = streams per warp, lpu = lanes per stream
extern "C" int emu_inflate_batch(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out_base,
                                 const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, int32_t *status,
                                 uint32_t *in_used, uint32_t n_units, size_t extent, int upw, int lpu, uint32_t *pieces_out) {
  InflateWs w;
  std::vector<uint32_t> tokens(extent + 256);
  w.hstride = (extent >> SPEC_HSHIFT) + 64;
  std::vector<uint32_t> htokens((SPEC_MAX_G - 1) * w.hstride + 64);
  std::vector<uint32_t> pieces((size_t)n_units * PIECE_WORDS + 16);
  std::vector<uint8_t> us((size_t)n_units * USCRATCH_BYTES + 256);
  w.tokens = tokens.data();
  w.htokens = htokens.data();
  w.pieces = pieces.data();
  w.uscratch = us.data();
  // the kernel reads aligned 16-byte blocks around every unit: give the input a padded home
  const uint32_t n_warps = (n_units + upw - 1) / upw;
  B200Z_LAUNCH(k_inflate_decode<false>, n_warps, 32, 0, 0, in_base, in_off, in_len, out_off, out_cap, w, out_len, status, in_used,
               n_units, upw, lpu, 0, 0);
  B200Z_LAUNCH(k_inflate_expand<false>, (n_units + 7) / 8, 256, 0, 0, w, in_base, in_off, out_base, out_off, out_cap, out_len, status,
               n_units, 0);
  g_last_pieces = pieces;
  if (pieces_out)
    for (uint32_t u = 0; u < n_units; ++u) pieces_out[u] = pieces[(size_t)u * PIECE_WORDS];
  return 0;
}

// debugging aid: the whole piece table of the last call
extern "C" const uint32_t *emu_last_pieces() { return g_last_pieces.data(); }

// k_inflate_fast (inflate_fast.cuh) alone: flags[u] = 1 where the unit was finished in shared memory, 0 where it was left to
// the exact kernels (then out_len / status / in_used of that unit are untouched).  `blocks` CTAs share the units.
extern "C" int emu_inflate_fast(const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out_base,
                                const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, int32_t *status,
                                uint32_t *in_used, uint32_t n_units, uint32_t blocks, uint32_t *flags) {
  uint32_t next_unit = 0;
  B200Z_LAUNCH(k_inflate_fast, blocks, fp::NTT, 0, 0, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, in_used,
               n_units, flags, 1u, &next_unit);
  return 0;
}
