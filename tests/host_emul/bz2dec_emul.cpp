// bz2dec_emul.cpp -- TEST INFRASTRUCTURE: the BZip2 decode kernels (K6 magic scan, K7 entropy decode, K8 inverse BWT /
// RLE / CRC) of archive_b200/csrc/bzip2_kernels.cu executed on the CUDA execution-model emulation, so that the CPU test
// tier covers them.  The kernels and their launch order are compiled from a generated copy of the product file
// (gen_emul.py: only the <<<>>> syntax, the dynamic shared memory declaration and a prefetch hint differ).
//
// The entry point has the shape of b200z_bzip2_decode_shard with world = 1: every block candidate is decoded and
// reported; the caller walks the chain with archive_b200/shard.py (bz2_walk_chain), the same host logic the multi-GPU
// path uses.
#include "cuda_emu.h"
#include "b200z_internal.h"
namespace b200z {
void count_launch() {}
}  // namespace b200z
#define BZ_SLOT_BYTES 64  // K8: most segments of the test blocks overflow their slot, so both halves of the walk are covered
#include "_gen/bzip2_kernels_emu.inc"

#include <algorithm>
#include <vector>

using namespace b200z;
static uint32_t g_last_quirk = 0, g_last_fast = 0;
extern "C" uint32_t emu_bzip2_last_fast() { return g_last_fast; }  // blocks of the last call that k_bz2_entropy_fast finished
extern "C" uint32_t emu_bzip2_last_quirk() { return g_last_quirk; }  // blocks of the last call that took the literal path

static uint32_t be32_at(const uint8_t *in, size_t n, uint64_t bit) {
  uint64_t v = 0;
  const size_t b0 = (size_t)(bit >> 3);
  for (int i = 0; i < 5; ++i) v = (v << 8) | (b0 + i < n ? in[b0 + i] : 0);
  return (uint32_t)(v >> (8 - (bit & 7)));
}

// -> 0, or -1 (bad signature / level: decodeStream false), -2 (truncated header: RangeError), -3 (out_cap / blocks_cap
// too small; *out_len / *n_blocks hold the need).  *empty = 1 when the stream is just the 4 header bytes.
extern "C" int emu_bzip2_blocks(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *out_len,
                                b200z_bz2_block *blocks, size_t blocks_cap, size_t *n_blocks, int *empty) {
  *out_len = 0;
  *n_blocks = 0;
  *empty = 0;
  static const uint8_t sig[3] = {0x42, 0x5a, 0x68};
  for (int i = 0; i < 3; ++i) {
    if ((size_t)i >= in_len) return -2;
    if (in[i] != sig[i]) return -1;
  }
  if (in_len < 4) return -2;
  const int level = (int)in[3] - 0x30;
  if (level < 0 || level > 9) return -1;
  if (in_len == 4) {
    *empty = 1;
    return 0;
  }
  const uint32_t nblock_max = (uint32_t)level * 100000u;
  const uint64_t total_bits = (uint64_t)in_len * 8;
  // the staged input: 4-byte aligned, zero padded (b200z_api.cu stage_input + the 64-byte memset)
  std::vector<uint32_t> staged((in_len + 64 + 3) / 4 + 80, 0u);
  memcpy(staged.data(), in, in_len);
  const uint8_t *d_in = (const uint8_t *)staged.data();

  const uint32_t cand_cap = 1u << 16;
  std::vector<unsigned long long> cand(cand_cap);
  uint32_t ncand = 0;
  if (bz2_launch_scan(d_in, in_len, cand.data(), &ncand, cand_cap, nullptr) != cudaSuccess) return -9;
  if (ncand > cand_cap) return -9;
  cand.resize(ncand);
  std::sort(cand.begin(), cand.end(),
            [](unsigned long long a, unsigned long long b) { return (a & ~(1ull << 63)) < (b & ~(1ull << 63)); });
  std::vector<unsigned long long> blk_bits;
  for (unsigned long long c : cand)
    if (!(c >> 63)) blk_bits.push_back(c);
  const uint32_t nb = (uint32_t)blk_bits.size();
  const uint32_t nbk = nb ? nb : 1;
  const uint32_t chunks_max = (nblock_max + 1023) / 1024;

  std::vector<uint32_t> rec_val((size_t)nbk * nblock_max), rec_pos((size_t)nbk * nblock_max), tt((size_t)nbk * nblock_max);
  std::vector<uint8_t> sym8((size_t)nbk * nblock_max), raw((size_t)nbk * nblock_max);
  std::vector<uint32_t> n_rec(nbk), nblock(nbk), orig_ptr(nbk), rnd(nbk), block_crc(nbk), cycle_len(nbk);
  std::vector<int32_t> status(nbk), irregular(nbk);
  std::vector<unsigned long long> end_bit(nbk), block_out(nbk), block_off(nbk + 1);
  std::vector<uint32_t> seg_len((size_t)nbk * 4098), seg_next((size_t)nbk * 4098), seg_off((size_t)nbk * 4098);
  std::vector<uint32_t> slice_state((size_t)nbk * 1024), slice_out((size_t)nbk * 1024);
  std::vector<uint32_t> chist((size_t)nbk * chunks_max * 256);

  if (nb) {
    Bz2Entropy e;
    e.words = staged.data();
    e.n_bytes = in_len;
    e.blk_bit = blk_bits.data();
    e.n_blocks = nb;
    e.nblock_max = nblock_max;
    e.rec_val = rec_val.data(); e.rec_pos = rec_pos.data(); e.n_rec = n_rec.data(); e.nblock = nblock.data();
    e.orig_ptr = orig_ptr.data(); e.randomised = rnd.data(); e.end_bit = end_bit.data(); e.status = status.data();
    std::vector<uint32_t> fastf(nb, 0u);
    e.fast_flag = fastf.data();
    e.sym8 = sym8.data();
    if (bz2_launch_entropy(e, nullptr) != cudaSuccess) return -9;
    g_last_fast = 0;
    for (uint32_t k = 0; k < nb; ++k) g_last_fast += fastf[k];
    std::vector<uint32_t> quirk;  // as in b200z_api.cu: damaged blocks the reference keeps decoding
    for (uint32_t k = 0; k < nb; ++k)
      if (status[k] == -3) quirk.push_back(k);
    g_last_quirk = (uint32_t)quirk.size();
    if (!quirk.empty() && bz2_launch_entropy_literal(e, quirk.data(), (uint32_t)quirk.size(), nullptr) != cudaSuccess) return -9;
  }
  std::vector<BzChainHost> chain;
  std::vector<uint32_t> chain_of(nb, 0xffffffffu);
  for (uint32_t k = 0; k < nb; ++k)
    if (status[k] == 0) {
      chain_of[k] = (uint32_t)chain.size();
      chain.push_back({k, nblock[k], n_rec[k], orig_ptr[k], rnd[k] ? 1u : 0u});
    }
  const uint32_t nc = (uint32_t)chain.size();
  std::vector<uint8_t> dout(out_cap + 64);
  if (nc) {
    Bz2Ibwt w;
    w.chain = chain.data(); w.n_chain = nc; w.nblock_max = nblock_max;
    w.rec_val = rec_val.data(); w.rec_pos = rec_pos.data(); w.sym8 = sym8.data(); w.chist = chist.data(); w.tt = tt.data();
    std::vector<uint32_t> walk_ctr(4, 0u), seg_resume((size_t)nbk * 4098);
    std::vector<uint8_t> slots((size_t)nbk * bz2_slot_bytes_per_block());
    w.walk_ctr = walk_ctr.data();
    w.seg_resume = seg_resume.data();
    w.slots = slots.data();
    w.seg_len = seg_len.data(); w.seg_next = seg_next.data(); w.seg_off = seg_off.data(); w.irregular = irregular.data();
    w.cycle_len = cycle_len.data(); w.raw = raw.data(); w.slice_state = slice_state.data(); w.slice_out = slice_out.data();
    w.block_out = block_out.data(); w.block_off = block_off.data(); w.block_crc = block_crc.data();
    w.out = dout.data(); w.out_cap = out_cap;
    for (const BzChainHost &ce : chain) w.any_randomised = w.any_randomised || (ce.flags & 1u);
    if (bz2_launch_ibwt(w, nullptr) != cudaSuccess) return -9;
  }
  size_t nrep = 0;
  auto push = [&](const b200z_bz2_block &b) {
    if (nrep < blocks_cap) blocks[nrep] = b;
    nrep++;
  };
  for (uint32_t k = 0; k < nb; ++k) {
    b200z_bz2_block b{};
    b.start_bit = blk_bits[k];
    b.end_bit = end_bit[k];
    b.status = status[k];
    b.crc_stored = blk_bits[k] + 80 <= total_bits ? be32_at(in, in_len, blk_bits[k] + 48) : 0u;
    const uint32_t c = chain_of[k];
    if (c != 0xffffffffu) {
      b.out_bytes = block_off[c + 1] - block_off[c];
      b.crc_calc = block_crc[c];
      if (irregular[c] == 2) b.flags |= B200Z_BZ2_OVERRUN;
      else if (irregular[c]) b.flags |= B200Z_BZ2_CORRUPT_CYCLE;
    }
    push(b);
  }
  for (unsigned long long c : cand)
    if (c >> 63) {
      b200z_bz2_block b{};
      b.start_bit = c & ~(1ull << 63);
      b.end_bit = b.start_bit + 80;
      b.flags = B200Z_BZ2_EOS;
      b.crc_stored = b.start_bit + 80 <= total_bits ? be32_at(in, in_len, b.start_bit + 48) : 0u;
      push(b);
    }
  *n_blocks = nrep;
  const size_t n_local = nc ? (size_t)block_off[nc] : 0;
  *out_len = n_local;
  if (nrep > blocks_cap || n_local > out_cap) return -3;
  if (n_local) memcpy(out, dout.data(), n_local);
  return 0;
}
