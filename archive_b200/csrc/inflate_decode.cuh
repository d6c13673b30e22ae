// inflate_decode.cuh -- per-stream DEFLATE decode logic shared by the sm_100a kernel
// (inflate_kernels.cu) and the host-side logic emulation used by the CPU tests.
//
// Restates (reference, paths relative to /root/reference/):
//   lib/src/codecs/zlib/inflate.dart:104-401, lib/src/codecs/zlib/_huffman_table.dart:9-46
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200z.h"

#ifndef B200Z_LBITS
#define B200Z_LBITS 9
#endif
#ifndef B200Z_DBITS
#define B200Z_DBITS 8
#endif

#ifdef __CUDA_ARCH__
// 32-bit shared-window addressing for the per-lane LUTs: keeps the hot loop free of 64-bit generic pointers
#define B200Z_SADDR(p) ((uint32_t)__cvta_generic_to_shared(p))
__device__ __forceinline__ uint32_t b200z_lds16(uint32_t base, uint32_t idx) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + idx * 2u));
  return v;
}
__device__ __forceinline__ uint32_t b200z_lds32(uint32_t base, uint32_t idx) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + idx * 4u));
  return v;
}
#define B200Z_LDS16(base, idx) b200z_lds16(base, idx)
#define B200Z_LDS32(base, idx) b200z_lds32(base, idx)
#define B200Z_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
typedef uint32_t b200z_saddr;
#define B200Z_ANY(x) __any_sync(0xffffffffu, (x))
#define B200Z_BALLOT(x) __ballot_sync(0xffffffffu, (x))
__device__ __forceinline__ uint32_t b200z_opaque(uint32_t v) {
  uint32_t o;
  asm volatile("mov.b32 %0, %1;" : "=r"(o) : "r"(v));  // keeps a loop invariant in a register (no rematerialisation)
  return o;
}
#define B200Z_OPAQUE(x) b200z_opaque(x)
#define B200Z_LDG(p) __ldg(p)
#define B200Z_BREV(x) __brev(x)
#else
#define B200Z_SADDR(p) (p)
#define B200Z_LDS16(base, idx) ((uint32_t)((const uint16_t *)(base))[idx])
#define B200Z_LDS32(base, idx) (((const uint32_t *)(base))[idx])
#define B200Z_PREFETCH(p) ((void)0)
typedef const void *b200z_saddr;
#define B200Z_ANY(x) (x)
#define B200Z_BALLOT(x) ((x) ? 1u : 0u)
#define B200Z_OPAQUE(x) (x)
#define B200Z_LDG(p) (*(p))
static inline uint32_t b200z_host_brev(uint32_t v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
}
#define B200Z_BREV(x) b200z_host_brev(x)
#endif
#ifdef __CUDACC__
#define B200Z_HD __device__ __forceinline__
#define B200Z_CONST __constant__
#else
#define B200Z_HD inline
#define B200Z_CONST static const
#include <algorithm>
using std::max;
using std::min;
#endif

namespace b200z {

// ---------------------------------------------------------------------------------------------
// constants (inflate.dart:738-894)
// ---------------------------------------------------------------------------------------------
B200Z_CONST uint8_t c_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
// length symbol 257+i -> (base << 4) | extra_bits
B200Z_CONST uint16_t c_len_tab[32] = {
    (3 << 4) | 0,   (4 << 4) | 0,   (5 << 4) | 0,   (6 << 4) | 0,   (7 << 4) | 0,   (8 << 4) | 0,
    (9 << 4) | 0,   (10 << 4) | 0,  (11 << 4) | 1,  (13 << 4) | 1,  (15 << 4) | 1,  (17 << 4) | 1,
    (19 << 4) | 2,  (23 << 4) | 2,  (27 << 4) | 2,  (31 << 4) | 2,  (35 << 4) | 3,  (43 << 4) | 3,
    (51 << 4) | 3,  (59 << 4) | 3,  (67 << 4) | 4,  (83 << 4) | 4,  (99 << 4) | 4,  (115 << 4) | 4,
    (131 << 4) | 5, (163 << 4) | 5, (195 << 4) | 5, (227 << 4) | 5, (258 << 4) | 0, 0, 0, 0};
// distance symbol -> (base << 4) | extra_bits
B200Z_CONST uint32_t c_dist_tab[32] = {
    (1 << 4) | 0,     (2 << 4) | 0,     (3 << 4) | 0,      (4 << 4) | 0,      (5 << 4) | 1,      (7 << 4) | 1,
    (9 << 4) | 2,     (13 << 4) | 2,    (17 << 4) | 3,     (25 << 4) | 3,     (33 << 4) | 4,     (49 << 4) | 4,
    (65 << 4) | 5,    (97 << 4) | 5,    (129 << 4) | 6,    (193 << 4) | 6,    (257 << 4) | 7,    (385 << 4) | 7,
    (513 << 4) | 8,   (769 << 4) | 8,   (1025 << 4) | 9,   (1537 << 4) | 9,   (2049 << 4) | 10,  (3073 << 4) | 10,
    (4097 << 4) | 11, (6145 << 4) | 11, (8193 << 4) | 12,  (12289 << 4) | 12, (16385 << 4) | 13, (24577 << 4) | 13,
    0, 0};

// token encoding (uint32):
//   literal : 0x80000000 | byte
//   match   : (len << 16) | dist          len 1..258 (bit 31/30 clear), dist 1..32768
//   stored  : 0x40000000 | (pos >> 30) << 16 | len (3..65535), followed by ONE payload word =
//             pos & 0x3fffffff (pos = byte offset of the run in the unit's input; top bits 00 so a
//             payload never looks like a stored token); the pair never straddles a group of 32
//             tokens (a nop pads).
//   nop     : 0
#define TOK_LIT 0x80000000u
#define TOK_STORED 0x40000000u

constexpr int LBITS = B200Z_LBITS;  // primary literal/length LUT bits
constexpr int DBITS = B200Z_DBITS;  // primary distance LUT bits
constexpr int LUT_HALFWORDS = (1 << LBITS) + (1 << DBITS);
constexpr int LANE_STRIDE_WORDS = LUT_HALFWORDS / 2 + 1;  // +1 word: same index -> different bank per lane
constexpr int CONST_WORDS = 16 + 32 + 64;                  // len table (32 x u16) + dist table (32 x u32) + xtab (64 x u32)

static inline size_t inflate_decode_smem_bytes(int warps_per_block, int units_per_warp) {
  return (size_t)(CONST_WORDS + warps_per_block * units_per_warp * LANE_STRIDE_WORDS) * 4;
}

// Canonical-code side tables for codes longer than the LUT (rare): per lane, in local memory.
struct SlowTab {
  uint16_t first[16];  // first canonical code of each length
  uint16_t count[16];  // number of codes of each length
  uint16_t offs[16];   // index into perm of the first symbol of each length
  uint16_t perm[288];  // symbols sorted by (length, symbol)
  uint8_t maxlen;      // HuffmanTable.maxCodeLength (_huffman_table.dart:12-15)
};
struct SlowTabD {
  uint16_t first[16];
  uint16_t count[16];
  uint16_t offs[16];
  uint8_t perm[32];
  uint8_t maxlen;
};

// ---------------------------------------------------------------------------------------------
// bit reader: LSB-first (inflate.dart:159-184), refilled 32 aligned bits at a time.
//   rem_bits() is the exact number of stream bits not yet consumed; it is what the reference's
//   "isEOS while _bitBufferLen < n" tests (inflate.dart:166-168,192-195) see.
// ---------------------------------------------------------------------------------------------
struct BitReader {
  const uint32_t *w;  // 16-byte aligned word base of the unit
  uint32_t nextw;  // word `widx`, requested one refill ahead of its use (hides the L1/L2 latency)
  uint64_t buf;
  int cnt;          // bits in buf (may include `pad` invalid bits once widx >= nw)
  uint32_t widx;    // next word to load
  uint32_t nw;      // words covering the unit
  uint32_t lead;    // byte offset of the unit inside word 0
  uint32_t in_len;  // unit bytes

  B200Z_HD void seek(uint32_t byte_pos) {
    uint32_t a = lead + byte_pos;
    widx = a >> 2;
    uint32_t sh = (a & 3) * 8;
    uint32_t v = (widx < nw) ? B200Z_LDG(w + widx) : 0u;
    widx++;
    nextw = (widx < nw) ? B200Z_LDG(w + widx) : 0u;
    buf = (uint64_t)(v >> sh);
    cnt = 32 - (int)sh;
  }
  B200Z_HD void refill() {
    if (cnt < 32) {
      buf |= (uint64_t)nextw << cnt;
      cnt += 32;
      widx++;
      nextw = (widx < nw) ? B200Z_LDG(w + widx) : 0u;
    }
  }
  // all bits in buf valid and >= 32 of them after refill()
  B200Z_HD bool fast() const { return widx < nw; }
  B200Z_HD long long rem_bits() const {
    return (long long)cnt + 32ll * ((long long)nw - (long long)widx) -
           (32ll * nw - 8ll * ((long long)lead + in_len));
  }
  B200Z_HD uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
  B200Z_HD void drop(int n) {
    buf >>= n;
    cnt -= n;
  }
  // _readBits: -1 when fewer than n bits remain (then nothing is consumed that matters)
  B200Z_HD int read_bits_checked(int n) {
    if (n == 0) return 0;
    refill();
    if (!fast() && rem_bits() < n) return -1;
    int v = (int)peek(n);
    drop(n);
    return v;
  }
};

// ---------------------------------------------------------------------------------------------
// Build the LUT + slow tables for one alphabet from code lengths (HuffmanTable ctor restated for a
// two-level layout).  Returns false when the set is over-subscribed (reference: later writes win in
// a flat table -- garbage; here: B200Z_U_BADCODE).
// ---------------------------------------------------------------------------------------------
template <int TBITS, typename PermT>
B200Z_HD bool build_table(const uint8_t *lens, int n, uint16_t *lut, uint16_t *first,
                                            uint16_t *count, uint16_t *offs, PermT *perm, uint8_t *maxlen,
                                            int lut_skip_eq = -1, int lut_skip_above = 0x7fffffff) {
  for (int l = 0; l < 16; ++l) count[l] = 0;
  int mx = 0;
  for (int i = 0; i < n; ++i) {
    int l = lens[i];
    count[l]++;
    mx = max(mx, l);
  }
  *maxlen = (uint8_t)mx;
  count[0] = 0;
  // Kraft check
  int left = 1;
  bool over = false;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= count[l];
    if (left < 0) over = true;
  }
  uint16_t next[16];
  {
    int code = 0, o = 0;
    for (int l = 1; l < 16; ++l) {
      code = (code + count[l - 1]) << 1;
      first[l] = (uint16_t)code;
      next[l] = (uint16_t)code;
      offs[l] = (uint16_t)o;
      o += count[l];
    }
    first[0] = 0;
    offs[0] = 0;
  }
  uint32_t *lut32 = reinterpret_cast<uint32_t *>(lut);
  for (int i = 0; i < (1 << TBITS) / 2; ++i) lut32[i] = 0;
  if (over) return false;
  uint16_t run[16];
  for (int l = 0; l < 16; ++l) run[l] = offs[l];
  for (int s = 0; s < n; ++s) {
    int l = lens[s];
    if (l == 0) continue;
    uint32_t c = next[l]++;
    perm[run[l]++] = (PermT)s;
    if (l <= TBITS && s != lut_skip_eq && s <= lut_skip_above) {
      uint32_t r = B200Z_BREV(c) >> (32 - l);
      uint16_t e = (uint16_t)((s << 4) | l);
      for (uint32_t j = r; j < (1u << TBITS); j += (1u << l)) lut[j] = e;
    }
  }
  return true;
}

// canonical decode of a code longer than TBITS (or a hole).  Returns length, 0 = no code matches.
template <int TBITS, typename PermT>
B200Z_HD int slow_decode(uint32_t bits15, const uint16_t *first, const uint16_t *count,
                                           const uint16_t *offs, const PermT *perm, int maxlen, int *sym) {
  uint32_t rev = B200Z_BREV(bits15) >> 17;  // first stream bit = MSB of a 15-bit value
  for (int l = 1; l <= maxlen; ++l) {
    uint32_t code = rev >> (15 - l);
    uint32_t d = code - first[l];
    if (d < count[l]) {
      *sym = perm[offs[l] + d];
      return l;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One stream: DEFLATE bits -> token stream.  Runs as one LANE of k_inflate_decode (and, compiled as
// plain C++, inside tests/host_emul to check the logic against the oracle without a GPU).
// ---------------------------------------------------------------------------------------------
struct UnitResult {
  uint32_t ntok, out_len, in_used;
  int32_t status;
};

B200Z_HD UnitResult inflate_decode_unit(bool active, const uint8_t *in, uint32_t in_len, uint32_t cap, uint32_t *tok,
                                        uint16_t *lut_l, uint16_t *lut_d, const uint16_t *s_len_tab,
                                        const uint32_t *s_dist_tab, const uint32_t *s_xtab) {
  const b200z_saddr lutl_s = B200Z_SADDR(lut_l), lutd_s = B200Z_SADDR(lut_d), xtab_s = B200Z_SADDR(s_xtab);
  SlowTab sl;
  SlowTabD sd;
  uint8_t lens[320];

  BitReader br;
  {
    uintptr_t a = reinterpret_cast<uintptr_t>(in);
    br.lead = (uint32_t)(a & 15);
    br.w = reinterpret_cast<const uint32_t *>(a - br.lead);
    br.in_len = in_len;
    br.nw = (uint32_t)(((uint64_t)br.lead + br.in_len + 3) >> 2);
    br.seek(0);
  }
  uint32_t nt = 0;
  uint32_t olen = 0;
  int st = B200Z_U_EOS;
  bool in_block = false;
  bool final_block = false;
  int maxl = 0, maxd = 0;
  bool mode_dist = false;  // bulk path: the next symbol is a distance code
  uint32_t mlen_pending = 0;

  bool done = !active;
  // loop invariants of the bulk loop, pinned in registers
  const b200z_saddr lutl_r = B200Z_OPAQUE(lutl_s), lutd_r = B200Z_OPAQUE(lutd_s), xtab_r = B200Z_OPAQUE(xtab_s);
  unsigned live;
  while ((live = B200Z_BALLOT(!done)) != 0u) {  // warp-uniform: every lane reconverges here
    // ---------------- bulk inner loop: warp-uniform, ONE SYMBOL per lane per turn, branch-light.  The same
    // instructions decode a literal/length symbol or a distance symbol (a lane that has just read a length
    // code reads its distance code on the next turn), so literal lanes and match lanes do not diverge.  A
    // lane speculates the symbol from two table look-ups and commits only if nothing special happened:
    // LUT miss (long code, end-of-block and invalid symbols are deliberately absent from the LUT),
    // back-reference before the start, output full, or fewer than 64 unloaded bits left.  Anything special
    // drops the warp to the exact step below for one turn.  (A token is <= 48 bits, so with >= 64 unloaded
    // bits at its start no end-of-stream test is needed in here.)
    bool fast_ok = !done && in_block;
    for (;;) {
      const bool can = fast_ok && (mode_dist || br.widx + 2u <= br.nw);
      if (B200Z_BALLOT(can) != live) break;
      if (!can) continue;  // lanes whose stream is finished just keep voting
      br.refill();
      const bool dm = mode_dist;
      const uint32_t bits = (uint32_t)br.buf;
      const uint32_t e = B200Z_LDS16(dm ? lutd_r : lutl_r, bits & (dm ? ((1u << DBITS) - 1u) : ((1u << LBITS) - 1u)));
      uint32_t n = e & 15u;
      uint32_t sym = e >> 4;
      bool odd = false;
      if (n == 0u) {
        // LUT miss (rare, divergent): a code longer than the LUT, or one of the symbols kept out of it.  Long codes
        // are decoded here by the canonical walk; end-of-block / invalid symbols / holes go to the exact step.
        int sy = 0;
        const int ln = dm ? slow_decode<DBITS, uint8_t>(bits & 0x7fffu, sd.first, sd.count, sd.offs, sd.perm, maxd, &sy)
                          : slow_decode<LBITS, uint16_t>(bits & 0x7fffu, sl.first, sl.count, sl.offs, sl.perm, maxl, &sy);
        n = (uint32_t)ln;
        sym = (uint32_t)sy;
        odd = ln == 0 || (dm ? sy > 29 : (sy == 256 || sy > 285));
      }
      const uint32_t xi = dm ? sym + 32u : (sym > 256u ? sym - 257u : 63u);
      const uint32_t x = B200Z_LDS32(xtab_r, xi & 63u);
      const uint32_t xb = x & 15u;
      const uint32_t val = (x >> 4) + ((bits >> n) & ~(0xffffffffu << xb));
      const bool islit = !dm && sym < 256u;
      const bool islen = !dm && sym > 256u;
      const uint32_t nolen = olen + (islit ? 1u : dm ? mlen_pending : 0u);
      const bool special = odd || (dm && val > olen) || nolen > cap;
      if (!special) {
        const uint32_t tot = n + xb;
        br.buf >>= tot;
        br.cnt -= (int)tot;
        if (islit || dm) tok[nt++] = islit ? (TOK_LIT | sym) : ((mlen_pending << 16) | val);
        olen = nolen;
        mlen_pending = islen ? val : mlen_pending;
        mode_dist = islen;
      }
      fast_ok = !special;
    }
    if (!done) do {
    if (in_block && (mode_dist || br.widx + 2u <= br.nw)) {
      // ---------------- bulk path: ONE SYMBOL per turn, the same instructions for literal/length and
      // distance symbols (a lane that has just read a length code reads its distance code on the next
      // turn), so the lanes of a warp stay converged.  >= 64 stream bits are still unloaded when a
      // literal/length symbol starts, so no end-of-stream test is needed here (a token is <= 48 bits);
      // everything near the end of the stream goes through the exact per-token path below.
      br.refill();
      const bool dm = mode_dist;
      const uint32_t bits = (uint32_t)br.buf;
      uint32_t e = B200Z_LDS16(dm ? lutd_s : lutl_s, bits & (dm ? ((1u << DBITS) - 1u) : ((1u << LBITS) - 1u)));
      int n = (int)(e & 15u);
      int sym = (int)(e >> 4);
      if (n == 0) {
        if (dm) {
          n = slow_decode<DBITS, uint8_t>(bits & 0x7fffu, sd.first, sd.count, sd.offs, sd.perm, maxd, &sym);
          if (n == 0) sym = 0;  // hole in the flat table: (len 0, sym 0) (_huffman_table.dart:22)
        } else {
          n = slow_decode<LBITS, uint16_t>(bits & 0x7fffu, sl.first, sl.count, sl.offs, sl.perm, maxl, &sym);
          if (n == 0) {
            st = B200Z_U_BADCODE;
            done = true; break;
          }
        }
      }
      br.drop(n);
      // base + extra bits: lengths at [0,32), distances at [32,64); literals/EOB read the all-zero entry 63
      const uint32_t xi = dm ? 32u + (uint32_t)sym : (sym > 256 ? (uint32_t)(sym - 257) : 63u);
      const uint32_t x = B200Z_LDS32(xtab_s, xi & 63u);
      const int xb = (int)(x & 15u);
      const uint32_t val = (x >> 4) + ((uint32_t)br.buf & ((1u << xb) - 1u));
      br.drop(xb);
      if (!dm) {
        if (sym < 256) {
          if (olen >= cap) {
            st = B200Z_U_NOSPC;
            done = true; break;
          }
          tok[nt++] = TOK_LIT | (uint32_t)sym;
          olen++;
        } else if (sym == 256) {
          in_block = false;
        } else if (sym > 285) {
          st = B200Z_U_STOP;
          done = true; break;
        } else {
          mlen_pending = val;
          mode_dist = true;
        }
      } else {
        mode_dist = false;
        if (sym > 29) {
          st = B200Z_U_STOP;
          done = true; break;
        }
        if (val > olen) {  // writeBackReference before the start of the output (output_memory_stream.dart:83-86)
          st = B200Z_U_RANGE;
          done = true; break;
        }
        if (olen + mlen_pending > cap) {
          st = B200Z_U_NOSPC;
          done = true; break;
        }
        tok[nt++] = (mlen_pending << 16) | val;
        olen += mlen_pending;
      }
      break;  // next symbol
    }
    if (!in_block) {
      // ---------------- block boundary: _inflate loop + _parseBlock (inflate.dart:111-156) -------------
      if (final_block) {
        st = B200Z_U_DONE;
        done = true; break;
      }
      br.refill();
      if (br.rem_bits() < 8) {  // isEOS: every byte already pulled into the bit buffer
        st = B200Z_U_EOS;
        done = true; break;
      }
      uint32_t hdr = br.peek(3);
      br.drop(3);
      final_block = hdr & 1;
      uint32_t type = hdr >> 1;
      if (type == 0) {
        // ---- stored (inflate.dart:213-235) ----
        int k = (int)(br.rem_bits() & 7);
        br.drop(k);
        long long rem_bytes = br.rem_bits() >> 3;
        uint32_t pos = br.in_len - (uint32_t)rem_bytes;
        long long len = -1, nlen;
        if (rem_bytes >= 2) {
          br.refill();
          len = br.peek(16);
          br.drop(16);
          rem_bytes -= 2;
          pos += 2;
        } else {
          rem_bytes = 0;  // the short read swallowed what was left
          pos = br.in_len;
        }
        if (rem_bytes >= 2) {
          br.refill();
          nlen = (long long)br.peek(16) ^ 0xffff;
          br.drop(16);
          rem_bytes -= 2;
          pos += 2;
        } else {
          nlen = -1ll ^ 0xffff;
          rem_bytes = 0;
          pos = br.in_len;
        }
        if (len != 0 && len != nlen) {
          st = B200Z_U_STOP;
          done = true; break;
        }
        if (len > rem_bytes) {
          st = B200Z_U_STOP;
          done = true; break;
        }
        if (len > 0) {
          if ((unsigned long long)olen + len > cap) {
            st = B200Z_U_NOSPC;
            done = true; break;
          }
          if (len < 3) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(br.w) + br.lead + pos;
            for (int i = 0; i < (int)len; ++i) tok[nt++] = TOK_LIT | src[i];
          } else {
            if ((nt & 31) == 31) tok[nt++] = 0;
            tok[nt++] = TOK_STORED | ((pos >> 30) << 16) | (uint32_t)len;
            tok[nt++] = pos & 0x3fffffffu;
          }
          olen += (uint32_t)len;
        }
        br.seek(pos + (uint32_t)len);
        break;  // next token
      } else if (type == 1) {
        // ---- fixed tables (inflate.dart:408-735): 288 lit/len lengths, 30 distance lengths ----
        for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        build_table<LBITS, uint16_t>(lens, 288, lut_l, sl.first, sl.count, sl.offs, sl.perm, &sl.maxlen, 256, 285);
        for (int i = 0; i < 30; ++i) lens[i] = 5;
        build_table<DBITS, uint8_t>(lens, 30, lut_d, sd.first, sd.count, sd.offs, sd.perm, &sd.maxlen, -1, 29);
      } else if (type == 2) {
        // ---- dynamic (inflate.dart:239-298) ----
        int hlit = br.read_bits_checked(5);
        if (hlit < 0) { st = B200Z_U_STOP; done = true; break; }
        hlit += 257;
        if (hlit > 288) { st = B200Z_U_STOP; done = true; break; }
        int hdist = br.read_bits_checked(5);
        if (hdist < 0) { st = B200Z_U_STOP; done = true; break; }
        hdist += 1;
        if (hdist > 32) { st = B200Z_U_STOP; done = true; break; }
        int hclen = br.read_bits_checked(4);
        if (hclen < 0) { st = B200Z_U_STOP; done = true; break; }
        hclen += 4;
        if (hclen > 19) { st = B200Z_U_STOP; done = true; break; }
        for (int i = 0; i < 19; ++i) lens[i] = 0;
        bool bad = false;
        for (int i = 0; i < hclen; ++i) {
          int l = br.read_bits_checked(3);
          if (l < 0) { bad = true; break; }
          lens[c_order[i]] = (uint8_t)l;
        }
        if (bad) { st = B200Z_U_STOP; done = true; break; }
        // code-length alphabet: 7-bit LUT in the (not yet built) lit/len LUT area
        uint8_t clmax;
        {
          uint16_t f[16], c[16], o[16];
          uint8_t pm[19];
          if (!build_table<7, uint8_t>(lens, 19, lut_l, f, c, o, pm, &clmax)) { st = B200Z_U_BADCODE; done = true; break; }
        }
        // _decode (inflate.dart:345-401)
        const int num = hlit + hdist;
        int i = 0, prev = 0;
        int err = 0;
        while (i < num) {
          br.refill();
          if (!br.fast() && br.rem_bits() < clmax) { err = B200Z_U_STOP; break; }
          uint32_t e = lut_l[br.peek(7)];
          int l = e & 15;
          int code = e >> 4;
          // l == 0: hole in an incomplete set -- the reference's flat table yields (len 0, sym 0)
          // (_huffman_table.dart:22), i.e. a zero length for this symbol and no bits consumed.
          br.drop(l);
          int repeat;
          if (code < 16) {
            lens[i++] = (uint8_t)code;
            prev = code;
            continue;
          } else if (code == 16) {
            repeat = br.read_bits_checked(2);
            if (repeat < 0) { err = B200Z_U_STOP; break; }
            repeat += 3;
          } else if (code == 17) {
            repeat = br.read_bits_checked(3);
            if (repeat < 0) { err = B200Z_U_STOP; break; }
            repeat += 3;
            prev = 0;
          } else {
            repeat = br.read_bits_checked(7);
            if (repeat < 0) { err = B200Z_U_STOP; break; }
            repeat += 11;
            prev = 0;
          }
          if (i + repeat > num) { err = B200Z_U_THROW; break; }
          for (int k = 0; k < repeat; ++k) lens[i++] = (uint8_t)prev;
        }
        if (err) { st = err; done = true; break; }
        for (int k = hdist; k < 32; ++k) lens[hlit + k] = 0;
        bool ok = build_table<DBITS, uint8_t>(lens + hlit, hdist, lut_d, sd.first, sd.count, sd.offs, sd.perm, &sd.maxlen, -1, 29);
        ok = build_table<LBITS, uint16_t>(lens, hlit, lut_l, sl.first, sl.count, sl.offs, sl.perm, &sl.maxlen, 256, 285) && ok;
        if (!ok) { st = B200Z_U_BADCODE; done = true; break; }
      } else {
        st = B200Z_U_STOP;
        done = true; break;
      }
      maxl = sl.maxlen;
      maxd = sd.maxlen;
      in_block = true;
    }

    // ---------------- one token: _decodeHuffman (inflate.dart:300-343) -------------------------------
    br.refill();
    const bool careful = !br.fast();
    if (careful && br.rem_bits() < maxl) {  // _readCodeByTable short read (quirk Q1)
      st = B200Z_U_STOP;
      done = true; break;
    }
    uint32_t e = lut_l[br.peek(LBITS)];
    int n = e & 15;
    int sym = e >> 4;
    if (n == 0) {
      n = slow_decode<LBITS, uint16_t>(br.peek(15), sl.first, sl.count, sl.offs, sl.perm, maxl, &sym);
      if (n == 0) {  // hole: reference would emit literal 0 for ever (or until OOM)
        st = B200Z_U_BADCODE;
        done = true; break;
      }
    }
    br.drop(n);
    if (sym < 256) {
      if (olen >= cap) {
        st = B200Z_U_NOSPC;
        done = true; break;
      }
      tok[nt++] = TOK_LIT | (uint32_t)sym;
      olen++;
      break;  // next token
    }
    if (sym == 256) {
      in_block = false;
      break;  // next token
    }
    if (sym > 285) {
      st = B200Z_U_STOP;
      done = true; break;
    }
    uint32_t le = s_len_tab[sym - 257];
    int lx = le & 15;
    int mlen = (int)(le >> 4);
    if (!careful) {
      mlen += (int)br.peek(lx);
      br.drop(lx);
    } else {
      int x = 0;
      if (lx) {
        if (br.rem_bits() < lx) x = -1;  // _readBits -> -1 is ADDED to the base (inflate.dart:323)
        else { x = (int)br.peek(lx); br.drop(lx); }
      }
      mlen += x;
    }
    br.refill();
    const bool careful2 = !br.fast();
    if (careful2 && br.rem_bits() < maxd) {
      st = B200Z_U_STOP;
      done = true; break;
    }
    uint32_t de = lut_d[br.peek(DBITS)];
    int dn = de & 15;
    int dsym = de >> 4;
    if (dn == 0) {
      dn = slow_decode<DBITS, uint8_t>(br.peek(15), sd.first, sd.count, sd.offs, sd.perm, maxd, &dsym);
      if (dn == 0) dsym = 0;  // hole in the flat table: (len 0, sym 0) (_huffman_table.dart:22)
    }
    br.drop(dn);
    if (dsym > 29) {
      st = B200Z_U_STOP;
      done = true; break;
    }
    uint32_t dd = s_dist_tab[dsym];
    int dx = dd & 15;
    int dist = (int)(dd >> 4);
    if (!careful2) {
      dist += (int)br.peek(dx);
      br.drop(dx);
    } else {
      int x = 0;
      if (dx) {
        if (br.rem_bits() < dx) x = -1;
        else { x = (int)br.peek(dx); br.drop(dx); }
      }
      dist += x;
    }
    // writeBackReference (output_memory_stream.dart:79-98)
    if (dist <= 0 || (uint32_t)dist > olen) {  // dist 0 only via a truncated extra-bits read
      st = B200Z_U_RANGE;
      done = true; break;
    }
    if (olen + (uint32_t)mlen > cap) {
      st = B200Z_U_NOSPC;
      done = true; break;
    }
    tok[nt++] = ((uint32_t)mlen << 16) | (uint32_t)dist;
    olen += (uint32_t)mlen;
    } while (0);
  }

  UnitResult r;
  r.ntok = nt;
  r.out_len = olen;
  r.status = st;
  {
    long long rem = br.rem_bits();
    if (rem < 0) rem = 0;
    r.in_used = br.in_len - (uint32_t)(rem >> 3);  // whole unread bytes are given back (inflate.dart:337-340)
  }
  return r;
}

}  // namespace b200z
