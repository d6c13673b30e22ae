/*
 * oracle/bzip2_enc.c -- CPU ORACLE (test infrastructure only; see orc.h).
 *
 * Restates lib/src/codecs/bzip2_encoder.dart (BZip2Encoder: always BZh9, never randomised) and
 * lib/src/codecs/bzip2/bz2_bit_writer.dart:
 *   encodeStream :25-81, _writeBlock :83-110, _addCharToBlock/_addPairToBlock :2013-2071 (RLE1 + CRC),
 *   _blockSort :880-928 with _mainSort :1247-1503 / _mainQSort3 :1505-1697 / _mainSimpleSort :1699-1787 /
 *   _mainGtU :1789-2011 and _fallbackSort :930-1073 / _fallbackQSort3 :1075-1217 / _fallbackSimpleSort :1219-1245,
 *   _generateMTFValues :139-265, _sendMTFValues :267-745, _hbMakeCodeLengths :747-864, _hbAssignCodes :866-878.
 *
 * PARITY UNPINNED by the reference's own tests (round trip only); cross-checked byte-for-byte against
 * libbz2 (bz2.compress(x, 9)), the library the Dart code was derived from (tests/test_oracle_codecs.py).
 */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

#define BZ_N_RADIX 2
#define BZ_N_QSORT 12
#define BZ_N_SHELL 18
#define BZ_N_OVERSHOOT (BZ_N_RADIX + BZ_N_QSORT + BZ_N_SHELL + 2)
#define BZ_MAX_ALPHA 258
#define BZ_N_GROUPS 6
#define BZ_G_SIZE 50
#define BZ_N_ITERS 4
#define BZ_MAX_SELECTORS (2 + (900000 / BZ_G_SIZE))
#define BZ_RUNA 0
#define BZ_RUNB 1

typedef struct {
  orc_oms *out;
  int bit_buffer, bit_pos; /* Bz2BitWriter: _bitPos starts at 8 */
  const uint8_t *in;
  size_t in_len, in_pos;
  int32_t nblock, nblock_max, work_factor, budget, orig_ptr, n_in_use, n_mtf;
  uint32_t block_crc;
  int state_in_ch, state_in_len;
  uint32_t *arr1, *arr2, *ftab;
  uint8_t *block;
  uint16_t *mtfv;
  uint8_t in_use[256], unseq_to_seq[256];
  uint8_t *selector, *selector_mtf;
  uint8_t len[BZ_N_GROUPS][BZ_MAX_ALPHA];
  int32_t code[BZ_N_GROUPS][BZ_MAX_ALPHA], rfreq[BZ_N_GROUPS][BZ_MAX_ALPHA], mtf_freq[BZ_MAX_ALPHA];
  uint32_t len_pack[BZ_MAX_ALPHA][4];
} bzenc;

/* ---- Bz2BitWriter.writeBits(numBits, value) bz2_bit_writer.dart:29-68 (MSB first) ---- */
static void bw_bits(bzenc *s, int nbits, uint32_t value) {
  while (nbits > 0) {
    nbits--;
    int b = (value >> nbits) & 1;
    s->bit_buffer = (s->bit_buffer << 1) | b;
    s->bit_pos--;
    if (s->bit_pos == 0) {
      orc_oms_write_byte(s->out, s->bit_buffer);
      s->bit_pos = 8;
      s->bit_buffer = 0;
    }
  }
}
static void bw_flush(bzenc *s) {
  if (s->bit_pos != 8) bw_bits(s, s->bit_pos, 0);
}

static uint32_t crc_tab[256];
static void crc_init(void) {
  if (crc_tab[1]) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i << 24;
    for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : c << 1;
    crc_tab[i] = c;
  }
}
static inline uint32_t upd_crc(int v, uint32_t crc) { return (crc << 8) ^ crc_tab[((crc >> 24) ^ v) & 0xff]; }

/* ---- RLE1 front end (:2013-2071) ---- */
static void add_pair_to_block(bzenc *s) {
  uint8_t ch = (uint8_t)s->state_in_ch;
  for (int i = 0; i < s->state_in_len; i++) s->block_crc = upd_crc(ch, s->block_crc);
  s->in_use[s->state_in_ch] = 1;
  switch (s->state_in_len) {
    case 1:
      s->block[s->nblock++] = ch;
      break;
    case 2:
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      break;
    case 3:
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      break;
    default:
      s->in_use[s->state_in_len - 4] = 1;
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = ch;
      s->block[s->nblock++] = (uint8_t)(s->state_in_len - 4);
      break;
  }
}
static void add_char_to_block(bzenc *s, int b) {
  if (b != s->state_in_ch && s->state_in_len == 1) {
    s->block_crc = upd_crc(s->state_in_ch, s->block_crc);
    s->in_use[s->state_in_ch] = 1;
    s->block[s->nblock++] = (uint8_t)s->state_in_ch;
    s->state_in_ch = b;
  } else {
    if (b != s->state_in_ch || s->state_in_len == 255) {
      if (s->state_in_ch < 256) add_pair_to_block(s);
      s->state_in_ch = b;
      s->state_in_len = 1;
    } else {
      s->state_in_len++;
    }
  }
}

/* ---- fallback sort (:930-1245) ---- */
static void fallback_simple_sort(uint32_t *fmap, uint32_t *eclass, int32_t lo, int32_t hi) {
  int32_t i, j, tmp;
  uint32_t ec_tmp;
  if (lo == hi) return;
  if (hi - lo > 3) {
    for (i = hi - 4; i >= lo; i--) {
      tmp = (int32_t)fmap[i];
      ec_tmp = eclass[tmp];
      for (j = i + 4; j <= hi && ec_tmp > eclass[fmap[j]]; j += 4) fmap[j - 4] = fmap[j];
      fmap[j - 4] = (uint32_t)tmp;
    }
  }
  for (i = hi - 1; i >= lo; i--) {
    tmp = (int32_t)fmap[i];
    ec_tmp = eclass[tmp];
    for (j = i + 1; j <= hi && ec_tmp > eclass[fmap[j]]; j++) fmap[j - 1] = fmap[j];
    fmap[j - 1] = (uint32_t)tmp;
  }
}
#define FSWAP(a, b) { uint32_t t_ = (a); (a) = (b); (b) = t_; }
static int fallback_qsort3(uint32_t *fmap, uint32_t *eclass, int32_t lo_st, int32_t hi_st) {
  int32_t stack_lo[100], stack_hi[100];
  int32_t sp = 0, un_lo, un_hi, lt_lo, gt_hi, n, m, lo, hi;
  int64_t med;
  uint32_t r = 0, r3;
  stack_lo[sp] = lo_st; stack_hi[sp] = hi_st; sp++;
  while (sp > 0) {
    if (sp >= 100 - 1) return 0;
    sp--;
    lo = stack_lo[sp];
    hi = stack_hi[sp];
    if (hi - lo < 10) {
      fallback_simple_sort(fmap, eclass, lo, hi);
      continue;
    }
    r = ((r * 7621) + 1) % 32768;
    r3 = r % 3;
    if (r3 == 0) med = eclass[fmap[lo]];
    else if (r3 == 1) med = eclass[fmap[(lo + hi) >> 1]];
    else med = eclass[fmap[hi]];
    un_lo = lt_lo = lo;
    un_hi = gt_hi = hi;
    for (;;) {
      for (;;) {
        if (un_lo > un_hi) break;
        int64_t nn = (int64_t)eclass[fmap[un_lo]] - med;
        if (nn == 0) {
          FSWAP(fmap[un_lo], fmap[lt_lo]);
          lt_lo++;
          un_lo++;
          continue;
        }
        if (nn > 0) break;
        un_lo++;
      }
      for (;;) {
        if (un_lo > un_hi) break;
        int64_t nn = (int64_t)eclass[fmap[un_hi]] - med;
        if (nn == 0) {
          FSWAP(fmap[un_hi], fmap[gt_hi]);
          gt_hi--;
          un_hi--;
          continue;
        }
        if (nn < 0) break;
        un_hi--;
      }
      if (un_lo > un_hi) break;
      FSWAP(fmap[un_lo], fmap[un_hi]);
      un_lo++;
      un_hi--;
    }
    if (un_hi != un_lo - 1) return 0;
    if (gt_hi < lt_lo) continue;
    n = (lt_lo - lo) < (un_lo - lt_lo) ? (lt_lo - lo) : (un_lo - lt_lo);
    for (int32_t a = lo, b = un_lo - n, c = n; c > 0; a++, b++, c--) FSWAP(fmap[a], fmap[b]);
    m = (hi - gt_hi) < (gt_hi - un_hi) ? (hi - gt_hi) : (gt_hi - un_hi);
    for (int32_t a = un_lo, b = hi - m + 1, c = m; c > 0; a++, b++, c--) FSWAP(fmap[a], fmap[b]);
    n = lo + un_lo - lt_lo - 1;
    m = hi - (gt_hi - un_hi) + 1;
    if (n - lo > hi - m) {
      stack_lo[sp] = lo; stack_hi[sp] = n; sp++;
      stack_lo[sp] = m; stack_hi[sp] = hi; sp++;
    } else {
      stack_lo[sp] = m; stack_hi[sp] = hi; sp++;
      stack_lo[sp] = lo; stack_hi[sp] = n; sp++;
    }
  }
  return 1;
}
#define SET_BH(zz) bhtab[(zz) >> 5] |= (1u << ((zz) & 31))
#define CLEAR_BH(zz) bhtab[(zz) >> 5] &= ~(1u << ((zz) & 31))
#define ISSET_BH(zz) (bhtab[(zz) >> 5] & (1u << ((zz) & 31)))
#define WORD_BH(zz) bhtab[(zz) >> 5]
#define UNALIGNED_BH(zz) ((zz) & 0x01f)
static int fallback_sort(uint32_t *fmap, uint32_t *eclass, uint32_t *bhtab, int32_t nblock) {
  int32_t ftab[257], ftab_copy[256];
  int32_t H, i, j, k, l, r, cc, cc1, n_not_done, n_bhtab;
  uint8_t *eclass8 = (uint8_t *)eclass;
  for (i = 0; i < 257; i++) ftab[i] = 0;
  for (i = 0; i < nblock; i++) ftab[eclass8[i]]++;
  for (i = 0; i < 256; i++) ftab_copy[i] = ftab[i];
  for (i = 1; i < 257; i++) ftab[i] += ftab[i - 1];
  for (i = 0; i < nblock; i++) {
    j = eclass8[i];
    k = ftab[j] - 1;
    ftab[j] = k;
    fmap[k] = (uint32_t)i;
  }
  n_bhtab = 2 + (nblock / 32);
  for (i = 0; i < n_bhtab; i++) bhtab[i] = 0;
  for (i = 0; i < 256; i++) SET_BH(ftab[i]);
  for (i = 0; i < 32; i++) {
    SET_BH(nblock + 2 * i);
    CLEAR_BH(nblock + 2 * i + 1);
  }
  H = 1;
  for (;;) {
    j = 0;
    for (i = 0; i < nblock; i++) {
      if (ISSET_BH(i)) j = i;
      k = (int32_t)fmap[i] - H;
      if (k < 0) k += nblock;
      eclass[k] = (uint32_t)j;
    }
    n_not_done = 0;
    r = -1;
    for (;;) {
      k = r + 1;
      while (ISSET_BH(k) && UNALIGNED_BH(k)) k++;
      if (ISSET_BH(k)) {
        while (WORD_BH(k) == 0xffffffffu) k += 32;
        while (ISSET_BH(k)) k++;
      }
      l = k - 1;
      if (l >= nblock) break;
      while (!ISSET_BH(k) && UNALIGNED_BH(k)) k++;
      if (!ISSET_BH(k)) {
        while (WORD_BH(k) == 0x00000000u) k += 32;
        while (!ISSET_BH(k)) k++;
      }
      r = k - 1;
      if (r >= nblock) break;
      if (r > l) {
        n_not_done += (r - l + 1);
        if (!fallback_qsort3(fmap, eclass, l, r)) return 0;
        cc = -1;
        for (i = l; i <= r; i++) {
          cc1 = (int32_t)eclass[fmap[i]];
          if (cc != cc1) {
            SET_BH(i);
            cc = cc1;
          }
        }
      }
    }
    H *= 2;
    if (H > nblock || n_not_done == 0) break;
  }
  j = 0;
  for (i = 0; i < nblock; i++) {
    while (ftab_copy[j] == 0) j++;
    ftab_copy[j]--;
    eclass8[fmap[i]] = (uint8_t)j;
  }
  return j < 256;
}

/* ---- main sort (:1247-2011) ---- */
static int main_gtu(uint32_t i1, uint32_t i2, const uint8_t *block, const uint16_t *quadrant, uint32_t nblock, int32_t *budget) {
  int32_t k;
  uint8_t c1, c2;
  uint16_t s1, s2;
  if (i1 == i2) return 0;
  for (int t = 0; t < 12; ++t) {
    c1 = block[i1];
    c2 = block[i2];
    if (c1 != c2) return c1 > c2;
    i1++;
    i2++;
  }
  k = (int32_t)nblock + 8;
  do {
    for (int t = 0; t < 8; ++t) {
      c1 = block[i1];
      c2 = block[i2];
      if (c1 != c2) return c1 > c2;
      s1 = quadrant[i1];
      s2 = quadrant[i2];
      if (s1 != s2) return s1 > s2;
      i1++;
      i2++;
    }
    if (i1 >= nblock) i1 -= nblock;
    if (i2 >= nblock) i2 -= nblock;
    k -= 8;
    (*budget)--;
  } while (k >= 0);
  return 0;
}
static const int32_t incs[14] = {1, 4, 13, 40, 121, 364, 1093, 3280, 9841, 29524, 88573, 265720, 797161, 2391484};
static void main_simple_sort(uint32_t *ptr, const uint8_t *block, const uint16_t *quadrant, int32_t nblock, int32_t lo,
                             int32_t hi, int32_t d, int32_t *budget) {
  int32_t i, j, h, big_n, hp;
  uint32_t v;
  big_n = hi - lo + 1;
  if (big_n < 2) return;
  hp = 0;
  while (incs[hp] < big_n) hp++;
  hp--;
  for (; hp >= 0; hp--) {
    h = incs[hp];
    i = lo + h;
    for (;;) {
      for (int rep = 0; rep < 3; ++rep) { /* the reference unrolls this copy three times */
        if (i > hi) goto next_h;
        v = ptr[i];
        j = i;
        while (main_gtu(ptr[j - h] + d, v + d, block, quadrant, (uint32_t)nblock, budget)) {
          ptr[j] = ptr[j - h];
          j = j - h;
          if (j <= (lo + h - 1)) break;
        }
        ptr[j] = v;
        i++;
      }
      if (*budget < 0) return;
    }
  next_h:;
  }
}
static uint8_t mmed3(uint8_t a, uint8_t b, uint8_t c) {
  uint8_t t;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) {
    b = c;
    if (a > b) b = a;
  }
  return b;
}
static int main_qsort3(uint32_t *ptr, const uint8_t *block, const uint16_t *quadrant, int32_t nblock, int32_t lo_st,
                       int32_t hi_st, int32_t d_st, int32_t *budget) {
  int32_t un_lo, un_hi, lt_lo, gt_hi, n, m, med, sp, lo, hi, d;
  int32_t stack_lo[100], stack_hi[100], stack_d[100];
  int32_t next_lo[3], next_hi[3], next_d[3];
  sp = 0;
  stack_lo[sp] = lo_st; stack_hi[sp] = hi_st; stack_d[sp] = d_st; sp++;
  while (sp > 0) {
    if (sp >= 100 - 2) return 0;
    sp--;
    lo = stack_lo[sp]; hi = stack_hi[sp]; d = stack_d[sp];
    if (hi - lo < 20 || d > (BZ_N_RADIX + BZ_N_QSORT)) {
      main_simple_sort(ptr, block, quadrant, nblock, lo, hi, d, budget);
      if (*budget < 0) return 1;
      continue;
    }
    med = mmed3(block[ptr[lo] + d], block[ptr[hi] + d], block[ptr[(lo + hi) >> 1] + d]);
    un_lo = lt_lo = lo;
    un_hi = gt_hi = hi;
    for (;;) {
      for (;;) {
        if (un_lo > un_hi) break;
        n = (int32_t)block[ptr[un_lo] + d] - med;
        if (n == 0) {
          FSWAP(ptr[un_lo], ptr[lt_lo]);
          lt_lo++;
          un_lo++;
          continue;
        }
        if (n > 0) break;
        un_lo++;
      }
      for (;;) {
        if (un_lo > un_hi) break;
        n = (int32_t)block[ptr[un_hi] + d] - med;
        if (n == 0) {
          FSWAP(ptr[un_hi], ptr[gt_hi]);
          gt_hi--;
          un_hi--;
          continue;
        }
        if (n < 0) break;
        un_hi--;
      }
      if (un_lo > un_hi) break;
      FSWAP(ptr[un_lo], ptr[un_hi]);
      un_lo++;
      un_hi--;
    }
    if (un_hi != un_lo - 1) return 0;
    if (gt_hi < lt_lo) {
      stack_lo[sp] = lo; stack_hi[sp] = hi; stack_d[sp] = d + 1; sp++;
      continue;
    }
    n = (lt_lo - lo) < (un_lo - lt_lo) ? (lt_lo - lo) : (un_lo - lt_lo);
    for (int32_t a = lo, b = un_lo - n, c = n; c > 0; a++, b++, c--) FSWAP(ptr[a], ptr[b]);
    m = (hi - gt_hi) < (gt_hi - un_hi) ? (hi - gt_hi) : (gt_hi - un_hi);
    for (int32_t a = un_lo, b = hi - m + 1, c = m; c > 0; a++, b++, c--) FSWAP(ptr[a], ptr[b]);
    n = lo + un_lo - lt_lo - 1;
    m = hi - (gt_hi - un_hi) + 1;
    next_lo[0] = lo; next_hi[0] = n; next_d[0] = d;
    next_lo[1] = m; next_hi[1] = hi; next_d[1] = d;
    next_lo[2] = n + 1; next_hi[2] = m - 1; next_d[2] = d + 1;
#define NSIZE(a) (next_hi[a] - next_lo[a])
#define NSWAP(a, b) { int32_t t_; t_ = next_lo[a]; next_lo[a] = next_lo[b]; next_lo[b] = t_; t_ = next_hi[a]; next_hi[a] = next_hi[b]; next_hi[b] = t_; t_ = next_d[a]; next_d[a] = next_d[b]; next_d[b] = t_; }
    if (NSIZE(0) < NSIZE(1)) NSWAP(0, 1);
    if (NSIZE(1) < NSIZE(2)) NSWAP(1, 2);
    if (NSIZE(0) < NSIZE(1)) NSWAP(0, 1);
    if (NSIZE(0) < NSIZE(1)) return 0;
    if (NSIZE(1) < NSIZE(2)) return 0;
    for (int a = 0; a < 3; ++a) { stack_lo[sp] = next_lo[a]; stack_hi[sp] = next_hi[a]; stack_d[sp] = next_d[a]; sp++; }
  }
  return 1;
}
#define SETMASK (1u << 21)
#define CLEARMASK (~SETMASK)
#define BIGFREQ(b) (ftab[((b) + 1) << 8] - ftab[(b) << 8])
static int main_sort(uint32_t *ptr, uint8_t *block, uint16_t *quadrant, uint32_t *ftab, int32_t nblock, int32_t *budget) {
  int32_t i, j, k, ss, sb, running_order[256], copy_start[256], copy_end[256];
  uint8_t big_done[256], c1;
  uint16_t s;
  for (i = 65536; i >= 0; i--) ftab[i] = 0;
  j = block[0] << 8;
  for (i = nblock - 1; i >= 0; i--) {
    quadrant[i] = 0;
    j = (j >> 8) | ((uint16_t)block[i] << 8);
    ftab[j]++;
  }
  for (i = 0; i < BZ_N_OVERSHOOT; i++) {
    block[nblock + i] = block[i];
    quadrant[nblock + i] = 0;
  }
  for (i = 1; i <= 65536; i++) ftab[i] += ftab[i - 1];
  s = (uint16_t)(block[0] << 8);
  for (i = nblock - 1; i >= 0; i--) {
    s = (uint16_t)((s >> 8) | (block[i] << 8));
    j = (int32_t)ftab[s] - 1;
    ftab[s] = (uint32_t)j;
    ptr[j] = (uint32_t)i;
  }
  for (i = 0; i <= 255; i++) {
    big_done[i] = 0;
    running_order[i] = i;
  }
  {
    int32_t vv, h = 1;
    do h = 3 * h + 1;
    while (h <= 256);
    do {
      h = h / 3;
      for (i = h; i <= 255; i++) {
        vv = running_order[i];
        j = i;
        while (BIGFREQ(running_order[j - h]) > BIGFREQ(vv)) {
          running_order[j] = running_order[j - h];
          j = j - h;
          if (j <= (h - 1)) break;
        }
        running_order[j] = vv;
      }
    } while (h != 1);
  }
  for (i = 0; i <= 255; i++) {
    ss = running_order[i];
    for (j = 0; j <= 255; j++) {
      if (j != ss) {
        sb = (ss << 8) + j;
        if (!(ftab[sb] & SETMASK)) {
          int32_t lo = (int32_t)(ftab[sb] & CLEARMASK), hi = (int32_t)(ftab[sb + 1] & CLEARMASK) - 1;
          if (hi > lo) {
            if (!main_qsort3(ptr, block, quadrant, nblock, lo, hi, BZ_N_RADIX, budget)) return 0;
            if (*budget < 0) return 1;
          }
        }
        ftab[sb] |= SETMASK;
      }
    }
    if (big_done[ss]) return 0;
    for (j = 0; j <= 255; j++) {
      copy_start[j] = (int32_t)(ftab[(j << 8) + ss] & CLEARMASK);
      copy_end[j] = (int32_t)(ftab[(j << 8) + ss + 1] & CLEARMASK) - 1;
    }
    for (j = (int32_t)(ftab[ss << 8] & CLEARMASK); j < copy_start[ss]; j++) {
      k = (int32_t)ptr[j] - 1;
      if (k < 0) k += nblock;
      c1 = block[k];
      if (!big_done[c1]) ptr[copy_start[c1]++] = (uint32_t)k;
    }
    for (j = (int32_t)(ftab[(ss + 1) << 8] & CLEARMASK) - 1; j > copy_end[ss]; j--) {
      k = (int32_t)ptr[j] - 1;
      if (k < 0) k += nblock;
      c1 = block[k];
      if (!big_done[c1]) ptr[copy_end[c1]--] = (uint32_t)k;
    }
    if (!((copy_start[ss] - 1 == copy_end[ss]) || (copy_start[ss] == 0 && copy_end[ss] == nblock - 1))) return 0;
    for (j = 0; j <= 255; j++) ftab[(j << 8) + ss] |= SETMASK;
    big_done[ss] = 1;
    if (i < 255) {
      int32_t bb_start = (int32_t)(ftab[ss << 8] & CLEARMASK);
      int32_t bb_size = (int32_t)(ftab[(ss + 1) << 8] & CLEARMASK) - bb_start;
      int32_t shifts = 0;
      if (bb_size > 0) {
        while ((bb_size >> shifts) > 65534) shifts++;
        for (j = bb_size - 1; j >= 0; j--) {
          int32_t a2update = (int32_t)ptr[bb_start + j];
          uint16_t q_val = (uint16_t)(j >> shifts);
          quadrant[a2update] = q_val;
          if (a2update < BZ_N_OVERSHOOT) quadrant[a2update + nblock] = q_val;
        }
      }
    }
  }
  return 1;
}

/* _blockSort :880-928 */
static int block_sort(bzenc *s) {
  int32_t nblock = s->nblock;
  if (nblock < 10000) {
    if (!fallback_sort(s->arr1, s->arr2, s->ftab, nblock)) return 0;
  } else {
    int32_t i = nblock + BZ_N_OVERSHOOT;
    if (i & 1) i++;
    uint16_t *quadrant = (uint16_t *)(&(s->block[i]));
    int32_t wfact = s->work_factor;
    if (wfact < 1) wfact = 1;
    if (wfact > 100) wfact = 100;
    s->budget = nblock * ((wfact - 1) / 3);
    if (!main_sort(s->arr1, s->block, quadrant, s->ftab, nblock, &s->budget)) return 0;
    if (s->budget < 0) {
      if (!fallback_sort(s->arr1, s->arr2, s->ftab, nblock)) return 0;
    }
  }
  s->orig_ptr = -1;
  for (int32_t i = 0; i < nblock; i++)
    if (s->arr1[i] == 0) {
      s->orig_ptr = i;
      break;
    }
  return s->orig_ptr != -1;
}

/* _generateMTFValues :139-265 */
static int generate_mtf_values(bzenc *s) {
  uint8_t yy[256];
  int32_t i, j, z_pend = 0, wr = 0, eob;
  s->n_in_use = 0;
  for (i = 0; i < 256; i++)
    if (s->in_use[i]) s->unseq_to_seq[i] = (uint8_t)s->n_in_use++;
  eob = s->n_in_use + 1;
  memset(s->mtf_freq, 0, sizeof s->mtf_freq);
  for (i = 0; i < s->n_in_use; i++) yy[i] = (uint8_t)i;
#define FLUSH_ZPEND()                                                    \
  if (z_pend > 0) {                                                      \
    z_pend--;                                                            \
    for (;;) {                                                           \
      if (z_pend & 1) { s->mtfv[wr++] = BZ_RUNB; s->mtf_freq[BZ_RUNB]++; } \
      else { s->mtfv[wr++] = BZ_RUNA; s->mtf_freq[BZ_RUNA]++; }            \
      if (z_pend < 2) break;                                             \
      z_pend = (z_pend - 2) / 2;                                         \
    }                                                                    \
    z_pend = 0;                                                          \
  }
  for (i = 0; i < s->nblock; i++) {
    if (wr > i) return 0;
    j = (int32_t)s->arr1[i] - 1;
    if (j < 0) j += s->nblock;
    uint8_t ll_i = s->unseq_to_seq[s->block[j]];
    if (ll_i >= s->n_in_use) return 0;
    if (yy[0] == ll_i) {
      z_pend++;
    } else {
      FLUSH_ZPEND();
      {
        uint8_t rtmp = yy[1], rtmp2;
        int32_t ryy_j = 1;
        yy[1] = yy[0];
        while (ll_i != rtmp) {
          ryy_j++;
          rtmp2 = rtmp;
          rtmp = yy[ryy_j];
          yy[ryy_j] = rtmp2;
        }
        yy[0] = rtmp;
        j = ryy_j;
        s->mtfv[wr++] = (uint16_t)(j + 1);
        s->mtf_freq[j + 1]++;
      }
    }
  }
  FLUSH_ZPEND();
  s->mtfv[wr++] = (uint16_t)eob;
  s->mtf_freq[eob]++;
  s->n_mtf = wr;
  return 1;
}

/* _hbMakeCodeLengths :747-864 */
static int hb_make_code_lengths(uint8_t *len, const int32_t *freq, int32_t alpha, int32_t max_len) {
  int32_t heap[BZ_MAX_ALPHA + 2], weight[BZ_MAX_ALPHA * 2], parent[BZ_MAX_ALPHA * 2];
  int32_t n_nodes, n_heap, n1, n2, i, j, k;
  int too_long;
  for (i = 0; i < alpha; i++) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
#define WEIGHTOF(z) ((z) & 0xffffff00)
#define DEPTHOF(z) ((z) & 0x000000ff)
#define MYMAX(a, b) ((a) > (b) ? (a) : (b))
#define ADDWEIGHTS(a, b) (int32_t)((WEIGHTOF(a) + WEIGHTOF(b)) | (1 + MYMAX(DEPTHOF(a), DEPTHOF(b))))
#define UPHEAP(z) { int32_t zz = z, tmp = heap[zz]; while (weight[tmp] < weight[heap[zz >> 1]]) { heap[zz] = heap[zz >> 1]; zz >>= 1; } heap[zz] = tmp; }
#define DOWNHEAP(z) { int32_t zz = z, yy, tmp = heap[zz]; for (;;) { yy = zz << 1; if (yy > n_heap) break; if (yy < n_heap && weight[heap[yy + 1]] < weight[heap[yy]]) yy++; if (weight[tmp] < weight[heap[yy]]) break; heap[zz] = heap[yy]; zz = yy; } heap[zz] = tmp; }
  for (;;) {
    n_nodes = alpha;
    n_heap = 0;
    heap[0] = 0;
    weight[0] = 0;
    parent[0] = -2;
    for (i = 1; i <= alpha; i++) {
      parent[i] = -1;
      n_heap++;
      heap[n_heap] = i;
      UPHEAP(n_heap);
    }
    if (n_heap >= BZ_MAX_ALPHA + 2) return 0;
    while (n_heap > 1) {
      n1 = heap[1]; heap[1] = heap[n_heap]; n_heap--; DOWNHEAP(1);
      n2 = heap[1]; heap[1] = heap[n_heap]; n_heap--; DOWNHEAP(1);
      n_nodes++;
      parent[n1] = parent[n2] = n_nodes;
      weight[n_nodes] = ADDWEIGHTS(weight[n1], weight[n2]);
      parent[n_nodes] = -1;
      n_heap++;
      heap[n_heap] = n_nodes;
      UPHEAP(n_heap);
    }
    if (n_nodes >= BZ_MAX_ALPHA * 2) return 0;
    too_long = 0;
    for (i = 1; i <= alpha; i++) {
      j = 0;
      k = i;
      while (parent[k] >= 0) {
        k = parent[k];
        j++;
      }
      len[i - 1] = (uint8_t)j;
      if (j > max_len) too_long = 1;
    }
    if (!too_long) break;
    for (i = 1; i <= alpha; i++) {
      j = weight[i] >> 8;
      j = 1 + (j / 2);
      weight[i] = j << 8;
    }
  }
  return 1;
}

/* _sendMTFValues :267-745 */
static int send_mtf_values(bzenc *s) {
  int32_t v, t, i, j, gs, ge, bt, bc, iter, n_selectors = 0, alpha, min_len, max_len, sel_ctr, n_groups;
  uint16_t cost[BZ_N_GROUPS];
  int32_t fave[BZ_N_GROUPS];
  uint16_t *mtfv = s->mtfv;
  alpha = s->n_in_use + 2;
  for (t = 0; t < BZ_N_GROUPS; t++)
    for (v = 0; v < alpha; v++) s->len[t][v] = 15;
  if (s->n_mtf <= 0) return 0;
  if (s->n_mtf < 200) n_groups = 2;
  else if (s->n_mtf < 600) n_groups = 3;
  else if (s->n_mtf < 1200) n_groups = 4;
  else if (s->n_mtf < 2400) n_groups = 5;
  else n_groups = 6;
  {
    int32_t n_part = n_groups, rem_f = s->n_mtf, t_freq, a_freq;
    gs = 0;
    while (n_part > 0) {
      t_freq = rem_f / n_part;
      ge = gs - 1;
      a_freq = 0;
      while (a_freq < t_freq && ge < alpha - 1) {
        ge++;
        a_freq += s->mtf_freq[ge];
      }
      if (ge > gs && n_part != n_groups && n_part != 1 && ((n_groups - n_part) % 2 == 1)) {
        a_freq -= s->mtf_freq[ge];
        ge--;
      }
      for (v = 0; v < alpha; v++) s->len[n_part - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
      n_part--;
      gs = ge + 1;
      rem_f -= a_freq;
    }
  }
  for (iter = 0; iter < BZ_N_ITERS; iter++) {
    for (t = 0; t < n_groups; t++) fave[t] = 0;
    for (t = 0; t < n_groups; t++)
      for (v = 0; v < alpha; v++) s->rfreq[t][v] = 0;
    if (n_groups == 6) {
      for (v = 0; v < alpha; v++) {
        s->len_pack[v][0] = ((uint32_t)s->len[1][v] << 16) | s->len[0][v];
        s->len_pack[v][1] = ((uint32_t)s->len[3][v] << 16) | s->len[2][v];
        s->len_pack[v][2] = ((uint32_t)s->len[5][v] << 16) | s->len[4][v];
      }
    }
    n_selectors = 0;
    gs = 0;
    for (;;) {
      if (gs >= s->n_mtf) break;
      ge = gs + BZ_G_SIZE - 1;
      if (ge >= s->n_mtf) ge = s->n_mtf - 1;
      for (t = 0; t < n_groups; t++) cost[t] = 0;
      if (n_groups == 6 && 50 == ge - gs + 1) {
        uint32_t cost01 = 0, cost23 = 0, cost45 = 0;
        for (int nn = 0; nn < 50; ++nn) {
          uint16_t icv = mtfv[gs + nn];
          cost01 += s->len_pack[icv][0];
          cost23 += s->len_pack[icv][1];
          cost45 += s->len_pack[icv][2];
        }
        cost[0] = cost01 & 0xffff; cost[1] = cost01 >> 16;
        cost[2] = cost23 & 0xffff; cost[3] = cost23 >> 16;
        cost[4] = cost45 & 0xffff; cost[5] = cost45 >> 16;
      } else {
        for (i = gs; i <= ge; i++) {
          uint16_t icv = mtfv[i];
          for (t = 0; t < n_groups; t++) cost[t] = (uint16_t)(cost[t] + s->len[t][icv]);
        }
      }
      bc = 999999999;
      bt = -1;
      for (t = 0; t < n_groups; t++)
        if (cost[t] < bc) {
          bc = cost[t];
          bt = t;
        }
      fave[bt]++;
      s->selector[n_selectors++] = (uint8_t)bt;
      for (i = gs; i <= ge; i++) s->rfreq[bt][mtfv[i]]++;
      gs = ge + 1;
    }
    for (t = 0; t < n_groups; t++)
      if (!hb_make_code_lengths(s->len[t], s->rfreq[t], alpha, 17)) return 0;
  }
  if (n_groups >= 8) return 0;
  if (!(n_selectors < 32768 && n_selectors <= BZ_MAX_SELECTORS)) return 0;
  {
    uint8_t pos[BZ_N_GROUPS], ll_i, tmp2, tmp;
    for (i = 0; i < n_groups; i++) pos[i] = (uint8_t)i;
    for (i = 0; i < n_selectors; i++) {
      ll_i = s->selector[i];
      j = 0;
      tmp = pos[j];
      while (ll_i != tmp) {
        j++;
        tmp2 = tmp;
        tmp = pos[j];
        pos[j] = tmp2;
      }
      pos[0] = tmp;
      s->selector_mtf[i] = (uint8_t)j;
    }
  }
  for (t = 0; t < n_groups; t++) {
    min_len = 32;
    max_len = 0;
    for (i = 0; i < alpha; i++) {
      if (s->len[t][i] > max_len) max_len = s->len[t][i];
      if (s->len[t][i] < min_len) min_len = s->len[t][i];
    }
    if (max_len > 17) return 0;
    if (min_len < 1) return 0;
    /* _hbAssignCodes :866-878 */
    int32_t vec = 0;
    for (int32_t n = min_len; n <= max_len; n++) {
      for (i = 0; i < alpha; i++)
        if (s->len[t][i] == n) s->code[t][i] = vec++;
      vec <<= 1;
    }
  }
  {
    uint8_t in_use16[16];
    for (i = 0; i < 16; i++) {
      in_use16[i] = 0;
      for (j = 0; j < 16; j++)
        if (s->in_use[i * 16 + j]) in_use16[i] = 1;
    }
    for (i = 0; i < 16; i++) bw_bits(s, 1, in_use16[i] ? 1 : 0);
    for (i = 0; i < 16; i++)
      if (in_use16[i])
        for (j = 0; j < 16; j++) bw_bits(s, 1, s->in_use[i * 16 + j] ? 1 : 0);
  }
  bw_bits(s, 3, (uint32_t)n_groups);
  bw_bits(s, 15, (uint32_t)n_selectors);
  for (i = 0; i < n_selectors; i++) {
    for (j = 0; j < s->selector_mtf[i]; j++) bw_bits(s, 1, 1);
    bw_bits(s, 1, 0);
  }
  for (t = 0; t < n_groups; t++) {
    int32_t curr = s->len[t][0];
    bw_bits(s, 5, (uint32_t)curr);
    for (i = 0; i < alpha; i++) {
      while (curr < s->len[t][i]) {
        bw_bits(s, 2, 2);
        curr++;
      }
      while (curr > s->len[t][i]) {
        bw_bits(s, 2, 3);
        curr--;
      }
      bw_bits(s, 1, 0);
    }
  }
  sel_ctr = 0;
  gs = 0;
  for (;;) {
    if (gs >= s->n_mtf) break;
    ge = gs + BZ_G_SIZE - 1;
    if (ge >= s->n_mtf) ge = s->n_mtf - 1;
    if (s->selector[sel_ctr] >= n_groups) return 0;
    for (i = gs; i <= ge; i++) bw_bits(s, s->len[s->selector[sel_ctr]][mtfv[i]], (uint32_t)s->code[s->selector[sel_ctr]][mtfv[i]]);
    gs = ge + 1;
    sel_ctr++;
  }
  return sel_ctr == n_selectors;
}

/* _compressBlock :112-137 */
static int compress_block(bzenc *s) {
  static const uint8_t magic[6] = {0x31, 0x41, 0x59, 0x26, 0x53, 0x59};
  if (s->nblock > 0) {
    if (!block_sort(s)) return 0;
    for (int i = 0; i < 6; ++i) bw_bits(s, 8, magic[i]);
    bw_bits(s, 32, s->block_crc);
    bw_bits(s, 1, 0);
    bw_bits(s, 24, (uint32_t)s->orig_ptr);
    if (!generate_mtf_values(s)) return 0;
    if (!send_mtf_values(s)) return 0;
  }
  return 1;
}

int orc_bzip2_encode_bytes(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  crc_init();
  bzenc *s = (bzenc *)calloc(1, sizeof(bzenc));
  orc_oms o;
  orc_oms_init(&o, 0x8000);
  s->out = &o;
  s->bit_pos = 8;
  s->in = in;
  s->in_len = n;
  const int32_t N = 100000 * 9;
  s->nblock_max = N - 19;
  s->work_factor = 30;
  s->arr1 = (uint32_t *)calloc((size_t)N, 4);
  s->arr2 = (uint32_t *)calloc((size_t)N + BZ_N_OVERSHOOT, 4);
  s->ftab = (uint32_t *)calloc(65537, 4);
  s->block = (uint8_t *)s->arr2;
  s->mtfv = (uint16_t *)s->arr1;
  s->selector = (uint8_t *)calloc(BZ_MAX_SELECTORS, 1);
  s->selector_mtf = (uint8_t *)calloc(BZ_MAX_SELECTORS, 1);
  bw_bits(s, 8, 0x42);
  bw_bits(s, 8, 0x5a);
  bw_bits(s, 8, 0x68);
  bw_bits(s, 8, 0x30 + 9);
  uint32_t combined = 0;
  int ok = 1;
  while (s->in_pos < s->in_len) { /* encodeStream :61-69 + _writeBlock :83-110 */
    memset(s->in_use, 0, 256);
    s->nblock = 0;
    s->block_crc = 0xffffffffu;
    s->state_in_ch = 256;
    s->state_in_len = 0;
    while (s->nblock < s->nblock_max && s->in_pos < s->in_len) add_char_to_block(s, s->in[s->in_pos++]);
    if (s->state_in_ch < 256) add_pair_to_block(s);
    s->state_in_ch = 256;
    s->state_in_len = 0;
    s->block_crc = ~s->block_crc;
    if (!compress_block(s)) {
      ok = 0;
      break;
    }
    combined = (combined << 1) | (combined >> 31);
    combined ^= s->block_crc;
  }
  if (ok) {
    static const uint8_t eos[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};
    for (int i = 0; i < 6; ++i) bw_bits(s, 8, eos[i]);
    bw_bits(s, 32, combined);
    bw_flush(s);
  }
  free(s->arr1);
  free(s->arr2);
  free(s->ftab);
  free(s->selector);
  free(s->selector_mtf);
  free(s);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return ok ? ORC_OK : ORC_FALSE;
}
