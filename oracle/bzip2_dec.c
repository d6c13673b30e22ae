/*
 * oracle/bzip2_dec.c -- CPU ORACLE (test infrastructure only; see orc.h).
 *
 * Restates, with the same control flow and quirks:
 *   lib/src/codecs/bzip2_decoder.dart          (BZip2Decoder)
 *   lib/src/codecs/bzip2/bz2_bit_reader.dart   (Bz2BitReader, MSB-first)
 *   lib/src/codecs/bzip2/bzip2.dart            (CRC: poly 0x04c11db7, MSB-first)
 */
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

#include "bz2_rnums.h"
#include "orc.h"

#define BZ_N_GROUPS 6
#define BZ_G_SIZE 50
#define BZ_MAX_ALPHA 258
#define BZ_MAX_CODE_LEN 23
#define BZ_MAX_SELECTORS (2 + (900000 / BZ_G_SIZE))
#define MTFA_SIZE 4096
#define MTFL_SIZE 16

static __thread jmp_buf *bz_jmp;
static void bz_throw(void) { longjmp(*bz_jmp, 1); }

static uint32_t bzcrc_tab[256];
static void bzcrc_init(void) {
  if (bzcrc_tab[1]) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i << 24;
    for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : c << 1;
    bzcrc_tab[i] = c;
  }
}
/* BZip2.updateCrc bzip2.dart:11-14 */
static inline uint32_t bz_update_crc(int value, uint32_t crc) {
  return (crc << 8) ^ bzcrc_tab[((crc >> 24) & 0xff) ^ (value & 0xff)];
}
uint32_t orc_bz2_crc(const uint8_t *p, size_t n) {
  bzcrc_init();
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) c = bz_update_crc(p[i], c);
  return c ^ 0xffffffffu;
}

/* Bz2BitReader bz2_bit_reader.dart:4-50 */
typedef struct {
  const uint8_t *buf;
  int64_t len, pos;
  int bit_buffer, bit_pos;
} bitrd;
static int in_read_byte(bitrd *b) {
  if (b->pos >= b->len) bz_throw(); /* InputMemoryStream.readByte past the end: RangeError */
  return b->buf[b->pos++];
}
static int64_t read_bits(bitrd *b, int n) {
  static const int mask[9] = {0, 1, 3, 7, 15, 31, 63, 127, 255};
  if (n == 0) return 0;
  if (b->bit_pos == 0) {
    b->bit_pos = 8;
    b->bit_buffer = in_read_byte(b);
  }
  int64_t value = 0;
  while (n > b->bit_pos) {
    value = (value << b->bit_pos) + (b->bit_buffer & mask[b->bit_pos]);
    n -= b->bit_pos;
    b->bit_pos = 8;
    b->bit_buffer = in_read_byte(b);
  }
  if (n > 0) {
    if (b->bit_pos == 0) {
      b->bit_pos = 8;
      b->bit_buffer = in_read_byte(b);
    }
    value = (value << n) + ((b->bit_buffer >> (b->bit_pos - n)) & mask[n]);
    b->bit_pos -= n;
  }
  return value;
}

typedef struct {
  bitrd br;
  orc_oms *out;
  int block_size_100k;
  uint32_t *tt;
  uint8_t in_use16[16], in_use[256], seq_to_unseq[256];
  uint8_t mtfa[MTFA_SIZE];
  int32_t mtfbase[256 / MTFL_SIZE];
  uint8_t *selector_mtf, *selector;
  int32_t limit[BZ_N_GROUPS][BZ_MAX_ALPHA], base[BZ_N_GROUPS][BZ_MAX_ALPHA], perm[BZ_N_GROUPS][BZ_MAX_ALPHA];
  int32_t min_lens[BZ_N_GROUPS];
  int32_t unzftab[256], cftab[257];
  uint8_t len[BZ_N_GROUPS][BZ_MAX_ALPHA];
  int num_selectors, group_pos, group_no, g_sel, g_minlen, num_in_use;
} bzdec;

/* _hbCreateDecodeTables :774-813 */
static void hb_create_decode_tables(int32_t *limit, int32_t *base, int32_t *perm, const uint8_t *length, int min_len,
                                    int max_len, int alpha) {
  int pp = 0;
  for (int i = min_len; i <= max_len; i++)
    for (int j = 0; j < alpha; j++)
      if (length[j] == i) perm[pp++] = j;
  for (int i = 0; i < BZ_MAX_CODE_LEN; i++) base[i] = 0;
  for (int i = 0; i < alpha; i++) base[length[i] + 1]++;
  for (int i = 1; i < BZ_MAX_CODE_LEN; i++) base[i] += base[i - 1];
  for (int i = 0; i < BZ_MAX_CODE_LEN; i++) limit[i] = 0;
  int32_t vec = 0;
  for (int i = min_len; i <= max_len; i++) {
    vec += (base[i + 1] - base[i]);
    limit[i] = vec - 1;
    vec <<= 1;
  }
  for (int i = min_len + 1; i <= max_len; i++) base[i] = ((limit[i - 1] + 1) << 1) - base[i];
}

/* _getMtfVal :732-772 */
static int get_mtf_val(bzdec *d) {
  if (d->group_pos == 0) {
    d->group_no++;
    if (d->group_no >= d->num_selectors) return -1;
    d->group_pos = BZ_G_SIZE;
    d->g_sel = d->selector[d->group_no];
    d->g_minlen = d->min_lens[d->g_sel];
  }
  d->group_pos--;
  int zn = d->g_minlen;
  int64_t zvec = read_bits(&d->br, zn);
  const int32_t *lim = d->limit[d->g_sel], *bas = d->base[d->g_sel], *prm = d->perm[d->g_sel];
  for (;;) {
    if (zn > 20) return -1;
    if (zvec <= lim[zn]) break;
    zn++;
    int64_t zj = read_bits(&d->br, 1);
    zvec = (zvec << 1) | zj;
  }
  if (zvec - bas[zn] < 0 || zvec - bas[zn] >= BZ_MAX_ALPHA) return -1;
  return prm[zvec - bas[zn]];
}

static void out_byte(bzdec *d, int ch, uint32_t *crc) {
  orc_oms_write_byte(d->out, ch);
  *crc = bz_update_crc(ch, *crc);
}

/* _readCompressed :113-730.  Returns 0 and *crc_out, or -1 (data error). */
static int read_compressed(bzdec *d, uint32_t *crc_out) {
  bitrd *br = &d->br;
  int block_randomized = (int)read_bits(br, 1);
  int64_t orig_ptr = read_bits(br, 8);
  orig_ptr = (orig_ptr << 8) | read_bits(br, 8);
  orig_ptr = (orig_ptr << 8) | read_bits(br, 8);

  for (int i = 0; i < 16; ++i) d->in_use16[i] = (uint8_t)read_bits(br, 1);
  memset(d->in_use, 0, 256);
  for (int i = 0, k = 0; i < 16; ++i, k += 16)
    if (d->in_use16[i])
      for (int j = 0; j < 16; ++j) d->in_use[k + j] = (uint8_t)read_bits(br, 1);
  /* _makeMaps :815-823 */
  d->num_in_use = 0;
  memset(d->seq_to_unseq, 0, 256);
  for (int i = 0; i < 256; ++i)
    if (d->in_use[i]) d->seq_to_unseq[d->num_in_use++] = (uint8_t)i;
  if (d->num_in_use == 0) return -1;
  int alpha = d->num_in_use + 2;

  int num_groups = (int)read_bits(br, 3);
  if (num_groups < 2 || num_groups > 6) return -1;
  d->num_selectors = (int)read_bits(br, 15);
  if (d->num_selectors < 1) return -1;
  memset(d->selector_mtf, 0, BZ_MAX_SELECTORS);
  memset(d->selector, 0, BZ_MAX_SELECTORS);
  for (int i = 0; i < d->num_selectors; ++i) {
    int j = 0;
    for (;;) {
      if (read_bits(br, 1) == 0) break;
      j++;
      if (j >= num_groups) return -1;
    }
    if (i >= BZ_MAX_SELECTORS) bz_throw(); /* _selectorMtf[i]: RangeError (:168) */
    d->selector_mtf[i] = (uint8_t)j;
  }
  {
    uint8_t pos[BZ_N_GROUPS];
    for (int i = 0; i < num_groups; ++i) pos[i] = (uint8_t)i;
    for (int i = 0; i < d->num_selectors; ++i) {
      int v = d->selector_mtf[i];
      uint8_t tmp = pos[v];
      while (v > 0) {
        pos[v] = pos[v - 1];
        v--;
      }
      pos[0] = tmp;
      d->selector[i] = tmp;
    }
  }
  for (int t = 0; t < num_groups; ++t) {
    memset(d->len[t], 0, BZ_MAX_ALPHA);
    int64_t c = read_bits(br, 5);
    for (int i = 0; i < alpha; ++i) {
      for (;;) {
        if (c < 1 || c > 20) return -1;
        if (read_bits(br, 1) == 0) break;
        if (read_bits(br, 1) == 0) c++;
        else c--;
      }
      d->len[t][i] = (uint8_t)c;
    }
  }
  for (int t = 0; t < num_groups; t++) {
    memset(d->limit[t], 0, sizeof d->limit[t]);
    memset(d->base[t], 0, sizeof d->base[t]);
    memset(d->perm[t], 0, sizeof d->perm[t]);
    int min_len = 32, max_len = 0;
    for (int i = 0; i < alpha; ++i) {
      if (d->len[t][i] > max_len) max_len = d->len[t][i];
      if (d->len[t][i] < min_len) min_len = d->len[t][i];
    }
    hb_create_decode_tables(d->limit[t], d->base[t], d->perm[t], d->len[t], min_len, max_len, alpha);
    d->min_lens[t] = min_len;
  }

  int eob = d->num_in_use + 1;
  int64_t nblock_max = 100000ll * d->block_size_100k;
  memset(d->unzftab, 0, sizeof d->unzftab);
  {
    int kk = MTFA_SIZE - 1;
    for (int ii = 256 / MTFL_SIZE - 1; ii >= 0; ii--) {
      for (int jj = MTFL_SIZE - 1; jj >= 0; jj--) {
        d->mtfa[kk] = (uint8_t)(ii * MTFL_SIZE + jj);
        kk--;
      }
      d->mtfbase[ii] = kk + 1;
    }
  }
  int64_t nblock = 0;
  d->group_pos = 0;
  d->group_no = -1;
  int next_sym = get_mtf_val(d);
  if (next_sym < 0) return -1;
  int uc = 0;
  uint32_t *tt = d->tt;
  for (;;) {
    if (next_sym == eob) break;
    if (next_sym == 0 || next_sym == 1) {
      int64_t es = -1, N = 1;
      do {
        if (N >= 2 * 1024 * 1024) return -1;
        if (next_sym == 0) es = es + N;
        else if (next_sym == 1) es = es + 2 * N;
        N = N * 2;
        next_sym = get_mtf_val(d);
      } while (next_sym == 0 || next_sym == 1);
      es++;
      uc = d->seq_to_unseq[d->mtfa[d->mtfbase[0]]];
      d->unzftab[uc] += (int32_t)es;
      while (es > 0) {
        if (nblock >= nblock_max) return -1;
        tt[nblock] = (uint32_t)uc;
        nblock++;
        es--;
      }
      continue;
    } else {
      if (nblock >= nblock_max) return -1;
      int nn = next_sym - 1; /* next_sym == -1 (error from _getMtfVal) -> nn = -2: Dart indexes _mtfa[pp - 2] */
      if (next_sym < 0) {
        /* nn = -2 < mtflSize: uc = _mtfa[pp + nn]; the two while loops do not run; _mtfa[pp] = uc.
         * pp = _mtfbase[0] >= 2 except right after a re-pack... restate literally: */
      }
      if (nn < MTFL_SIZE) {
        int pp = d->mtfbase[0];
        if (pp + nn < 0) bz_throw();
        uc = d->mtfa[pp + nn];
        while (nn > 3) {
          int z = pp + nn;
          d->mtfa[z] = d->mtfa[z - 1];
          d->mtfa[z - 1] = d->mtfa[z - 2];
          d->mtfa[z - 2] = d->mtfa[z - 3];
          d->mtfa[z - 3] = d->mtfa[z - 4];
          nn -= 4;
        }
        while (nn > 0) {
          d->mtfa[pp + nn] = d->mtfa[pp + nn - 1];
          nn--;
        }
        d->mtfa[pp] = (uint8_t)uc;
      } else {
        int lno = nn / MTFL_SIZE, off = nn % MTFL_SIZE;
        int pp = d->mtfbase[lno] + off;
        uc = d->mtfa[pp];
        while (pp > d->mtfbase[lno]) {
          d->mtfa[pp] = d->mtfa[pp - 1];
          pp--;
        }
        d->mtfbase[lno]++;
        while (lno > 0) {
          d->mtfbase[lno]--;
          d->mtfa[d->mtfbase[lno]] = d->mtfa[d->mtfbase[lno - 1] + MTFL_SIZE - 1];
          lno--;
        }
        d->mtfbase[0]--;
        d->mtfa[d->mtfbase[0]] = (uint8_t)uc;
        if (d->mtfbase[0] == 0) {
          int kk = MTFA_SIZE - 1;
          for (int ii = 256 / MTFL_SIZE - 1; ii >= 0; ii--) {
            for (int jj = MTFL_SIZE - 1; jj >= 0; jj--) {
              d->mtfa[kk] = d->mtfa[d->mtfbase[ii] + jj];
              kk--;
            }
            d->mtfbase[ii] = kk + 1;
          }
        }
      }
      d->unzftab[d->seq_to_unseq[uc]]++;
      tt[nblock] = d->seq_to_unseq[uc];
      nblock++;
      next_sym = get_mtf_val(d);
      continue;
    }
  }

  if (orig_ptr < 0 || orig_ptr >= nblock) return -1;
  for (int i = 0; i <= 255; i++)
    if (d->unzftab[i] < 0 || d->unzftab[i] > nblock) return -1;
  d->cftab[0] = 0;
  for (int i = 1; i <= 256; i++) d->cftab[i] = d->unzftab[i - 1];
  for (int i = 1; i <= 256; i++) d->cftab[i] += d->cftab[i - 1];
  for (int i = 0; i <= 256; i++)
    if (d->cftab[i] < 0 || d->cftab[i] > nblock) return -1;
  for (int i = 1; i <= 256; i++)
    if (d->cftab[i - 1] > d->cftab[i]) return -1;
  for (int64_t i = 0; i < nblock; i++) {
    uc = tt[i] & 0xff;
    tt[d->cftab[uc]] |= (uint32_t)(i << 8);
    d->cftab[uc]++;
  }

  uint32_t crc = 0xffffffffu;
  int64_t tlimit = 100000ll * d->block_size_100k;
  uint32_t t_pos = tt[orig_ptr] >> 8;
  int64_t num_block_used = 0;
  int k0;
  int r_n_to_go = 0, r_t_pos = 0;
#define RAND_NEXT()                    \
  if (r_n_to_go == 0) {                \
    r_n_to_go = bz2_rnums[r_t_pos];    \
    r_t_pos++;                         \
    if (r_t_pos == 512) r_t_pos = 0;   \
  }
#define TT_NEXT_UNCHECKED(k)           \
  do {                                 \
    if (t_pos >= (uint64_t)tlimit) bz_throw(); /* _tt[tPos] RangeError */ \
    t_pos = tt[t_pos];                 \
    (k) = t_pos & 0xff;                \
    t_pos >>= 8;                       \
  } while (0)
  if (block_randomized) {
    if (t_pos >= tlimit) return -1;
    t_pos = tt[t_pos];
    k0 = t_pos & 0xff;
    t_pos >>= 8;
    num_block_used++;
    RAND_NEXT();
    r_n_to_go--;
    k0 ^= (r_n_to_go == 1) ? 1 : 0;
  } else {
    if (t_pos >= tlimit) {
      *crc_out = crc;
      return 0;
    }
    t_pos = tt[t_pos];
    k0 = t_pos & 0xff;
    t_pos >>= 8;
    num_block_used++;
  }
  int64_t out_len = 0;
  int out_ch = 0;
  int64_t save_nblock_pp = nblock + 1;
  int64_t n_used = num_block_used;
  int c_k0 = k0;
  int k1;
  if (block_randomized) {
    /* :492-608 -- NOTE quirk Q6: rNToGo is decremented only on the first read of each turn */
    for (;;) {
      for (;;) {
        if (out_len == 0) break;
        out_byte(d, out_ch, &crc);
        out_len--;
      }
      if (n_used == save_nblock_pp) {
        *crc_out = crc;
        return 0;
      }
      if (n_used > save_nblock_pp) return -1;
      out_len = 1;
      out_ch = k0;
      TT_NEXT_UNCHECKED(k1);
      RAND_NEXT();
      r_n_to_go--;
      k1 ^= (r_n_to_go == 1) ? 1 : 0;
      n_used++;
      if (n_used == save_nblock_pp) continue;
      if (k1 != k0) {
        k0 = k1;
        continue;
      }
      out_len = 2;
      TT_NEXT_UNCHECKED(k1);
      RAND_NEXT();
      k1 ^= (r_n_to_go == 1) ? 1 : 0;
      n_used++;
      if (n_used == save_nblock_pp) continue;
      if (k1 != k0) {
        k0 = k1;
        continue;
      }
      out_len = 3;
      TT_NEXT_UNCHECKED(k1);
      RAND_NEXT();
      k1 ^= (r_n_to_go == 1) ? 1 : 0;
      n_used++;
      if (n_used == save_nblock_pp) continue;
      if (k1 != k0) {
        k0 = k1;
        continue;
      }
      TT_NEXT_UNCHECKED(k1);
      RAND_NEXT();
      k1 ^= (r_n_to_go == 1) ? 1 : 0;
      n_used++;
      out_len = k1 + 4;
      TT_NEXT_UNCHECKED(k0);
      RAND_NEXT();
      k0 ^= (r_n_to_go == 1) ? 1 : 0;
      n_used++;
    }
  } else {
    /* :610-727 */
    for (;;) {
      if (out_len > 0) {
        for (;;) {
          if (out_len == 1) break;
          out_byte(d, out_ch, &crc);
          out_len--;
        }
        out_byte(d, out_ch, &crc);
      }
      if (n_used > save_nblock_pp) return -1;
      if (n_used == save_nblock_pp) {
        *crc_out = crc;
        return 0;
      }
      out_ch = c_k0;
#define TT_NEXT_CHECKED(k)             \
  do {                                 \
    if (t_pos >= tlimit) return -1;    \
    t_pos = tt[t_pos];                 \
    (k) = t_pos & 0xff;                \
    t_pos >>= 8;                       \
  } while (0)
      TT_NEXT_CHECKED(k1);
      n_used++;
      if (k1 != c_k0) {
        c_k0 = k1;
        out_byte(d, out_ch, &crc);
        out_len = 0;
        continue;
      }
      if (n_used == save_nblock_pp) {
        out_byte(d, out_ch, &crc);
        out_len = 0;
        continue;
      }
      out_len = 2;
      TT_NEXT_CHECKED(k1);
      n_used++;
      if (n_used == save_nblock_pp) continue;
      if (k1 != c_k0) {
        c_k0 = k1;
        continue;
      }
      out_len = 3;
      TT_NEXT_CHECKED(k1);
      n_used++;
      if (n_used == save_nblock_pp) continue;
      if (k1 != c_k0) {
        c_k0 = k1;
        continue;
      }
      TT_NEXT_CHECKED(k1);
      n_used++;
      out_len = k1 + 4;
      TT_NEXT_CHECKED(c_k0);
      n_used++;
    }
  }
}

/* decodeStream :20-88 */
static int bz_decode_run(bzdec *d, int verify) {
  bitrd *br = &d->br;
  d->group_pos = 0;
  d->group_no = 0;
  if (read_bits(br, 8) != 0x42 || read_bits(br, 8) != 0x5a || read_bits(br, 8) != 0x68) return ORC_FALSE;
  d->block_size_100k = (int)read_bits(br, 8) - 0x30;
  if (d->block_size_100k < 0 || d->block_size_100k > 9) return ORC_FALSE;
  d->tt = (uint32_t *)calloc((size_t)d->block_size_100k * 100000 + 1, sizeof(uint32_t));
  uint32_t combined = 0;
  while (!(br->pos >= br->len)) {
    /* _readBlockType :90-111 */
    int eos = 1, compressed = 1, type;
    static const int cm[6] = {0x31, 0x41, 0x59, 0x26, 0x53, 0x59}, em[6] = {0x17, 0x72, 0x45, 0x38, 0x50, 0x90};
    type = 0;
    for (int i = 0; i < 6; ++i) {
      int b = (int)read_bits(br, 8);
      if (b != cm[i]) compressed = 0;
      if (b != em[i]) eos = 0;
      if (!eos && !compressed) {
        type = -1;
        break;
      }
    }
    if (type < 0) return ORC_FALSE;
    if (compressed) {
      uint32_t stored = 0;
      for (int i = 0; i < 4; ++i) stored = (stored << 8) | (uint32_t)read_bits(br, 8);
      uint32_t crc;
      if (read_compressed(d, &crc) < 0) return ORC_FALSE;
      crc ^= 0xffffffffu;
      if (verify && crc != stored) return ORC_FALSE;
      combined = (combined << 1) | (combined >> 31);
      combined ^= crc;
    } else {
      uint32_t stored = 0;
      for (int i = 0; i < 4; ++i) stored = (stored << 8) | (uint32_t)read_bits(br, 8);
      if (verify && stored != combined) return ORC_FALSE;
      return ORC_OK;
    }
  }
  return ORC_OK;
}

int orc_bzip2_decode_bytes(const uint8_t *in, size_t n, int verify, uint8_t **out, size_t *out_len) {
  bzcrc_init();
  bzdec *d = (bzdec *)calloc(1, sizeof(bzdec));
  orc_oms o;
  orc_oms_init(&o, 0x8000);
  d->br.buf = in;
  d->br.len = (int64_t)n;
  d->out = &o;
  d->selector_mtf = (uint8_t *)calloc(BZ_MAX_SELECTORS, 1);
  d->selector = (uint8_t *)calloc(BZ_MAX_SELECTORS, 1);
  jmp_buf jb, *saved = bz_jmp;
  bz_jmp = &jb;
  int st;
  if (setjmp(jb) == 0) st = bz_decode_run(d, verify);
  else st = ORC_THROW;
  bz_jmp = saved;
  free(d->tt);
  free(d->selector_mtf);
  free(d->selector);
  free(d);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}
