/*
 * oracle/deflate.c -- CPU ORACLE (test infrastructure only; see orc.h).
 *
 * Restates lib/src/codecs/zlib/deflate.dart (class Deflate, _HuffmanTree, _StaticTree): the zlib-1.1.x
 * lineage encoder (via JZlib / zlib.NET) INCLUDING the early-flush heuristic in _trTally (:549-562) that
 * stock zlib compiles out, the old deflate_stored (:691-737), memLevel 8 and the 16-bit bit buffer.
 * Also the framing of _zlib_encoder_web.dart:27-73 and _gzip_encoder_web.dart:27-100.
 *
 * The constant tables (_distCode :1711, lengthCode :2226, baseLength :2485, baseDist :2517, staticLTree
 * :2801, staticDTree :3380) are regenerated here from their defining rules (RFC 1951 section 3.2.5/3.2.6)
 * rather than copied; tests/test_oracle_deflate.py checks the encoder against system zlib on inputs where
 * the heuristic cannot fire.
 *
 * PARITY UNPINNED by the reference's own tests: they only round-trip Deflate output.
 */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

enum {
  MAX_BITS = 15, BL_CODES = 19, D_CODES = 30, LITERALS = 256, LENGTH_CODES = 29,
  L_CODES = LITERALS + 1 + LENGTH_CODES, HEAP_SIZE = 2 * L_CODES + 1, END_BLOCK = 256,
  REP_3_6 = 16, REPZ_3_10 = 17, REPZ_11_138 = 18, MIN_MATCH = 3, MAX_MATCH = 258,
  MIN_LOOKAHEAD = MAX_MATCH + MIN_MATCH + 1, BUF_SIZE = 16, MAX_BL_BITS = 7,
  Z_BINARY = 0, Z_ASCII = 1, Z_UNKNOWN = 2,
  FN_STORED = 0, FN_FAST = 1, FN_SLOW = 2,
  NEED_MORE = 0, BLOCK_DONE = 1, FINISH_STARTED = 2, FINISH_DONE = 3,
  STORED_BLOCK = 0, STATIC_TREES = 1, DYN_TREES = 2
};

/* ---- constant tables, generated from their rules -------------------------------------------- */
static const uint8_t extra_lbits[LENGTH_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                                  2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint8_t extra_dbits[D_CODES] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6,
                                             6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t extra_blbits[BL_CODES] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
static const uint8_t bl_order[BL_CODES] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
static uint16_t static_ltree[(L_CODES + 2) * 2], static_dtree[D_CODES * 2];
static uint8_t dist_code_tab[512], length_code_tab[256];
static int base_length[LENGTH_CODES], base_dist[D_CODES];
static int tables_ready;

static unsigned bi_reverse(unsigned code, int len) { /* _reverseBits :2773-2781 */
  unsigned res = 0;
  do {
    res |= code & 1;
    code >>= 1;
    res <<= 1;
  } while (--len > 0);
  return res >> 1;
}
static void gen_codes(uint16_t *tree, int max_code, const uint16_t *bl_count) { /* _genCodes :2746-2767 */
  uint16_t next_code[MAX_BITS + 1];
  unsigned code = 0;
  for (int bits = 1; bits <= MAX_BITS; bits++) next_code[bits] = (uint16_t)(code = (code + bl_count[bits - 1]) << 1);
  for (int n = 0; n <= max_code; n++) {
    int len = tree[n * 2 + 1];
    if (len == 0) continue;
    tree[n * 2] = (uint16_t)bi_reverse(next_code[len]++, len);
  }
}
static void tables_init(void) {
  if (tables_ready) return;
  int length = 0, code, n, dist = 0;
  for (code = 0; code < LENGTH_CODES - 1; code++) {
    base_length[code] = length;
    for (n = 0; n < (1 << extra_lbits[code]); n++) length_code_tab[length++] = (uint8_t)code;
  }
  length_code_tab[length - 1] = (uint8_t)code; /* 258 -> code 28 */
  base_length[28] = 0;                          /* baseLength[28] = 0 in the reference table */
  for (code = 0; code < 16; code++) {
    base_dist[code] = dist;
    for (n = 0; n < (1 << extra_dbits[code]); n++) dist_code_tab[dist++] = (uint8_t)code;
  }
  dist >>= 7;
  for (; code < D_CODES; code++) {
    base_dist[code] = dist << 7;
    for (n = 0; n < (1 << (extra_dbits[code] - 7)); n++) dist_code_tab[256 + dist++] = (uint8_t)code;
  }
  uint16_t bl_count[MAX_BITS + 1];
  memset(bl_count, 0, sizeof bl_count);
  n = 0;
  while (n <= 143) static_ltree[n++ * 2 + 1] = 8, bl_count[8]++;
  while (n <= 255) static_ltree[n++ * 2 + 1] = 9, bl_count[9]++;
  while (n <= 279) static_ltree[n++ * 2 + 1] = 7, bl_count[7]++;
  while (n <= 287) static_ltree[n++ * 2 + 1] = 8, bl_count[8]++;
  gen_codes(static_ltree, L_CODES + 1, bl_count);
  for (n = 0; n < D_CODES; n++) {
    static_dtree[n * 2 + 1] = 5;
    static_dtree[n * 2] = (uint16_t)bi_reverse((unsigned)n, 5);
  }
  tables_ready = 1;
}
static int d_code(int dist) { return dist < 256 ? dist_code_tab[dist] : dist_code_tab[256 + (dist >> 7)]; } /* :2786 */

typedef struct {
  const uint16_t *static_tree; /* or NULL */
  const uint8_t *extra_bits;
  int extra_base, elems, max_length;
} static_desc;
typedef struct {
  uint16_t *dyn_tree;
  int max_code;
  const static_desc *stat;
} tree_desc;
static const static_desc static_l_desc = {static_ltree, extra_lbits, LITERALS + 1, L_CODES, MAX_BITS};
static const static_desc static_d_desc = {static_dtree, extra_dbits, 0, D_CODES, MAX_BITS};
static const static_desc static_bl_desc = {NULL, extra_blbits, 0, BL_CODES, MAX_BL_BITS};

typedef struct {
  int good, lazy, nice, chain, func;
} config;
static const config config_table[10] = { /* _getConfig :1250-1275 */
    {0, 0, 0, 0, FN_STORED},    {4, 4, 8, 4, FN_FAST},        {4, 5, 16, 8, FN_FAST},      {4, 6, 32, 32, FN_FAST},
    {4, 4, 16, 16, FN_SLOW},    {8, 16, 32, 32, FN_SLOW},     {8, 16, 128, 128, FN_SLOW},  {8, 32, 128, 256, FN_SLOW},
    {32, 128, 258, 1024, FN_SLOW}, {32, 258, 258, 4096, FN_SLOW}};

typedef struct {
  const uint8_t *in;
  int64_t in_len, in_pos;
  orc_oms *out;
  uint32_t crc;
  int64_t total;
  int level, strategy;
  config cfg;
  uint16_t dyn_ltree[HEAP_SIZE * 2], dyn_dtree[(2 * D_CODES + 1) * 2], bl_tree[(2 * BL_CODES + 1) * 2];
  int w_bits, w_size, w_mask, hash_bits, hash_size, hash_mask, hash_shift;
  uint8_t *window;
  uint16_t *prev, *head;
  int lit_bufsize;
  uint8_t *pending_buf;
  int pending_buf_size, d_buf, l_buf, pending;
  tree_desc l_desc, d_desc, bl_desc;
  int64_t bi_buf;
  int bi_valid, last_eob_len;
  int64_t opt_len, static_len;
  int last_lit, matches;
  uint32_t heap[2 * L_CODES + 1];
  int heap_len, heap_max;
  uint8_t depth[2 * L_CODES + 1];
  uint16_t bl_count[MAX_BITS + 1];
  int strstart, block_start, lookahead, match_length, prev_length, match_available, ins_h, match_start, prev_match;
  int window_size; /* _actualWindowSize */
  int data_type;
} deflate_t;

static int in_eos(const deflate_t *s) { return s->in_pos >= s->in_len; }

/* ---- pending buffer / bit output (:464-499, :640-675) ---------------------------------------- */
static void flush_pending(deflate_t *s) { /* _flushPending :1239-1248 */
  orc_oms_write_bytes(s->out, s->pending_buf, s->pending);
  s->pending = 0;
}
static void put_byte(deflate_t *s, int64_t c) { s->pending_buf[s->pending++] = (uint8_t)c; }
static void put_short(deflate_t *s, int64_t w) {
  put_byte(s, w);
  put_byte(s, w >> 8); /* _rshift(w, 8): w >= 0 here */
}
static void send_bits(deflate_t *s, int64_t value, int length) { /* _sendBits :487-499 */
  if (s->bi_valid > BUF_SIZE - length) {
    s->bi_buf = s->bi_buf | ((value << s->bi_valid) & 0xffff);
    put_short(s, s->bi_buf);
    s->bi_buf = value >> (BUF_SIZE - s->bi_valid);
    s->bi_valid += length - BUF_SIZE;
  } else {
    s->bi_buf = s->bi_buf | ((value << s->bi_valid) & 0xffff);
    s->bi_valid += length;
  }
}
static void send_code(deflate_t *s, int c, const uint16_t *tree) { send_bits(s, tree[c * 2], tree[c * 2 + 1]); }
static void bi_windup(deflate_t *s) { /* _biWindup :653-661 */
  if (s->bi_valid > 8) put_short(s, s->bi_buf);
  else if (s->bi_valid > 0) put_byte(s, s->bi_buf);
  s->bi_buf = 0;
  s->bi_valid = 0;
}
static void copy_block(deflate_t *s, int buf, int len, int header) { /* _copyBlock :665-675 */
  bi_windup(s);
  s->last_eob_len = 8;
  if (header) {
    put_short(s, len);
    put_short(s, (~len + 0x10000) & 0xffff);
  }
  if (len) {
    memcpy(s->pending_buf + s->pending, s->window + buf, (size_t)len);
    s->pending += len;
  }
}

/* ---- trees ------------------------------------------------------------------------------------ */
static void init_block(deflate_t *s) { /* _initBlock :278-286 */
  for (int i = 0; i < L_CODES; i++) s->dyn_ltree[i * 2] = 0;
  for (int i = 0; i < D_CODES; i++) s->dyn_dtree[i * 2] = 0;
  for (int i = 0; i < BL_CODES; i++) s->bl_tree[i * 2] = 0;
  /* NOTE: fillRange(0, lCodes * 2, 0) clears freq AND len fields of the first lCodes entries */
  memset(s->dyn_ltree, 0, L_CODES * 2 * sizeof(uint16_t));
  memset(s->dyn_dtree, 0, D_CODES * 2 * sizeof(uint16_t));
  memset(s->bl_tree, 0, BL_CODES * 2 * sizeof(uint16_t));
  s->dyn_ltree[END_BLOCK * 2] = 1;
  s->opt_len = s->static_len = 0;
  s->last_lit = s->matches = 0;
}
static int smaller(const uint16_t *tree, int n, int m, const uint8_t *depth) { /* _smaller :312-314 */
  return tree[n * 2] < tree[m * 2] || (tree[n * 2] == tree[m * 2] && depth[n] <= depth[m]);
}
static void pqdownheap(deflate_t *s, const uint16_t *tree, int k) { /* _pqdownheap :290-310 */
  int v = (int)s->heap[k];
  int j = k << 1;
  while (j <= s->heap_len) {
    if (j < s->heap_len && smaller(tree, (int)s->heap[j + 1], (int)s->heap[j], s->depth)) j++;
    if (smaller(tree, v, (int)s->heap[j], s->depth)) break;
    s->heap[k] = s->heap[j];
    k = j;
    j <<= 1;
  }
  s->heap[k] = (uint32_t)v;
}
static void gen_bitlen(deflate_t *s, tree_desc *desc) { /* _genBitlen :2567-2648 */
  uint16_t *tree = desc->dyn_tree;
  const uint16_t *stree = desc->stat->static_tree;
  const uint8_t *extra = desc->stat->extra_bits;
  int base = desc->stat->extra_base, max_length = desc->stat->max_length, max_code = desc->max_code;
  int h, n, m, bits, xbits, overflow = 0;
  int64_t f;
  for (bits = 0; bits <= MAX_BITS; bits++) s->bl_count[bits] = 0;
  tree[s->heap[s->heap_max] * 2 + 1] = 0;
  for (h = s->heap_max + 1; h < HEAP_SIZE; h++) {
    n = (int)s->heap[h];
    bits = tree[tree[n * 2 + 1] * 2 + 1] + 1;
    if (bits > max_length) {
      bits = max_length;
      overflow++;
    }
    tree[n * 2 + 1] = (uint16_t)bits;
    if (n > max_code) continue;
    s->bl_count[bits]++;
    xbits = 0;
    if (n >= base) xbits = extra[n - base];
    f = tree[n * 2];
    s->opt_len += f * (bits + xbits);
    if (stree) s->static_len += f * (stree[n * 2 + 1] + xbits);
  }
  if (overflow == 0) return;
  do {
    bits = max_length - 1;
    while (s->bl_count[bits] == 0) bits--;
    s->bl_count[bits]--;
    s->bl_count[bits + 1] = (uint16_t)(s->bl_count[bits + 1] + 2);
    s->bl_count[max_length]--;
    overflow -= 2;
  } while (overflow > 0);
  for (bits = max_length; bits != 0; bits--) {
    n = s->bl_count[bits];
    while (n != 0) {
      m = (int)s->heap[--h];
      if (m > max_code) continue;
      if (tree[m * 2 + 1] != bits) {
        s->opt_len = s->opt_len + ((int64_t)bits - tree[m * 2 + 1]) * tree[m * 2];
        tree[m * 2 + 1] = (uint16_t)bits;
      }
      n--;
    }
  }
}
static void build_tree(deflate_t *s, tree_desc *desc) { /* _buildTree :2656-2736 */
  uint16_t *tree = desc->dyn_tree;
  const uint16_t *stree = desc->stat->static_tree;
  int elems = desc->stat->elems;
  int n, m, max_code = -1, node;
  s->heap_len = 0;
  s->heap_max = HEAP_SIZE;
  for (n = 0; n < elems; n++) {
    if (tree[n * 2] != 0) {
      s->heap[++s->heap_len] = (uint32_t)(max_code = n);
      s->depth[n] = 0;
    } else {
      tree[n * 2 + 1] = 0;
    }
  }
  while (s->heap_len < 2) {
    node = (int)(s->heap[++s->heap_len] = (uint32_t)(max_code < 2 ? ++max_code : 0));
    tree[node * 2] = 1;
    s->depth[node] = 0;
    s->opt_len--;
    if (stree) s->static_len -= stree[node * 2 + 1];
  }
  desc->max_code = max_code;
  for (n = s->heap_len / 2; n >= 1; n--) pqdownheap(s, tree, n);
  node = elems;
  do {
    n = (int)s->heap[1];
    s->heap[1] = s->heap[s->heap_len--];
    pqdownheap(s, tree, 1);
    m = (int)s->heap[1];
    s->heap[--s->heap_max] = (uint32_t)n;
    s->heap[--s->heap_max] = (uint32_t)m;
    tree[node * 2] = (uint16_t)(tree[n * 2] + tree[m * 2]);
    s->depth[node] = (uint8_t)((s->depth[n] > s->depth[m] ? s->depth[n] : s->depth[m]) + 1);
    tree[n * 2 + 1] = tree[m * 2 + 1] = (uint16_t)node;
    s->heap[1] = (uint32_t)node++;
    pqdownheap(s, tree, 1);
  } while (s->heap_len >= 2);
  s->heap[--s->heap_max] = s->heap[1];
  gen_bitlen(s, desc);
  gen_codes(tree, max_code, s->bl_count);
}
static void scan_tree(deflate_t *s, uint16_t *tree, int max_code) { /* _scanTree :318-363 */
  int n, prevlen = -1, curlen, nextlen = tree[0 * 2 + 1], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  tree[(max_code + 1) * 2 + 1] = 0xffff;
  for (n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree[(n + 1) * 2 + 1];
    if (++count < max_count && curlen == nextlen) continue;
    else if (count < min_count) s->bl_tree[curlen * 2] = (uint16_t)(s->bl_tree[curlen * 2] + count);
    else if (curlen != 0) {
      if (curlen != prevlen) s->bl_tree[curlen * 2]++;
      s->bl_tree[REP_3_6 * 2]++;
    } else if (count <= 10) s->bl_tree[REPZ_3_10 * 2]++;
    else s->bl_tree[REPZ_11_138 * 2]++;
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}
static int build_bl_tree(deflate_t *s) { /* _buildBitLengthTree :367-392 */
  int max_blindex;
  scan_tree(s, s->dyn_ltree, s->l_desc.max_code);
  scan_tree(s, s->dyn_dtree, s->d_desc.max_code);
  build_tree(s, &s->bl_desc);
  for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
    if (s->bl_tree[bl_order[max_blindex] * 2 + 1] != 0) break;
  s->opt_len += 3 * (max_blindex + 1) + 5 + 5 + 4;
  return max_blindex;
}
static void send_tree(deflate_t *s, const uint16_t *tree, int max_code) { /* _sendTree :412-462 */
  int n, prevlen = -1, curlen, nextlen = tree[0 * 2 + 1], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  for (n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree[(n + 1) * 2 + 1];
    if (++count < max_count && curlen == nextlen) continue;
    else if (count < min_count) {
      do send_code(s, curlen, s->bl_tree);
      while (--count != 0);
    } else if (curlen != 0) {
      if (curlen != prevlen) {
        send_code(s, curlen, s->bl_tree);
        count--;
      }
      send_code(s, REP_3_6, s->bl_tree);
      send_bits(s, count - 3, 2);
    } else if (count <= 10) {
      send_code(s, REPZ_3_10, s->bl_tree);
      send_bits(s, count - 3, 3);
    } else {
      send_code(s, REPZ_11_138, s->bl_tree);
      send_bits(s, count - 11, 7);
    }
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}
static void send_all_trees(deflate_t *s, int lcodes, int dcodes, int blcodes) { /* _sendAllTrees :397-408 */
  send_bits(s, lcodes - 257, 5);
  send_bits(s, dcodes - 1, 5);
  send_bits(s, blcodes - 4, 4);
  for (int rank = 0; rank < blcodes; rank++) send_bits(s, s->bl_tree[bl_order[rank] * 2 + 1], 3);
  send_tree(s, s->dyn_ltree, lcodes - 1);
  send_tree(s, s->dyn_dtree, dcodes - 1);
}
static void compress_block(deflate_t *s, const uint16_t *ltree, const uint16_t *dtree) { /* _compressBlock :571-614 */
  int dist, lc, lx = 0, code, extra;
  if (s->last_lit != 0) {
    do {
      dist = ((s->pending_buf[s->d_buf + lx * 2] << 8) & 0xff00) | (s->pending_buf[s->d_buf + lx * 2 + 1] & 0xff);
      lc = s->pending_buf[s->l_buf + lx] & 0xff;
      lx++;
      if (dist == 0) {
        send_code(s, lc, ltree);
      } else {
        code = length_code_tab[lc];
        send_code(s, code + LITERALS + 1, ltree);
        extra = extra_lbits[code];
        if (extra != 0) {
          lc -= base_length[code];
          send_bits(s, lc, extra);
        }
        dist--;
        code = d_code(dist);
        send_code(s, code, dtree);
        extra = extra_dbits[code];
        if (extra != 0) {
          dist -= base_dist[code];
          send_bits(s, dist, extra);
        }
      }
    } while (lx < s->last_lit);
  }
  send_code(s, END_BLOCK, ltree);
  s->last_eob_len = ltree[END_BLOCK * 2 + 1];
}
static void set_data_type(deflate_t *s) { /* setDataType :621-637 */
  int n = 0;
  int64_t ascii_freq = 0, bin_freq = 0;
  while (n < 7) bin_freq += s->dyn_ltree[n++ * 2];
  while (n < 128) ascii_freq += s->dyn_ltree[n++ * 2];
  while (n < LITERALS) bin_freq += s->dyn_ltree[n++ * 2];
  s->data_type = bin_freq > (ascii_freq >> 2) ? Z_BINARY : Z_ASCII;
}
static void tr_stored_block(deflate_t *s, int buf, int stored_len, int eof) { /* _trStoredBlock :740-743 */
  send_bits(s, (STORED_BLOCK << 1) + (eof ? 1 : 0), 3);
  copy_block(s, buf, stored_len, 1);
}
static void tr_flush_block(deflate_t *s, int buf, int stored_len, int eof) { /* _trFlushBlock :747-807 */
  int64_t opt_lenb, static_lenb;
  int max_blindex = 0;
  if (s->level > 0) {
    if (s->data_type == Z_UNKNOWN) set_data_type(s);
    build_tree(s, &s->l_desc);
    build_tree(s, &s->d_desc);
    max_blindex = build_bl_tree(s);
    opt_lenb = (s->opt_len + 3 + 7) >> 3;
    static_lenb = (s->static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
  } else {
    opt_lenb = static_lenb = stored_len + 5;
  }
  if (stored_len + 4 <= opt_lenb && buf != -1) {
    tr_stored_block(s, buf, stored_len, eof);
  } else if (static_lenb == opt_lenb) {
    send_bits(s, (STATIC_TREES << 1) + (eof ? 1 : 0), 3);
    compress_block(s, static_ltree, static_dtree);
  } else {
    send_bits(s, (DYN_TREES << 1) + (eof ? 1 : 0), 3);
    send_all_trees(s, s->l_desc.max_code + 1, s->d_desc.max_code + 1, max_blindex + 1);
    compress_block(s, s->dyn_ltree, s->dyn_dtree);
  }
  init_block(s);
  if (eof) bi_windup(s);
}
static void flush_block_only(deflate_t *s, int eof) { /* _flushBlockOnly :677-682 */
  tr_flush_block(s, s->block_start >= 0 ? s->block_start : -1, s->strstart - s->block_start, eof);
  s->block_start = s->strstart;
  flush_pending(s);
}
/* Cross-check hook (tests only): with the heuristic off the output must equal stock zlib's, which compiles
 * TRUNCATE_BLOCK out.  The reference always has it ON. */
static int orc_truncate_heuristic = 1;
void orc_deflate_set_truncate_heuristic(int on) { orc_truncate_heuristic = on; }
static int tr_tally(deflate_t *s, int dist, int lc) { /* _trTally :531-568 */
  s->pending_buf[s->d_buf + s->last_lit * 2] = (uint8_t)(dist >> 8);
  s->pending_buf[s->d_buf + s->last_lit * 2 + 1] = (uint8_t)dist;
  s->pending_buf[s->l_buf + s->last_lit] = (uint8_t)lc;
  s->last_lit++;
  if (dist == 0) {
    s->dyn_ltree[lc * 2]++;
  } else {
    s->matches++;
    dist--;
    s->dyn_ltree[(length_code_tab[lc] + LITERALS + 1) * 2]++;
    s->dyn_dtree[d_code(dist) * 2]++;
  }
  if (orc_truncate_heuristic && (s->last_lit & 0x1fff) == 0 && s->level > 2) { /* TRUNCATE_BLOCK heuristic, compiled IN here */
    int64_t out_length = (int64_t)s->last_lit * 8;
    int64_t in_length = s->strstart - s->block_start;
    for (int dcode = 0; dcode < D_CODES; dcode++) out_length += (int64_t)s->dyn_dtree[dcode * 2] * (5 + extra_dbits[dcode]);
    out_length >>= 3;
    /* (_matches < _lastLit / 2) && outLength < inLength / 2  with Dart double division */
    if ((double)s->matches < (double)s->last_lit / 2.0 && (double)out_length < (double)in_length / 2.0) return 1;
  }
  return s->last_lit == s->lit_bufsize - 1;
}

/* ---- window / matching ------------------------------------------------------------------------ */
static int read_buf(deflate_t *s, int start, int size) { /* _readBuf :1214-1234 */
  if (size == 0 || in_eos(s)) return 0;
  int64_t len = s->in_len - s->in_pos;
  if (len > size) len = size;
  if (len == 0) return 0;
  memcpy(s->window + start, s->in + s->in_pos, (size_t)len);
  s->crc = orc_crc32(s->in + s->in_pos, (size_t)len, s->crc);
  s->in_pos += len;
  s->total += len;
  return (int)len;
}
static void fill_window(deflate_t *s) { /* _fillWindow :816-888 */
  do {
    int more = s->window_size - s->lookahead - s->strstart;
    if (more == 0 && s->strstart == 0 && s->lookahead == 0) {
      more = s->w_size;
    } else if (s->strstart >= s->w_size + s->w_size - MIN_LOOKAHEAD) {
      memcpy(s->window, s->window + s->w_size, (size_t)s->w_size);
      s->match_start -= s->w_size;
      s->strstart -= s->w_size;
      s->block_start -= s->w_size;
      int n = s->hash_size, p = n;
      do {
        int m = s->head[--p];
        s->head[p] = (uint16_t)(m >= s->w_size ? m - s->w_size : 0);
      } while (--n != 0);
      n = s->w_size;
      p = n;
      do {
        int m = s->prev[--p];
        s->prev[p] = (uint16_t)(m >= s->w_size ? m - s->w_size : 0);
      } while (--n != 0);
      more += s->w_size;
    }
    if (in_eos(s)) return;
    int n = read_buf(s, s->strstart + s->lookahead, more);
    s->lookahead += n;
    if (s->lookahead >= MIN_MATCH) {
      s->ins_h = s->window[s->strstart] & 0xff;
      s->ins_h = ((s->ins_h << s->hash_shift) ^ (s->window[s->strstart + 1] & 0xff)) & s->hash_mask;
    }
  } while (s->lookahead < MIN_LOOKAHEAD && !in_eos(s));
}
static int longest_match(deflate_t *s, int cur_match) { /* _longestMatch :1120-1206 */
  int chain_length = s->cfg.chain;
  int scan = s->strstart, match, len, best_len = s->prev_length;
  int limit = s->strstart > (s->w_size - MIN_LOOKAHEAD) ? s->strstart - (s->w_size - MIN_LOOKAHEAD) : 0;
  int nice_match = s->cfg.nice;
  int wmask = s->w_mask;
  int strend = s->strstart + MAX_MATCH;
  const uint8_t *w = s->window;
  uint8_t scan_end1 = w[scan + best_len - 1], scan_end = w[scan + best_len];
  if (s->prev_length >= s->cfg.good) chain_length >>= 2;
  if (nice_match > s->lookahead) nice_match = s->lookahead;
  do {
    match = cur_match;
    if (w[match + best_len] != scan_end || w[match + best_len - 1] != scan_end1 || w[match] != w[scan] ||
        w[++match] != w[scan + 1])
      continue;
    scan += 2;
    match++;
    do {
    } while (w[++scan] == w[++match] && w[++scan] == w[++match] && w[++scan] == w[++match] &&
             w[++scan] == w[++match] && w[++scan] == w[++match] && w[++scan] == w[++match] &&
             w[++scan] == w[++match] && w[++scan] == w[++match] && scan < strend);
    len = MAX_MATCH - (strend - scan);
    scan = strend - MAX_MATCH;
    if (len > best_len) {
      s->match_start = cur_match;
      best_len = len;
      if (len >= nice_match) break;
      scan_end1 = w[scan + best_len - 1];
      scan_end = w[scan + best_len];
    }
  } while ((cur_match = s->prev[cur_match & wmask]) > limit && --chain_length != 0);
  if (best_len <= s->lookahead) return best_len;
  return s->lookahead;
}
#define INSERT_STRING(s, hash_head)                                                                           \
  do {                                                                                                        \
    (s)->ins_h = (((s)->ins_h << (s)->hash_shift) ^ ((s)->window[(s)->strstart + (MIN_MATCH - 1)] & 0xff)) & \
                 (s)->hash_mask;                                                                              \
    (hash_head) = (s)->head[(s)->ins_h];                                                                      \
    (s)->prev[(s)->strstart & (s)->w_mask] = (s)->head[(s)->ins_h];                                           \
    (s)->head[(s)->ins_h] = (uint16_t)(s)->strstart;                                                          \
  } while (0)

static int deflate_stored(deflate_t *s) { /* _deflateStored :691-737 (flush == finish) */
  int max_block_size = 0xffff;
  if (max_block_size > s->pending_buf_size - 5) max_block_size = s->pending_buf_size - 5;
  for (;;) {
    if (s->lookahead <= 1) {
      fill_window(s);
      if (s->lookahead == 0) break;
    }
    s->strstart += s->lookahead;
    s->lookahead = 0;
    int max_start = s->block_start + max_block_size;
    if (s->strstart >= max_start) {
      s->lookahead = s->strstart - max_start;
      s->strstart = max_start;
      flush_block_only(s, 0);
    }
    if (s->strstart - s->block_start >= s->w_size - MIN_LOOKAHEAD) flush_block_only(s, 0);
  }
  flush_block_only(s, 1);
  return FINISH_DONE;
}
static int deflate_fast(deflate_t *s) { /* _deflateFast :895-992 */
  int hash_head = 0, bflush;
  for (;;) {
    if (s->lookahead < MIN_LOOKAHEAD) {
      fill_window(s);
      if (s->lookahead == 0) break;
    }
    if (s->lookahead >= MIN_MATCH) INSERT_STRING(s, hash_head);
    if (hash_head != 0 && ((s->strstart - hash_head) & 0xffff) <= s->w_size - MIN_LOOKAHEAD) {
      if (s->strategy != 2) s->match_length = longest_match(s, hash_head);
    }
    if (s->match_length >= MIN_MATCH) {
      bflush = tr_tally(s, s->strstart - s->match_start, s->match_length - MIN_MATCH);
      s->lookahead -= s->match_length;
      if (s->match_length <= s->cfg.lazy && s->lookahead >= MIN_MATCH) {
        s->match_length--;
        do {
          s->strstart++;
          INSERT_STRING(s, hash_head);
        } while (--s->match_length != 0);
        s->strstart++;
      } else {
        s->strstart += s->match_length;
        s->match_length = 0;
        s->ins_h = s->window[s->strstart] & 0xff;
        s->ins_h = ((s->ins_h << s->hash_shift) ^ (s->window[s->strstart + 1] & 0xff)) & s->hash_mask;
      }
    } else {
      bflush = tr_tally(s, 0, s->window[s->strstart] & 0xff);
      s->lookahead--;
      s->strstart++;
    }
    if (bflush) flush_block_only(s, 0);
  }
  flush_block_only(s, 1);
  return FINISH_DONE;
}
static int deflate_slow(deflate_t *s) { /* _deflateSlow :997-1118 */
  int hash_head = 0, bflush;
  for (;;) {
    if (s->lookahead < MIN_LOOKAHEAD) {
      fill_window(s);
      if (s->lookahead == 0) break;
    }
    if (s->lookahead >= MIN_MATCH) INSERT_STRING(s, hash_head);
    s->prev_length = s->match_length;
    s->prev_match = s->match_start;
    s->match_length = MIN_MATCH - 1;
    if (hash_head != 0 && s->prev_length < s->cfg.lazy &&
        ((s->strstart - hash_head) & 0xffff) <= s->w_size - MIN_LOOKAHEAD) {
      if (s->strategy != 2) s->match_length = longest_match(s, hash_head);
      if (s->match_length <= 5 &&
          (s->strategy == 1 || (s->match_length == MIN_MATCH && s->strstart - s->match_start > 4096)))
        s->match_length = MIN_MATCH - 1;
    }
    if (s->prev_length >= MIN_MATCH && s->match_length <= s->prev_length) {
      int max_insert = s->strstart + s->lookahead - MIN_MATCH;
      bflush = tr_tally(s, s->strstart - 1 - s->prev_match, s->prev_length - MIN_MATCH);
      s->lookahead -= s->prev_length - 1;
      s->prev_length -= 2;
      do {
        if (++s->strstart <= max_insert) INSERT_STRING(s, hash_head);
      } while (--s->prev_length != 0);
      s->match_available = 0;
      s->match_length = MIN_MATCH - 1;
      s->strstart++;
      if (bflush) flush_block_only(s, 0);
    } else if (s->match_available != 0) {
      bflush = tr_tally(s, 0, s->window[s->strstart - 1] & 0xff);
      if (bflush) flush_block_only(s, 0);
      s->strstart++;
      s->lookahead--;
    } else {
      s->match_available = 1;
      s->strstart++;
      s->lookahead--;
    }
  }
  if (s->match_available != 0) {
    tr_tally(s, 0, s->window[s->strstart - 1] & 0xff);
    s->match_available = 0;
  }
  flush_block_only(s, 1);
  return FINISH_DONE;
}

/* Deflate(bytes, level:, windowBits:) ctor path: _init :102-169 + _deflate(finish) :172-239 */
static int deflate_run(const uint8_t *in, size_t n, int level, int window_bits, orc_oms *out, uint32_t *crc) {
  const int mem_level = 8;
  if (window_bits < 9 || window_bits > 15 || level < 0 || level > 9) return ORC_THROW; /* LateInitializationError */
  tables_init();
  deflate_t *s = (deflate_t *)calloc(1, sizeof(deflate_t));
  s->in = in;
  s->in_len = (int64_t)n;
  s->out = out;
  s->cfg = config_table[level];
  s->w_bits = window_bits;
  s->w_size = 1 << window_bits;
  s->w_mask = s->w_size - 1;
  s->hash_bits = mem_level + 7;
  s->hash_size = 1 << s->hash_bits;
  s->hash_mask = s->hash_size - 1;
  s->hash_shift = (s->hash_bits + MIN_MATCH - 1) / MIN_MATCH;
  s->window = (uint8_t *)calloc((size_t)s->w_size * 2 + 8, 1);
  s->prev = (uint16_t *)calloc((size_t)s->w_size, sizeof(uint16_t));
  s->head = (uint16_t *)calloc((size_t)s->hash_size, sizeof(uint16_t));
  s->lit_bufsize = 1 << (mem_level + 6);
  s->pending_buf = (uint8_t *)calloc((size_t)s->lit_bufsize * 4 + 8, 1);
  s->pending_buf_size = s->lit_bufsize * 4;
  s->d_buf = s->lit_bufsize;
  s->l_buf = (1 + 2) * s->lit_bufsize;
  s->level = level;
  s->strategy = 0;
  s->data_type = Z_UNKNOWN;
  /* _trInit :255-276 */
  s->l_desc.dyn_tree = s->dyn_ltree;
  s->l_desc.stat = &static_l_desc;
  s->d_desc.dyn_tree = s->dyn_dtree;
  s->d_desc.stat = &static_d_desc;
  s->bl_desc.dyn_tree = s->bl_tree;
  s->bl_desc.stat = &static_bl_desc;
  s->bi_buf = 0;
  s->bi_valid = 0;
  s->last_eob_len = 8;
  init_block(s);
  /* _lmInit :241-252 */
  s->window_size = 2 * s->w_size;
  s->strstart = s->block_start = s->lookahead = 0;
  s->match_length = s->prev_length = MIN_MATCH - 1;
  s->match_available = 0;
  s->ins_h = 0;
  s->match_start = 0;
  switch (s->cfg.func) {
    case FN_STORED: deflate_stored(s); break;
    case FN_FAST: deflate_fast(s); break;
    default: deflate_slow(s); break;
  }
  flush_pending(s); /* getBytes() :72-75 */
  if (crc) *crc = s->crc;
  free(s->window);
  free(s->prev);
  free(s->head);
  free(s->pending_buf);
  free(s);
  return ORC_OK;
}

int orc_deflate_bytes(const uint8_t *in, size_t n, int level, int window_bits, uint8_t **out, size_t *out_len,
                      uint32_t *crc32_of_input) {
  orc_oms o;
  orc_oms_init(&o, 0x8000);
  int st = deflate_run(in, n, level, window_bits, &o, crc32_of_input);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}

/* ZLibEncoderWeb.encodeBytes  _zlib_encoder_web.dart:17-73 (FLEVEL 0 -> "78 01" at windowBits 15, quirk Q4) */
int orc_zlib_encode_bytes(const uint8_t *in, size_t n, int level, int window_bits, int raw, uint8_t **out,
                          size_t *out_len) {
  orc_oms o;
  orc_oms_init(&o, 0x8000);
  int st = ORC_OK;
  if (raw) {
    st = deflate_run(in, n, level, window_bits, &o, NULL);
  } else {
    int wb = window_bits < 0 ? 0 : window_bits > 15 ? 15 : window_bits;
    int cmf = ((wb - 8) << 4) | 8;
    orc_oms_write_byte(&o, cmf);
    int flag = 0, fcheck = 0;
    while ((cmf * 256 + (flag | fcheck)) % 31 != 0) fcheck++;
    flag |= fcheck;
    orc_oms_write_byte(&o, flag);
    uint32_t ad = orc_adler32(in, n, 1);
    st = deflate_run(in, n, level, window_bits, &o, NULL);
    orc_oms_write_byte(&o, (ad >> 24) & 0xff);
    orc_oms_write_byte(&o, (ad >> 16) & 0xff);
    orc_oms_write_byte(&o, (ad >> 8) & 0xff);
    orc_oms_write_byte(&o, ad & 0xff);
  }
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}

/* GZipEncoderWeb.encodeBytes  _gzip_encoder_web.dart:17-100 (MTIME = now in the reference: a parameter here) */
int orc_gzip_encode_bytes(const uint8_t *in, size_t n, int level, uint32_t mtime, uint8_t **out, size_t *out_len) {
  orc_oms o;
  orc_oms_init(&o, 0x8000);
  orc_oms_write_byte(&o, 0x1f);
  orc_oms_write_byte(&o, 0x8b);
  orc_oms_write_byte(&o, 8);
  orc_oms_write_byte(&o, 0);
  for (int i = 0; i < 4; ++i) orc_oms_write_byte(&o, (mtime >> (8 * i)) & 0xff);
  orc_oms_write_byte(&o, 0);
  orc_oms_write_byte(&o, 255);
  uint32_t crc = 0;
  int st = deflate_run(in, n, level, 15, &o, &crc);
  for (int i = 0; i < 4; ++i) orc_oms_write_byte(&o, (crc >> (8 * i)) & 0xff);
  for (int i = 0; i < 4; ++i) orc_oms_write_byte(&o, ((uint32_t)n >> (8 * i)) & 0xff);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}
