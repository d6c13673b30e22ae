// tests/host_emul/deflate_model.cpp -- TEST INFRASTRUCTURE.
//
// A CPU model of the DATA-PARALLEL REFORMULATION of the reference's Deflate (deflate.dart, levels 4-9) that the
// CUDA encoder (archive_b200/csrc/deflate_kernels.cu) implements.  It is checked byte-for-byte against the oracle
// (oracle/deflate.c, the line-by-line restatement) so that the reformulation itself is proven before any
// kernel is written, and it documents each stage's exact semantics:
//
//   S1 match table   for every position p: the result of _longestMatch (deflate.dart:1120-1206) under the two
//                    chain budgets it can be called with (maxChain, and maxChain >> 2 when prevLength >=
//                    goodLength).  At levels >= 4 every position is inserted in the hash chains in order
//                    (:1022-1029,1072-1081) and the walk stops at strStart - (wSize - MIN_LOOKAHEAD), so the
//                    candidates -- and therefore both results -- are a pure function of the data.
//   S2 parse graph   the lazy-match state machine (_deflateSlow :997-1118) has only four states per position
//                    once S1 is known: R (nothing pending), A2 (a literal pending, no match at p-1), A128 /
//                    A32 (a literal pending, match found at p-1 with the full / the reduced budget).  Every
//                    (position, state) node has exactly one successor; the parse is the path from (0, R).
//   S3 block cuts    _trTally (:531-568): flush at lastLit == 16383, or -- the reference's compiled-in
//                    TRUNCATE_BLOCK heuristic -- at lastLit % 8192 == 0 when matches < lastLit/2 and the
//                    estimated size < inLength/2.  Computable from prefix sums over the token stream.
//   S4 trees + bits  per block, exactly _trFlushBlock (:747-807).
//
// The model produces the raw DEFLATE stream; tests compare it with orc_deflate_bytes.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace {

enum { MIN_MATCH = 3, MAX_MATCH = 258, MIN_LOOKAHEAD = 262, WSIZE = 32768, MAX_DIST = WSIZE - MIN_LOOKAHEAD, TOO_FAR = 4096 };

struct Cfg {
  int good, lazy, nice, chain;
};
const Cfg kCfg[10] = {{0, 0, 0, 0},      {4, 4, 8, 4},      {4, 5, 16, 8},      {4, 6, 32, 32},     {4, 4, 16, 16},
                      {8, 16, 32, 32},   {8, 16, 128, 128}, {8, 32, 128, 256}, {32, 128, 258, 1024}, {32, 258, 258, 4096}};

struct Match {
  uint16_t len;   // 0 = no candidate beat length 2
  uint16_t dist;
};

inline uint32_t hash3(const uint8_t *p) { return (((uint32_t)(p[0] & 31) << 10) ^ ((uint32_t)p[1] << 5) ^ p[2]) & 0x7fff; }

// S1: both budgets in one chain walk.  `full` = result after `chain` candidates, `reduced` after `chain >> 2`.
void match_table(const uint8_t *d, size_t n, const Cfg &c, int64_t max_dist, std::vector<Match> &full, std::vector<Match> &reduced) {
  full.assign(n, Match{0, 0});
  reduced.assign(n, Match{0, 0});
  std::vector<int64_t> head(32768, -1), prev(n, -1);
  const int budget_full = c.chain, budget_red = c.chain >> 2;
  for (size_t p = 0; p + MIN_MATCH <= n; ++p) {
    uint32_t h = hash3(d + p);
    int64_t cur = head[h];
    prev[p] = cur;
    head[h] = (int64_t)p;
    const int64_t limit = (int64_t)p > max_dist ? (int64_t)p - max_dist : 0;
    const size_t lookahead = n - p;
    const int nice = (size_t)c.nice > lookahead ? (int)lookahead : c.nice;
    int best = MIN_MATCH - 1;
    int64_t best_pos = -1;
    int steps = 0;
    bool red_done = budget_red == 0;
    // first candidate: hashHead != 0 && strStart - hashHead <= MAX_DIST (:1036-1038); later ones: > limit (:1198)
    bool first = true;
    while (first ? (cur >= 1 && (int64_t)p - cur <= max_dist) : (cur > limit)) {
      first = false;
      // candidate length (exact common prefix, capped at MAX_MATCH and at the end of the data)
      size_t maxl = lookahead < (size_t)MAX_MATCH ? lookahead : (size_t)MAX_MATCH;
      size_t l = 0;
      const uint8_t *a = d + p, *b = d + cur;
      while (l < maxl && a[l] == b[l]) ++l;
      bool stop = false;
      if ((int)l > best) {
        best = (int)l;
        best_pos = cur;
        if ((int)l >= nice) stop = true;
      }
      ++steps;
      if (!red_done && (steps == budget_red || stop)) {
        reduced[p] = best >= MIN_MATCH ? Match{(uint16_t)best, (uint16_t)(p - best_pos)} : Match{0, 0};
        red_done = true;
      }
      if (stop || steps == budget_full) break;
      cur = prev[cur];
    }
    if (!red_done) reduced[p] = best >= MIN_MATCH ? Match{(uint16_t)best, (uint16_t)(p - best_pos)} : Match{0, 0};
    full[p] = best >= MIN_MATCH ? Match{(uint16_t)best, (uint16_t)(p - best_pos)} : Match{0, 0};
    // _longestMatch returns min(bestLen, lookAhead): only matters in the last bytes, where it kills the match
    auto clampm = [&](Match &m) {
      if (m.len > lookahead) m.len = (uint16_t)lookahead;
      if (m.len < MIN_MATCH) m = Match{0, 0};
    };
    clampm(full[p]);
    clampm(reduced[p]);
  }
}

// token: dist == 0 -> literal lc, else match (lc = len - 3)
struct Token {
  uint16_t dist;
  uint8_t lc;
  uint32_t pos_after;  // strStart when _trTally is called (block accounting)
};

enum State { R = 0, A2 = 1, AF = 2, AR = 3 };  // AF: match from the full budget pending, AR: from the reduced one

// S2: the lazy parse as a walk over (position, state) nodes.
void parse(const uint8_t *d, size_t n, const Cfg &c, const std::vector<Match> &full, const std::vector<Match> &reduced,
           std::vector<Token> &toks, std::vector<uint32_t> &tally_strstart, std::vector<uint32_t> &next_strstart,
           bool &last_is_pending_literal) {
  last_is_pending_literal = false;
  size_t p = 0;
  State s = R;
  while (p < n) {
    // state -> (match_available, prev_length, prev match)
    int pl = 2;
    Match pm{0, 0};
    if (s == AF) pm = full[p - 1];
    if (s == AR) pm = reduced[p - 1];
    if (s == AF || s == AR) pl = pm.len;
    const bool avail = s != R;
    // current match (:1031-1056)
    int cl = 2;
    Match cm{0, 0};
    State ns = A2;
    if (pl < c.lazy && p + MIN_MATCH <= n) {
      const bool red = pl >= c.good;
      cm = red ? reduced[p] : full[p];
      if (cm.len > pl) {
        cl = cm.len;
        ns = red ? AR : AF;
        if (cl == MIN_MATCH && cm.dist > TOO_FAR) {
          cl = 2;
          ns = A2;
        }
      } else {
        cl = pl;  // _longestMatch returns bestLen = prevLength: "not better"
      }
    }
    if (pl >= MIN_MATCH && cl <= pl) {
      // emit the previous match (:1059-1090)
      toks.push_back(Token{pm.dist, (uint8_t)(pl - MIN_MATCH), 0});
      tally_strstart.push_back((uint32_t)p);
      p = p + pl - 1;
      next_strstart.push_back((uint32_t)p);
      s = R;
    } else if (avail) {
      toks.push_back(Token{0, d[p - 1], 0});
      tally_strstart.push_back((uint32_t)p);
      next_strstart.push_back((uint32_t)p);  // a flush here happens BEFORE strStart++ (:1097-1101)
      p++;
      s = ns;
    } else {
      p++;
      s = ns;
    }
  }
  if (s != R) {  // :1110-1113 the pending literal
    toks.push_back(Token{0, d[n - 1], 0});
    tally_strstart.push_back((uint32_t)n);
    next_strstart.push_back((uint32_t)n);
    last_is_pending_literal = true;  // its _trTally result is ignored (:1111)
  }
}

// ---------------- S4: trees + bit output, as the reference does per block ----------------
enum { MAX_BITS = 15, BL_CODES = 19, D_CODES = 30, LITERALS = 256, LENGTH_CODES = 29, L_CODES = 286, HEAP_SIZE = 573, END_BLOCK = 256 };
const uint8_t extra_lbits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint8_t extra_dbits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t extra_blbits[19] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
const uint8_t bl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
uint8_t length_code[256], dist_code[512];
int base_length[29], base_dist[30];
uint16_t static_l[288 * 2], static_d[30 * 2];
bool tabs;
unsigned bitrev(unsigned c, int len) {
  unsigned r = 0;
  do {
    r |= c & 1;
    c >>= 1;
    r <<= 1;
  } while (--len > 0);
  return r >> 1;
}
void gen_codes(uint16_t *tree, int max_code, const uint16_t *bl_count) {
  uint16_t next[MAX_BITS + 1];
  unsigned code = 0;
  for (int b = 1; b <= MAX_BITS; b++) next[b] = (uint16_t)(code = (code + bl_count[b - 1]) << 1);
  for (int n = 0; n <= max_code; n++) {
    int len = tree[n * 2 + 1];
    if (len) tree[n * 2] = (uint16_t)bitrev(next[len]++, len);
  }
}
void init_tabs() {
  if (tabs) return;
  int length = 0, code, n, dist = 0;
  for (code = 0; code < 28; code++) {
    base_length[code] = length;
    for (n = 0; n < (1 << extra_lbits[code]); n++) length_code[length++] = (uint8_t)code;
  }
  length_code[length - 1] = 28;
  base_length[28] = 0;
  for (code = 0; code < 16; code++) {
    base_dist[code] = dist;
    for (n = 0; n < (1 << extra_dbits[code]); n++) dist_code[dist++] = (uint8_t)code;
  }
  dist >>= 7;
  for (; code < 30; code++) {
    base_dist[code] = dist << 7;
    for (n = 0; n < (1 << (extra_dbits[code] - 7)); n++) dist_code[256 + dist++] = (uint8_t)code;
  }
  uint16_t blc[16] = {0};
  for (n = 0; n < 288; n++) {
    int l = n < 144 ? 8 : n < 256 ? 9 : n < 280 ? 7 : 8;
    static_l[n * 2 + 1] = (uint16_t)l;
    blc[l]++;
  }
  gen_codes(static_l, 287, blc);
  for (n = 0; n < 30; n++) {
    static_d[n * 2 + 1] = 5;
    static_d[n * 2] = (uint16_t)bitrev(n, 5);
  }
  tabs = true;
}
inline int d_code(int dist) { return dist < 256 ? dist_code[dist] : dist_code[256 + (dist >> 7)]; }

struct TreeBuilder {  // _HuffmanTree._buildTree / _genBitlen (deflate.dart:2567-2736)
  uint32_t heap[HEAP_SIZE];
  int heap_len, heap_max;
  uint8_t depth[HEAP_SIZE];
  uint16_t bl_count[MAX_BITS + 1];
  int64_t opt_len = 0, static_len = 0;
  bool smaller(const uint16_t *t, int n, int m) { return t[n * 2] < t[m * 2] || (t[n * 2] == t[m * 2] && depth[n] <= depth[m]); }
  void down(const uint16_t *t, int k) {
    int v = heap[k], j = k << 1;
    while (j <= heap_len) {
      if (j < heap_len && smaller(t, heap[j + 1], heap[j])) j++;
      if (smaller(t, v, heap[j])) break;
      heap[k] = heap[j];
      k = j;
      j <<= 1;
    }
    heap[k] = v;
  }
  int build(uint16_t *tree, int elems, const uint16_t *stree, const uint8_t *extra, int base, int max_length) {
    int n, m, max_code = -1, node;
    heap_len = 0;
    heap_max = HEAP_SIZE;
    for (n = 0; n < elems; n++) {
      if (tree[n * 2]) {
        heap[++heap_len] = max_code = n;
        depth[n] = 0;
      } else
        tree[n * 2 + 1] = 0;
    }
    while (heap_len < 2) {
      node = heap[++heap_len] = (max_code < 2 ? ++max_code : 0);
      tree[node * 2] = 1;
      depth[node] = 0;
      opt_len--;
      if (stree) static_len -= stree[node * 2 + 1];
    }
    for (n = heap_len / 2; n >= 1; n--) down(tree, n);
    node = elems;
    do {
      n = heap[1];
      heap[1] = heap[heap_len--];
      down(tree, 1);
      m = heap[1];
      heap[--heap_max] = n;
      heap[--heap_max] = m;
      tree[node * 2] = (uint16_t)(tree[n * 2] + tree[m * 2]);
      depth[node] = (uint8_t)((depth[n] > depth[m] ? depth[n] : depth[m]) + 1);
      tree[n * 2 + 1] = tree[m * 2 + 1] = (uint16_t)node;
      heap[1] = node++;
      down(tree, 1);
    } while (heap_len >= 2);
    heap[--heap_max] = heap[1];
    // gen_bitlen
    int h, bits, xbits, overflow = 0;
    for (bits = 0; bits <= MAX_BITS; bits++) bl_count[bits] = 0;
    tree[heap[heap_max] * 2 + 1] = 0;
    for (h = heap_max + 1; h < HEAP_SIZE; h++) {
      n = heap[h];
      bits = tree[tree[n * 2 + 1] * 2 + 1] + 1;
      if (bits > max_length) bits = max_length, overflow++;
      tree[n * 2 + 1] = (uint16_t)bits;
      if (n > max_code) continue;
      bl_count[bits]++;
      xbits = n >= base ? extra[n - base] : 0;
      int64_t f = tree[n * 2];
      opt_len += f * (bits + xbits);
      if (stree) static_len += f * (stree[n * 2 + 1] + xbits);
    }
    if (overflow) {
      do {
        bits = max_length - 1;
        while (bl_count[bits] == 0) bits--;
        bl_count[bits]--;
        bl_count[bits + 1] += 2;
        bl_count[max_length]--;
        overflow -= 2;
      } while (overflow > 0);
      for (bits = max_length; bits != 0; bits--) {
        n = bl_count[bits];
        while (n != 0) {
          m = heap[--h];
          if (m > max_code) continue;
          if (tree[m * 2 + 1] != bits) {
            opt_len += ((int64_t)bits - tree[m * 2 + 1]) * tree[m * 2];
            tree[m * 2 + 1] = (uint16_t)bits;
          }
          n--;
        }
      }
    }
    gen_codes(tree, max_code, bl_count);
    return max_code;
  }
};

struct BitOut {
  std::vector<uint8_t> &o;
  uint64_t acc = 0;
  int nb = 0;
  explicit BitOut(std::vector<uint8_t> &v) : o(v) {}
  void put(unsigned v, int n) {
    acc |= (uint64_t)v << nb;
    nb += n;
    while (nb >= 8) {
      o.push_back((uint8_t)acc);
      acc >>= 8;
      nb -= 8;
    }
  }
  void align() {
    if (nb > 0) {
      o.push_back((uint8_t)acc);
      acc = 0;
      nb = 0;
    }
  }
};

void scan_tree(uint16_t *bl, uint16_t *tree, int max_code) {
  int prevlen = -1, curlen, nextlen = tree[1], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  tree[(max_code + 1) * 2 + 1] = 0xffff;
  for (int n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree[(n + 1) * 2 + 1];
    if (++count < max_count && curlen == nextlen) continue;
    else if (count < min_count) bl[curlen * 2] = (uint16_t)(bl[curlen * 2] + count);
    else if (curlen != 0) {
      if (curlen != prevlen) bl[curlen * 2]++;
      bl[16 * 2]++;
    } else if (count <= 10) bl[17 * 2]++;
    else bl[18 * 2]++;
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}
void send_tree(BitOut &b, const uint16_t *bl, const uint16_t *tree, int max_code) {
  int prevlen = -1, curlen, nextlen = tree[1], count = 0, max_count = 7, min_count = 4;
  if (nextlen == 0) max_count = 138, min_count = 3;
  auto code = [&](int c) { b.put(bl[c * 2], bl[c * 2 + 1]); };
  for (int n = 0; n <= max_code; n++) {
    curlen = nextlen;
    nextlen = tree[(n + 1) * 2 + 1];
    if (++count < max_count && curlen == nextlen) continue;
    else if (count < min_count) {
      do code(curlen);
      while (--count != 0);
    } else if (curlen != 0) {
      if (curlen != prevlen) {
        code(curlen);
        count--;
      }
      code(16);
      b.put(count - 3, 2);
    } else if (count <= 10) {
      code(17);
      b.put(count - 3, 3);
    } else {
      code(18);
      b.put(count - 11, 7);
    }
    count = 0;
    prevlen = curlen;
    if (nextlen == 0) max_count = 138, min_count = 3;
    else if (curlen == nextlen) max_count = 6, min_count = 3;
    else max_count = 7, min_count = 4;
  }
}
void compress_block(BitOut &b, const Token *t, size_t nt, const uint16_t *lt, const uint16_t *dt) {
  for (size_t i = 0; i < nt; ++i) {
    if (t[i].dist == 0) {
      b.put(lt[t[i].lc * 2], lt[t[i].lc * 2 + 1]);
    } else {
      int lc = t[i].lc, code = length_code[lc];
      b.put(lt[(code + 257) * 2], lt[(code + 257) * 2 + 1]);
      if (extra_lbits[code]) b.put(lc - base_length[code], extra_lbits[code]);
      int dist = t[i].dist - 1;
      code = d_code(dist);
      b.put(dt[code * 2], dt[code * 2 + 1]);
      if (extra_dbits[code]) b.put(dist - base_dist[code], extra_dbits[code]);
    }
  }
  b.put(lt[END_BLOCK * 2], lt[END_BLOCK * 2 + 1]);
}

}  // namespace

extern "C" int model_deflate_wb(const uint8_t *d, size_t n, int level, int window_bits, uint8_t **out, size_t *out_len,
                                uint64_t *n_tokens, uint64_t *n_blocks) {
  if (level < 4 || level > 9) return -1;
  init_tabs();
  const Cfg &c = kCfg[level];
  std::vector<Match> full, red;
  match_table(d, n, c, ((int64_t)1 << window_bits) - MIN_LOOKAHEAD, full, red);
  std::vector<Token> toks;
  std::vector<uint32_t> tally_ss, next_ss;
  bool last_pending = false;
  parse(d, n, c, full, red, toks, tally_ss, next_ss, last_pending);

  // S3 + S4: cut blocks with the _trTally rules and emit each as _trFlushBlock does.  data_type is set once, at
  // the first flush (it does not influence the output), the bit buffer runs on across blocks.
  std::vector<uint8_t> o;
  BitOut bo(o);
  size_t t0 = 0;
  uint32_t block_start = 0;
  uint64_t blocks = 0;
  const size_t NT = toks.size();
  // window base (absolute position of window index 0) at the iteration that starts at position p: _fillWindow slides by
  // wSize whenever lookahead < MIN_LOOKAHEAD and strStart >= 2 wSize - MIN_LOOKAHEAD (:816-857).  Blocks whose start has
  // slid out (blockStart < 0 -> buf == -1, :677-680) cannot be stored.
  const int64_t W = (int64_t)1 << window_bits;
  auto window_base = [&](int64_t p) {
    int64_t base = 0;
    for (;;) {
      int64_t loaded = base + 2 * W < (int64_t)n ? base + 2 * W : (int64_t)n;
      if (loaded - p < MIN_LOOKAHEAD && p - base >= 2 * W - MIN_LOOKAHEAD) base += W;
      else break;
    }
    return base;
  };
  auto flush = [&](size_t t1, uint32_t strstart, bool eof, int64_t iter_pos) {
    uint16_t lt[HEAP_SIZE * 2], dt[(2 * D_CODES + 1) * 2], bl[(2 * BL_CODES + 1) * 2];
    memset(lt, 0, sizeof lt);
    memset(dt, 0, sizeof dt);
    memset(bl, 0, sizeof bl);
    lt[END_BLOCK * 2] = 1;
    for (size_t i = t0; i < t1; ++i) {
      if (toks[i].dist == 0) lt[toks[i].lc * 2]++;
      else {
        lt[(length_code[toks[i].lc] + 257) * 2]++;
        dt[d_code(toks[i].dist - 1) * 2]++;
      }
    }
    TreeBuilder tb;
    int lmax = tb.build(lt, L_CODES, static_l, extra_lbits, 257, MAX_BITS);
    int dmax = tb.build(dt, D_CODES, static_d, extra_dbits, 0, MAX_BITS);
    scan_tree(bl, lt, lmax);
    scan_tree(bl, dt, dmax);
    tb.build(bl, BL_CODES, nullptr, extra_blbits, 0, 7);
    int max_bl;
    for (max_bl = BL_CODES - 1; max_bl >= 3; max_bl--)
      if (bl[bl_order[max_bl] * 2 + 1]) break;
    tb.opt_len += 3 * (max_bl + 1) + 5 + 5 + 4;
    int64_t opt_lenb = (tb.opt_len + 3 + 7) >> 3, static_lenb = (tb.static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    const int64_t stored_len = (int64_t)strstart - block_start;
    if (stored_len + 4 <= opt_lenb && (int64_t)block_start >= window_base(iter_pos)) {
      bo.put(0 + (eof ? 1 : 0), 3);
      bo.align();
      o.push_back((uint8_t)stored_len);
      o.push_back((uint8_t)(stored_len >> 8));
      o.push_back((uint8_t)~stored_len);
      o.push_back((uint8_t)(~stored_len >> 8));
      o.insert(o.end(), d + block_start, d + block_start + stored_len);
    } else if (static_lenb == opt_lenb) {
      bo.put(2 + (eof ? 1 : 0), 3);
      compress_block(bo, toks.data() + t0, t1 - t0, static_l, static_d);
    } else {
      bo.put(4 + (eof ? 1 : 0), 3);
      bo.put(lmax + 1 - 257, 5);
      bo.put(dmax + 1 - 1, 5);
      bo.put(max_bl + 1 - 4, 4);
      for (int r = 0; r <= max_bl; r++) bo.put(bl[bl_order[r] * 2 + 1], 3);
      send_tree(bo, bl, lt, lmax);
      send_tree(bo, bl, dt, dmax);
      compress_block(bo, toks.data() + t0, t1 - t0, lt, dt);
    }
    if (eof) bo.align();
    t0 = t1;
    block_start = strstart;
    blocks++;
  };
  uint64_t matches = 0;
  uint64_t dsum = 0;  // sum over the block's matches of (5 + extra_dbits)
  for (size_t i = 0; i < NT; ++i) {
    if (toks[i].dist) {
      matches++;
      dsum += 5 + extra_dbits[d_code(toks[i].dist - 1)];
    }
    const uint64_t last_lit = i + 1 - t0;
    bool fl = false;
    if ((last_lit & 0x1fff) == 0) {  // level > 2 always here
      uint64_t out_length = (last_lit * 8 + dsum) >> 3;
      uint64_t in_length = (uint64_t)tally_ss[i] - block_start;
      if ((double)matches < (double)last_lit / 2.0 && (double)out_length < (double)in_length / 2.0) fl = true;
    }
    if (last_lit == 16383) fl = true;
    // the very last token is followed by the final flush, which takes precedence only if no flush fired here
    if (fl && !(i + 1 == NT && last_pending)) {
      flush(i + 1, next_ss[i], false, tally_ss[i]);
      matches = 0;
      dsum = 0;
    }
  }
  flush(NT, (uint32_t)n, true, (int64_t)n);
  *out = (uint8_t *)malloc(o.size() ? o.size() : 1);
  memcpy(*out, o.data(), o.size());
  *out_len = o.size();
  if (n_tokens) *n_tokens = NT;
  if (n_blocks) *n_blocks = blocks;
  return 0;
}

extern "C" int model_deflate(const uint8_t *d, size_t n, int level, uint8_t **out, size_t *out_len, uint64_t *n_tokens,
                             uint64_t *n_blocks) {
  return model_deflate_wb(d, n, level, 15, out, out_len, n_tokens, n_blocks);
}
