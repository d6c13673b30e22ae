/*
 * oracle/inflate.c -- CPU ORACLE (test infrastructure only; see orc.h).
 *
 * Restates, with the same control flow and the same quirks:
 *   lib/src/codecs/zlib/inflate.dart            (Inflate)
 *   lib/src/codecs/zlib/_huffman_table.dart     (HuffmanTable)
 *   lib/src/util/output_memory_stream.dart      (OutputMemoryStream)
 *   lib/src/util/input_memory_stream.dart       (InputMemoryStream)
 *   lib/src/codecs/zlib/_gzip_decoder_web.dart  (gzip member loop + header skip)
 *   lib/src/codecs/zlib/_zlib_decoder_web.dart  (zlib stream loop + Adler verify)
 *   lib/src/util/{crc32,adler32}.dart
 * Dart `int` is 64-bit: all scalar state here is int64_t.
 */
#include "orc.h"

#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

/* ---- "throw": any Dart RangeError / exception unwinds to the public entry point ---- */
static __thread jmp_buf *orc_jmp;
static void orc_throw(void) { longjmp(*orc_jmp, 1); }
/* Holes in an incomplete Huffman set make the reference emit (len 0, sym 0) for ever (memory
 * exhaustion in Dart).  The oracle stops such a run once the output passes this bound. */
static int64_t orc_runaway_limit = (int64_t)1 << 28;
void orc_set_runaway_limit(int64_t n) { orc_runaway_limit = n; }
static void orc_runaway(void) { longjmp(*orc_jmp, 2); }

/* ------------------------------------------------------------------------------------
 * OutputMemoryStream  (output_memory_stream.dart)
 * ---------------------------------------------------------------------------------- */
#define OMS_DEFAULT 0x8000 /* :11 defaultBufferSize */

void orc_oms_init(orc_oms *o, int64_t size) {
  o->cap = size;
  o->len = 0;
  o->buf = (uint8_t *)calloc(size > 0 ? (size_t)size : 1, 1);
}
void orc_oms_free(orc_oms *o) {
  free(o->buf);
  o->buf = NULL;
}
/* _expandBuffer :127-136 -- geometric growth, fresh storage is zero (Uint8List) */
static void oms_expand(orc_oms *o, int64_t required) {
  int64_t min_len = o->cap + (required > 0 ? required : 1);
  int64_t nl = o->cap == 0 ? OMS_DEFAULT : o->cap * 2;
  if (nl < min_len) nl = min_len;
  uint8_t *nb = (uint8_t *)calloc((size_t)nl, 1);
  memcpy(nb, o->buf, (size_t)o->cap);
  free(o->buf);
  o->buf = nb;
  o->cap = nl;
}
/* writeByte :41-46 */
void orc_oms_write_byte(orc_oms *o, int v) {
  if (o->len >= orc_runaway_limit) orc_runaway();
  if (o->len == o->cap) oms_expand(o, 0);
  o->buf[o->len++] = (uint8_t)v;
}
/* writeBytes :49-58 */
void orc_oms_write_bytes(orc_oms *o, const uint8_t *p, int64_t n) {
  while (o->len + n > o->cap) oms_expand(o, (o->len + n) - o->cap);
  memcpy(o->buf + o->len, p, (size_t)n);
  o->len += n;
}
/* writeBackReference :79-98 */
static void oms_write_backref(orc_oms *o, int64_t distance, int64_t count) {
  if (o->len >= orc_runaway_limit) orc_runaway();
  while (o->len + count > o->cap) oms_expand(o, (o->len + count) - o->cap);
  int64_t src = o->len - distance;
  if (distance >= count) {
    /* setRange(length, length+count, _buffer, src): negative skipCount or a source range
     * running past the buffer end throws; count <= 0 with a valid src is a no-op / throws
     * on end < start. */
    if (count < 0) orc_throw();
    if (count == 0) return;
    if (src < 0 || src + count > o->cap) orc_throw();
    memmove(o->buf + o->len, o->buf + src, (size_t)count);
  } else {
    int64_t s = src, d = o->len, end = o->len + count;
    while (d < end) {
      if (s < 0) orc_throw(); /* _buffer[s] with s < 0 -> RangeError */
      o->buf[d++] = o->buf[s++];
    }
  }
  o->len += count;
}

/* ------------------------------------------------------------------------------------
 * InputMemoryStream helpers (input_memory_stream.dart, input_stream.dart)
 * ---------------------------------------------------------------------------------- */
static int ims_eos(const orc_ims *s) { return s->pos >= s->len; }          /* :60 */
static int64_t ims_length(const orc_ims *s) { return s->len - s->pos; }    /* :56 */
static int ims_read_byte(orc_ims *s) {                                     /* :121-124 */
  if (s->pos < 0 || s->pos >= s->len) orc_throw();
  return s->buf[s->pos++];
}
static void ims_rewind1(orc_ims *s) {                                      /* :88-91 */
  s->pos -= 1;
  if (s->pos < 0) s->pos = 0;
  if (s->pos > s->len) s->pos = s->len;
}
static int64_t ims_read_u16(orc_ims *s) {                                  /* input_stream.dart:67-74 */
  int64_t b1 = ims_read_byte(s), b2 = ims_read_byte(s);
  return s->big_endian ? ((b1 << 8) | b2) : ((b2 << 8) | b1);
}
static int64_t ims_read_u32(orc_ims *s) {                                  /* input_stream.dart:88-97 */
  int64_t b1 = ims_read_byte(s), b2 = ims_read_byte(s), b3 = ims_read_byte(s), b4 = ims_read_byte(s);
  return s->big_endian ? ((b1 << 24) | (b2 << 16) | (b3 << 8) | b4)
                       : ((b4 << 24) | (b3 << 16) | (b2 << 8) | b1);
}
/* readBytes(count): subset clamps to the remaining bytes (input_stream.dart:132-136,
 * input_memory_stream.dart:17-22); returns the clamped length and advances. */
static int64_t ims_skip_bytes(orc_ims *s, int64_t count) {
  int64_t rem = s->len - s->pos;
  if (rem < 0) rem = 0;
  if (count > rem) count = rem;
  if (count < 0) orc_throw();
  s->pos += count;
  return count;
}
/* readString() null-terminated (input_stream.dart:152-164) */
static void ims_skip_cstring(orc_ims *s) {
  if (ims_eos(s)) return;
  while (!ims_eos(s)) {
    if (ims_read_byte(s) == 0) return;
  }
}

/* ------------------------------------------------------------------------------------
 * HuffmanTable  (_huffman_table.dart:4-47): flat 1<<maxLen table, entry = len<<16 | sym,
 * no validation of over-subscribed / incomplete sets (later writes win, holes stay 0).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  uint32_t *table;
  int64_t max_len, min_len, size;
} huff;

static void huff_build(huff *h, const uint8_t *lengths, int n) {
  h->max_len = 0;
  h->min_len = 0x7fffffff;
  for (int i = 0; i < n; ++i) {
    if (lengths[i] > h->max_len) h->max_len = lengths[i];
    if (lengths[i] < h->min_len) h->min_len = lengths[i];
  }
  h->size = (int64_t)1 << h->max_len;
  h->table = (uint32_t *)calloc((size_t)h->size, sizeof(uint32_t));
  int64_t code = 0, skip = 2;
  for (int64_t bl = 1; bl <= h->max_len;) {
    for (int i = 0; i < n; ++i) {
      if (lengths[i] == bl) {
        int64_t reversed = 0, rt = code;
        for (int64_t j = 0; j < bl; ++j) {
          reversed = (reversed << 1) | (rt & 1);
          rt >>= 1;
        }
        for (int64_t j = reversed; j < h->size; j += skip) h->table[j] = (uint32_t)((bl << 16) | i);
        ++code;
      }
    }
    ++bl;
    code <<= 1;
    skip <<= 1;
  }
}
static void huff_free(huff *h) {
  free(h->table);
  h->table = NULL;
}

/* ------------------------------------------------------------------------------------
 * Inflate  (inflate.dart)
 * ---------------------------------------------------------------------------------- */
typedef struct {
  orc_ims *in;
  orc_oms *out;
  int64_t bit_buffer, bit_len; /* _bitBuffer/_bitBufferLen :404-405 */
} inflate_t;

static const uint8_t k_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; /* :738-758 */
static const int k_len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                   31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258}; /* :761-791 */
static const int k_len_extra[31] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                    3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0}; /* :794-826 */
static const int k_dist_base[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                    193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}; /* :829-860 */
static const int k_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                     6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13}; /* :863-894 */

/* _readBits :159-184 */
static int64_t read_bits(inflate_t *z, int64_t length) {
  if (length == 0) return 0;
  while (z->bit_len < length) {
    if (ims_eos(z->in)) return -1;
    int64_t octet = ims_read_byte(z->in);
    z->bit_buffer |= octet << z->bit_len;
    z->bit_len += 8;
  }
  int64_t octet = z->bit_buffer & (((int64_t)1 << length) - 1);
  z->bit_buffer >>= length;
  z->bit_len -= length;
  return octet;
}

/* _readCodeByTable :187-211 -- note: demands maxCodeLength bits be available (quirk Q1) */
static int64_t read_code(inflate_t *z, const huff *t) {
  while (z->bit_len < t->max_len) {
    if (ims_eos(z->in)) return -1;
    int64_t octet = ims_read_byte(z->in);
    z->bit_buffer |= octet << z->bit_len;
    z->bit_len += 8;
  }
  uint32_t cwl = t->table[z->bit_buffer & (((int64_t)1 << t->max_len) - 1)];
  int64_t cl = cwl >> 16;
  z->bit_buffer >>= cl;
  z->bit_len -= cl;
  return cwl & 0xffff;
}

/* _parseUncompressedBlock :213-235 */
static int parse_stored(inflate_t *z) {
  z->bit_buffer = 0;
  z->bit_len = 0;
  int64_t len = read_bits(z, 16);
  int64_t nlen = read_bits(z, 16) ^ 0xffff;
  if (len != 0 && len != nlen) return -1;
  if (len > ims_length(z->in)) return -1;
  /* readBytes(len) + writeStream(bytes) (output_memory_stream.dart:61-76) */
  int64_t p0 = z->in->pos;
  int64_t n = ims_skip_bytes(z->in, len);
  orc_oms_write_bytes(z->out, z->in->buf + p0, n);
  return 0;
}

/* _decodeHuffman :300-343 */
static int decode_huffman(inflate_t *z, const huff *litlen, const huff *dist) {
  for (;;) {
    int64_t code = read_code(z, litlen);
    if (code < 0 || code > 285) return -1;
    if (code == 256) break;
    if (code < 256) {
      orc_oms_write_byte(z->out, (int)(code & 0xff));
      continue;
    }
    int64_t ti = code - 257;
    int64_t code_length = k_len_base[ti] + read_bits(z, k_len_extra[ti]);
    int64_t dist_code = read_code(z, dist);
    if (dist_code < 0 || dist_code > 29) return -1;
    int64_t distance = k_dist_base[dist_code] + read_bits(z, k_dist_extra[dist_code]);
    oms_write_backref(z->out, distance, code_length);
  }
  while (z->bit_len >= 8) { /* :337-340 un-read whole bytes */
    z->bit_len -= 8;
    ims_rewind1(z->in);
  }
  return 0;
}

/* _decode :345-401 (code-length alphabet 16/17/18 run codes) */
static int decode_lengths(inflate_t *z, int num, const huff *table, uint8_t *cl) {
  int64_t prev = 0;
  int i = 0;
  while (i < num) {
    int64_t code = read_code(z, table);
    if (code == -1) return -1;
    int64_t repeat;
    switch (code) {
      case 16:
        repeat = read_bits(z, 2);
        if (repeat == -1) return -1;
        repeat += 3;
        while (repeat-- > 0) {
          if (i >= num) orc_throw(); /* codeLengths[i++] RangeError */
          cl[i++] = (uint8_t)prev;
        }
        break;
      case 17:
        repeat = read_bits(z, 3);
        if (repeat == -1) return -1;
        repeat += 3;
        while (repeat-- > 0) {
          if (i >= num) orc_throw();
          cl[i++] = 0;
        }
        prev = 0;
        break;
      case 18:
        repeat = read_bits(z, 7);
        if (repeat == -1) return -1;
        repeat += 11;
        while (repeat-- > 0) {
          if (i >= num) orc_throw();
          cl[i++] = 0;
        }
        prev = 0;
        break;
      default:
        if (code < 0 || code > 15) return -1;
        cl[i++] = (uint8_t)code;
        prev = code;
        break;
    }
  }
  return 0;
}

/* _parseDynamicHuffmanBlock :239-298 */
static int parse_dynamic(inflate_t *z) {
  int64_t hlit = read_bits(z, 5);
  if (hlit == -1) return -1;
  hlit += 257;
  if (hlit > 288) return -1;
  int64_t hdist = read_bits(z, 5);
  if (hdist == -1) return -1;
  hdist += 1;
  if (hdist > 32) return -1;
  int64_t hclen = read_bits(z, 4);
  if (hclen == -1) return -1;
  hclen += 4;
  if (hclen > 19) return -1;

  uint8_t cl19[19];
  memset(cl19, 0, sizeof cl19);
  for (int i = 0; i < hclen; ++i) {
    int64_t len = read_bits(z, 3);
    if (len == -1) return -1;
    cl19[k_order[i]] = (uint8_t)len;
  }
  huff clt;
  huff_build(&clt, cl19, 19);

  uint8_t lens[288 + 32];
  memset(lens, 0, sizeof lens);
  int r = decode_lengths(z, (int)(hlit + hdist), &clt, lens);
  huff_free(&clt);
  if (r == -1) return -1;

  huff lt, dt;
  huff_build(&lt, lens, (int)hlit);
  huff_build(&dt, lens + hlit, (int)hdist);
  r = decode_huffman(z, &lt, &dt);
  huff_free(&lt);
  huff_free(&dt);
  return r;
}

/* _parseFixedHuffmanBlock :236-237 with tables :408-735 */
static int parse_fixed(inflate_t *z) {
  static huff flt, fdt;
  static int ready;
  if (!ready) {
    uint8_t l[288], d[30];
    for (int i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    for (int i = 0; i < 30; ++i) d[i] = 5;
    huff_build(&flt, l, 288);
    huff_build(&fdt, d, 30);
    ready = 1;
  }
  return decode_huffman(z, &flt, &fdt);
}

/* _parseBlock :120-156; returns 1 = more blocks, 0 = done/stop */
static int parse_block(inflate_t *z) {
  if (ims_eos(z->in)) return 0;
  int64_t hdr = read_bits(z, 3);
  int final_block = (hdr & 1) != 0;
  int64_t type = hdr >> 1; /* hdr == -1 (only at EOS, unreachable here) -> type -1 -> default */
  switch (type) {
    case 0:
      if (parse_stored(z) == -1) return 0;
      break;
    case 1:
      if (parse_fixed(z) == -1) return 0;
      break;
    case 2:
      if (parse_dynamic(z) == -1) return 0;
      break;
    default:
      return 0;
  }
  return !final_block;
}

/* _inflate :104-116 (ctor path of Inflate(bytes) / Inflate.stream(input, output:)) */
static void inflate_run(orc_ims *in, orc_oms *out) {
  inflate_t z = {in, out, 0, 0};
  while (!ims_eos(in)) {
    if (!parse_block(&z)) return;
  }
}

int orc_inflate(orc_ims *in, orc_oms *out) {
  jmp_buf jb, *saved = orc_jmp;
  orc_jmp = &jb;
  int st = ORC_OK;
  int j = setjmp(jb);
  if (j == 0) inflate_run(in, out);
  else st = j == 2 ? ORC_RUNAWAY : ORC_THROW;
  orc_jmp = saved;
  return st;
}

/* ------------------------------------------------------------------------------------
 * crc32.dart:6-27 (table CRC-32, reflected 0xEDB88320), adler32.dart:29-52 (NMAX = 3800)
 * ---------------------------------------------------------------------------------- */
static uint32_t crc_tab[256];
static void crc_init(void) {
  if (crc_tab[1]) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_tab[i] = c;
  }
}
uint32_t orc_crc32(const uint8_t *p, size_t n, uint32_t crc) {
  crc_init();
  crc ^= 0xffffffffu;
  for (size_t i = 0; i < n; ++i) crc = crc_tab[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
  return crc ^ 0xffffffffu;
}
uint32_t orc_adler32(const uint8_t *p, size_t len, uint32_t adler) {
  const uint32_t base = 65521;
  uint32_t s1 = adler & 0xffff, s2 = adler >> 16;
  size_t i = 0;
  while (len > 0) {
    size_t n = 3800;
    if (n > len) n = len;
    len -= n;
    while (n-- > 0) {
      s1 += p[i++];
      s2 += s1;
    }
    s1 %= base;
    s2 %= base;
  }
  return (s2 << 16) | s1;
}

/* ------------------------------------------------------------------------------------
 * _zlib_decoder_web.dart:31-107
 * ---------------------------------------------------------------------------------- */
static int zlib_decode_run(orc_ims *in, orc_oms *out, int verify, int raw) {
  orc_oms buffer;
  int have = 0;
  while (!ims_eos(in)) {
    if (!raw) {
      int64_t cmf = ims_read_byte(in);
      int64_t flg = ims_read_byte(in);
      int64_t method = cmf & 8; /* :57 (sic: & 8, quirk Q2) */
      if (method != 8) {
        if (have) orc_oms_free(&buffer);
        return ORC_FALSE;
      }
      int64_t fdict = (flg & 32) >> 5;
      if (((cmf * 256) + flg) % 31 != 0) {
        if (have) orc_oms_free(&buffer);
        return ORC_FALSE;
      }
      if (fdict != 0) {
        ims_read_u32(in);
        if (have) orc_oms_free(&buffer);
        return ORC_FALSE;
      }
    }
    if (have) { /* :82-84 previous stream's bytes are committed only now */
      orc_oms_write_bytes(out, buffer.buf, buffer.len);
      orc_oms_free(&buffer);
      have = 0;
    }
    orc_oms_init(&buffer, OMS_DEFAULT);
    have = 1;
    inflate_run(in, &buffer); /* :87 Inflate.stream(input).getBytes() */
    if (!raw) {
      int64_t adler = ims_read_u32(in); /* byte order of the stream: BE via decodeBytes :25 */
      if (verify) {
        uint32_t a = orc_adler32(buffer.buf, (size_t)buffer.len, 1);
        if ((uint32_t)adler != a || (adler >> 32) != 0) {
          orc_oms_free(&buffer);
          return ORC_FALSE;
        }
      }
    }
  }
  if (have) {
    orc_oms_write_bytes(out, buffer.buf, buffer.len);
    orc_oms_free(&buffer);
  }
  return ORC_OK;
}

int orc_zlib_decode(orc_ims *in, orc_oms *out, int verify, int raw) {
  jmp_buf jb, *saved = orc_jmp;
  orc_jmp = &jb;
  int st;
  int j = setjmp(jb);
  if (j == 0) st = zlib_decode_run(in, out, verify, raw);
  else st = j == 2 ? ORC_RUNAWAY : ORC_THROW; /* (leaks the per-stream scratch on throw; oracle only) */
  orc_jmp = saved;
  return st;
}

/* ------------------------------------------------------------------------------------
 * _gzip_decoder_web.dart:27-138
 * ---------------------------------------------------------------------------------- */
static int gzip_read_header(orc_ims *in) { /* _readHeader :60-138 */
  int64_t sig = ims_read_u16(in);
  if (sig != 0x8b1f) return 0;
  int64_t cm = ims_read_byte(in);
  if (cm != 8) return 0;
  int64_t flags = ims_read_byte(in);
  ims_read_u32(in); /* mtime */
  ims_read_byte(in); /* xfl */
  ims_read_byte(in); /* os */
  if (flags & 0x04) {
    int64_t t = ims_read_u16(in);
    ims_skip_bytes(in, t);
  }
  if (flags & 0x08) ims_skip_cstring(in);
  if (flags & 0x10) ims_skip_cstring(in);
  if (flags & 0x02) ims_read_u16(in);
  return 1;
}

static int gzip_decode_run(orc_ims *in, orc_oms *out, int verify, int raw) {
  while (!ims_eos(in)) {
    int64_t start = in->pos;
    if (!gzip_read_header(in)) {
      in->pos = start; /* :34-36 fall back to zlib on the same (little-endian) stream */
      return zlib_decode_run(in, out, verify, raw);
    }
    inflate_run(in, out); /* :38 shared output stream */
    ims_read_u32(in);     /* crc, discarded (quirk Q3) */
    ims_read_u32(in);     /* isize, discarded */
  }
  return ORC_OK;
}

int orc_gzip_decode(orc_ims *in, orc_oms *out, int verify, int raw) {
  jmp_buf jb, *saved = orc_jmp;
  orc_jmp = &jb;
  int st;
  int j = setjmp(jb);
  if (j == 0) st = gzip_decode_run(in, out, verify, raw);
  else st = j == 2 ? ORC_RUNAWAY : ORC_THROW;
  orc_jmp = saved;
  return st;
}

/* ------------------------------------------------------------------------------------
 * flat wrappers
 * ---------------------------------------------------------------------------------- */
void orc_free(void *p) { free(p); }

int orc_inflate_bytes(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len, size_t *consumed) {
  orc_ims s = {in, (int64_t)n, 0, 0};
  orc_oms o;
  orc_oms_init(&o, OMS_DEFAULT);
  int st = orc_inflate(&s, &o);
  *out = o.buf;
  *out_len = (size_t)o.len;
  if (consumed) *consumed = (size_t)s.pos;
  return st;
}
int orc_gzip_decode_bytes(const uint8_t *in, size_t n, int verify, uint8_t **out, size_t *out_len) {
  orc_ims s = {in, (int64_t)n, 0, 0}; /* _gzip_decoder_web.dart:19-24: default little-endian */
  orc_oms o;
  orc_oms_init(&o, OMS_DEFAULT);
  int st = orc_gzip_decode(&s, &o, verify, 0);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}
/* GZipDecoderWeb().decodeBytes(bytes, verify:, raw:) -- `raw` only matters for input without a gzip header, which goes to
 * the zlib decoder with it (_gzip_decoder_web.dart:31-37) */
int orc_gzip_decode_bytes_raw(const uint8_t *in, size_t n, int verify, int raw, uint8_t **out, size_t *out_len) {
  orc_ims s = {in, (int64_t)n, 0, 0};
  orc_oms o;
  orc_oms_init(&o, OMS_DEFAULT);
  int st = orc_gzip_decode(&s, &o, verify, raw);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}
int orc_zlib_decode_bytes(const uint8_t *in, size_t n, int verify, int raw, uint8_t **out, size_t *out_len) {
  orc_ims s = {in, (int64_t)n, 0, 1}; /* _zlib_decoder_web.dart:21-28: big-endian */
  orc_oms o;
  orc_oms_init(&o, OMS_DEFAULT);
  int st = orc_zlib_decode(&s, &o, verify, raw);
  *out = o.buf;
  *out_len = (size_t)o.len;
  return st;
}
