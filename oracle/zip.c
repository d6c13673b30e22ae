/*
 * oracle/zip.c -- CPU ORACLE (test infrastructure only; see orc.h).
 *
 * Restates the ZIP reader of the reference as far as the B200 path replaces it:
 *   lib/src/codecs/zip/zip_directory.dart:25-183  ZipDirectory.read / _readZip64Data / _findSignature
 *   lib/src/codecs/zip/zip_file_header.dart:28-111 ZipFileHeader.read (incl. the zip64 extra field)
 *   lib/src/codecs/zip/zip_file.dart:73-149        ZipFile.read (local header, data descriptor)
 *   lib/src/codecs/zip/zip_file.dart:164-249       getStream / decompress: ZLibDecoder().decodeBytes(raw: true) on the
 *                                                  member's bytes, BZip2Decoder for method 12, copy otherwise
 * written as the Dart reads it: a cursor over the bytes whose reads throw past the end (input_memory_stream.dart:121-124).
 * Encrypted members are outside the scope (reported, not decoded).
 * Pinned by the reference's fixtures (tests/golden/zip/, expectations from CPython's zipfile: tests/golden/make_golden.py).
 */
#include <stdlib.h>
#include <string.h>

#include "orc.h"

typedef struct {
  const uint8_t *b;
  int64_t len, pos;
  int threw;
} cur;

static uint32_t rd(cur *c, int n) { /* little-endian readUint16 / readUint32 */
  uint32_t v = 0;
  for (int i = 0; i < n; i++) {
    if (c->pos >= c->len || c->pos < 0) {
      c->threw = 1;
      return 0;
    }
    v |= (uint32_t)c->b[c->pos++] << (8 * i);
  }
  return v;
}
static uint64_t rd64(cur *c) {
  uint64_t lo = rd(c, 4);
  uint64_t hi = rd(c, 4);
  return lo | (hi << 32);
}
static void skip(cur *c, int64_t n) { /* readBytes / readString of n bytes */
  if (n > 0 && c->pos + n > c->len) c->threw = 1;
  c->pos += n;
}

/* _findSignature :139-182 */
static int64_t find_signature(const uint8_t *b, int64_t len) {
  if (len <= 4) return -1;
  int64_t length = len - 4;
  int64_t chunk_size = length < 1024 ? length : 1024;
  int64_t start_pos = length - chunk_size;
  while (start_pos >= 0) {
    for (int64_t chunk_pos = chunk_size - 4; chunk_pos >= 0; --chunk_pos) {
      const uint8_t *p = b + start_pos + chunk_pos;
      uint32_t sig = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
      if (sig == 0x06054b50u) return start_pos + chunk_pos;
    }
    if (start_pos > 0 && start_pos < chunk_size) start_pos = 0;
    else start_pos -= chunk_size;
  }
  return -1;
}

int orc_zip_list(const uint8_t *b, size_t blen, orc_zip_entry *out, size_t cap, size_t *n_out) {
  *n_out = 0;
  int64_t len = (int64_t)blen;
  int64_t file_position = find_signature(b, len);
  if (file_position < 0) return ORC_OK;
  cur in = {b, len, file_position, 0};
  uint32_t sig = rd(&in, 4);
  if (in.threw) return ORC_THROW;
  if (sig != 0x06054b50u) return ORC_OK;
  rd(&in, 2); rd(&in, 2); rd(&in, 2); rd(&in, 2);
  uint64_t dir_size = rd(&in, 4), dir_offset = rd(&in, 4);
  uint32_t clen = rd(&in, 2);
  if (clen > 0) skip(&in, clen);
  if (in.threw) return ORC_THROW;
  /* _readZip64Data */
  int64_t loc_pos = file_position - 20;
  if (loc_pos >= 0) {
    cur z = {b, len, loc_pos, 0};
    if (rd(&z, 4) == 0x07064b50u) {
      rd(&z, 4);
      uint64_t z64off = rd64(&z);
      rd(&z, 4);
      cur e = {b, len, (int64_t)z64off, 0};
      uint32_t s2 = rd(&e, 4);
      if (e.threw) return ORC_THROW;
      if (s2 == 0x06064b50u) {
        rd64(&e); rd(&e, 2); rd(&e, 2); rd(&e, 4); rd(&e, 4); rd64(&e); rd64(&e);
        dir_size = rd64(&e);
        dir_offset = rd64(&e);
        if (e.threw) return ORC_THROW;
      }
    }
  }
  /* the central directory: a sub-stream [dir_offset, dir_offset + dir_size) */
  /* subset(position, length) (input_memory_stream.dart:15-27,111-119): the length is cut to what the archive holds; an
   * offset beyond it, or a negative offset / size, makes Uint8List.view throw; reads are bounded by the sub-stream */
  if ((int64_t)dir_offset < 0 || (int64_t)dir_offset > len || (int64_t)dir_size < 0) return ORC_THROW;
  int64_t d_end = (int64_t)dir_size > len - (int64_t)dir_offset ? len : (int64_t)(dir_offset + dir_size);
  cur d = {b, d_end, (int64_t)dir_offset, 0};
  size_t n = 0;
  while (d.pos < d_end) {
    uint32_t fsig = rd(&d, 4);
    if (d.threw) return ORC_THROW;
    if (fsig != 0x02014b50u) break;
    orc_zip_entry e;
    memset(&e, 0, sizeof e);
    e.version_made_by = rd(&d, 2);
    rd(&d, 2); rd(&d, 2); rd(&d, 2); rd(&d, 2); rd(&d, 2); rd(&d, 4);
    uint64_t comp = rd(&d, 4), uncomp = rd(&d, 4);
    uint32_t fname_len = rd(&d, 2), extra_len = rd(&d, 2), comment_len = rd(&d, 2);
    uint32_t disk = rd(&d, 2);
    rd(&d, 2);
    e.ext_attr = rd(&d, 4);
    uint64_t lho = rd(&d, 4);
    e.cd_name_off = (uint64_t)d.pos;
    e.cd_name_len = fname_len;
    if (fname_len > 0) skip(&d, fname_len);
    if (extra_len > 0) {
      cur x = {b, d.pos + extra_len, d.pos, 0};
      skip(&d, extra_len);
      if (d.threw) return ORC_THROW;
      if (extra_len >= 4) {
        while (x.len - x.pos >= 4) {
          uint32_t id = rd(&x, 2);
          uint32_t size = rd(&x, 2);
          cur eb = {b, x.pos + size, x.pos, 0};
          skip(&x, size);
          if (x.threw) return ORC_THROW;
          if (id == 1) {
            if (size >= 8 && uncomp == 0xffffffffu) { uncomp = rd64(&eb); size -= 8; }
            if (size >= 8 && comp == 0xffffffffu) { comp = rd64(&eb); size -= 8; }
            if (size >= 8 && lho == 0xffffffffu) { lho = rd64(&eb); size -= 8; }
            if (size >= 4 && disk == 0xffffu) { disk = rd(&eb, 4); size -= 4; }
          }
        }
      }
    }
    if (comment_len > 0) skip(&d, comment_len);
    if (d.threw) return ORC_THROW;
    /* ZipFile.read */
    e.local_header_off = lho;
    e.comp_size = comp;
    e.uncomp_size = uncomp;
    e.hint_uncomp_size = uncomp;
    cur f = {b, len, (int64_t)lho, 0};
    uint32_t lsig = rd(&f, 4);
    if (f.threw) return ORC_THROW;
    if (lsig == 0x04034b50u) {
      rd(&f, 2);
      e.flags = rd(&f, 2);
      e.method = rd(&f, 2);
      e.mod_time = rd(&f, 2);
      e.mod_date = rd(&f, 2);
      e.crc32 = rd(&f, 4);
      rd(&f, 4); rd(&f, 4);
      uint32_t fn_len = rd(&f, 2), ex_len = rd(&f, 2);
      e.name_off = (uint64_t)f.pos;
      e.name_len = fn_len;
      skip(&f, fn_len);
      skip(&f, ex_len);
      if (f.threw) return ORC_THROW;
      e.data_off = (uint64_t)f.pos;
      e.has_data = 1;
      int64_t avail = len - f.pos;
      if ((int64_t)comp < 0) return ORC_THROW; /* readBytes(negative count): Uint8List.view throws */
      if ((int64_t)comp > avail) e.comp_size = (uint64_t)avail; /* readBytes gives what is left */
      f.pos += (int64_t)e.comp_size;
      if (e.flags & 0x08) {
        uint32_t sig_or_crc = rd(&f, 4);
        if (sig_or_crc == 0x08074b50u) e.crc32 = rd(&f, 4);
        else e.crc32 = sig_or_crc;
        rd(&f, 4);
        e.uncomp_size = rd(&f, 4);
        if (f.threw) return ORC_THROW;
      }
    }
    if (n < cap) out[n] = e;
    n++;
  }
  *n_out = n;
  return ORC_OK;
}

/* content of one listed member as the reference produces it.  web_eos != 0: pure-Dart Inflate on exactly the member's
 * bytes (ZLibDecoderWeb(raw: true)); 0: as dart:io's zlib decodes it (all symbols) -- restated by letting Inflate see up to
 * 8 bytes that follow the member, which only satisfies its look-ahead (inflate.dart:192-195). */
int orc_zip_member(const uint8_t *b, size_t blen, const orc_zip_entry *e, int web_eos, uint8_t **out, size_t *out_len) {
  *out = NULL;
  *out_len = 0;
  if (!e->has_data) return ORC_OK;
  if (e->flags & 1) return ORC_FALSE;
  if (e->method == 8) {
    uint64_t pad = 0;
    if (!web_eos) {
      pad = blen - (e->data_off + e->comp_size);
      if (pad > 8) pad = 8;
    }
    size_t consumed;
    return orc_inflate_bytes(b + e->data_off, (size_t)(e->comp_size + pad), out, out_len, &consumed);
  }
  if (e->method == 12) return orc_bzip2_decode_bytes(b + e->data_off, (size_t)e->comp_size, 0, out, out_len);
  *out = (uint8_t *)malloc(e->comp_size ? e->comp_size : 1);
  memcpy(*out, b + e->data_off, e->comp_size);
  *out_len = e->comp_size;
  return ORC_OK;
}
