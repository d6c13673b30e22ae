/*
 * oracle/orc.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C restatement of the pure-Dart codec hot path of brendan-duncan/archive 4.2.0
 * (the reference is Dart; there is no Dart SDK in the build container, so the reference
 * itself cannot run -- see DESIGN.md "Oracle").  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.
 *
 * Parity pin: the decoders are pinned against every fixture / known-answer vector the
 * reference's own tests hold for this path (tests/golden/, see tests/test_oracle_golden.py).
 * The ENCODERS' compressed bytes are "parity unpinned" by the reference's tests (round-trip
 * only); they are pinned here only by restating the source and cross-checked
 * against system zlib / libbz2.
 */
#ifndef ORC_H
#define ORC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: what the Dart code would have done */
#define ORC_OK 0        /* ran to completion / returned true                        */
#define ORC_FALSE 1     /* a decodeStream() that returned false (partial output kept) */
#define ORC_THROW 2     /* the Dart code would have thrown (RangeError etc.)          */
#define ORC_RUNAWAY 3   /* the Dart code would never terminate (output grows without consuming input) */

/* InputMemoryStream (lib/src/util/input_memory_stream.dart:8-134) */
typedef struct {
  const uint8_t *buf;
  int64_t len;
  int64_t pos;
  int big_endian;
} orc_ims;

/* OutputMemoryStream (lib/src/util/output_memory_stream.dart:8-137) */
typedef struct {
  uint8_t *buf;
  int64_t len;
  int64_t cap;
} orc_oms;

void orc_oms_init(orc_oms *o, int64_t size);
void orc_oms_free(orc_oms *o);
void orc_oms_write_byte(orc_oms *o, int v);
void orc_oms_write_bytes(orc_oms *o, const uint8_t *p, int64_t n);

/* util */
uint32_t orc_crc32(const uint8_t *p, size_t n, uint32_t crc);  /* crc32.dart:6-27   */
uint32_t orc_adler32(const uint8_t *p, size_t n, uint32_t adler); /* adler32.dart:29-52 */
uint32_t orc_bz2_crc(const uint8_t *p, size_t n);              /* bzip2.dart:11-14 over a buffer */

/* Inflate (inflate.dart).  Decodes from in->pos; leaves in->pos where Dart would. */
int orc_inflate(orc_ims *in, orc_oms *out);

/* Framing: _gzip_decoder_web.dart / _zlib_decoder_web.dart */
int orc_gzip_decode(orc_ims *in, orc_oms *out, int verify, int raw);
int orc_zlib_decode(orc_ims *in, orc_oms *out, int verify, int raw);

/* Flat convenience wrappers for ctypes: return status; *out is malloc'ed (free with orc_free). */
int orc_inflate_bytes(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len, size_t *consumed);
int orc_gzip_decode_bytes(const uint8_t *in, size_t n, int verify, uint8_t **out, size_t *out_len);
int orc_zlib_decode_bytes(const uint8_t *in, size_t n, int verify, int raw, uint8_t **out, size_t *out_len);

/* Deflate (deflate.dart). level 0..9, windowBits 9..15. Returns ORC_THROW for invalid params. */
int orc_deflate_bytes(const uint8_t *in, size_t n, int level, int window_bits, uint8_t **out,
                      size_t *out_len, uint32_t *crc32_of_input);
int orc_zlib_encode_bytes(const uint8_t *in, size_t n, int level, int window_bits, int raw,
                          uint8_t **out, size_t *out_len);
int orc_gzip_encode_bytes(const uint8_t *in, size_t n, int level, uint32_t mtime, uint8_t **out,
                          size_t *out_len);

/* BZip2 (bzip2_decoder.dart / bzip2_encoder.dart) */
int orc_bzip2_decode_bytes(const uint8_t *in, size_t n, int verify, uint8_t **out, size_t *out_len);
int orc_bzip2_encode_bytes(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);

void orc_free(void *p);
void orc_set_runaway_limit(int64_t n);
void orc_deflate_set_truncate_heuristic(int on); /* tests only: off == stock zlib behaviour */

#ifdef __cplusplus
}
#endif
/* ZIP container (zip.c): directory listing and member content. */
typedef struct {
  uint64_t local_header_off, data_off, comp_size, uncomp_size, hint_uncomp_size, name_off, cd_name_off;
  uint32_t name_len, cd_name_len, crc32, method, flags, mod_time, mod_date, ext_attr, version_made_by, has_data;
} orc_zip_entry;
int orc_zip_list(const uint8_t *b, size_t blen, orc_zip_entry *out, size_t cap, size_t *n_out);
int orc_zip_member(const uint8_t *b, size_t blen, const orc_zip_entry *e, int web_eos, uint8_t **out, size_t *out_len);

/* ZipEncoder container (zip_enc.c). */
typedef struct {
  const char *name;      /* UTF-8, zero terminated                                    */
  const uint8_t *content;
  size_t content_len;
  int method;            /* 0 none, 1 deflate, 2 bzip2 (CompressionType)             */
  int is_file;
  uint32_t mode;
  uint32_t dos_time, dos_date;
  const char *comment;   /* may be NULL                                               */
} orc_zip_member_in;
int orc_zip_encode(const orc_zip_member_in *m, size_t n, int level, const char *comment, uint8_t **out, size_t *out_len);

#endif
