/* zipc_oracle.h -- CPU restatement of dbuenzli/zipc's hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a plain-C restatement of the algorithms in
 *   /root/reference/src/zipc_deflate.ml   (checksums, inflate, deflate, zlib framing)
 *   /root/reference/src/zipc.ml           (File / Member rules, archive encode + decode, DOS time)
 * Every function cites the reference file:line it follows.
 *
 * It is NOT product code.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load it.  The product (libzipc_b200.so) never links,
 * loads or calls anything in this directory.
 *
 * Pinning: the reference is OCaml and no OCaml toolchain exists in this image, so the reference
 * itself cannot be run here (oracle/_ref is "unbuildable": see DESIGN.md).  The oracle is pinned
 * against every vector the reference's own tests hold for this path (test/test.ml:14-129 and the
 * embedded zip-docs.zip fixture), against the survey-time independent emulation's known answers
 *  and against system zlib; see tests/test_oracle_*.py.
 */
#ifndef ZIPC_ORACLE_H
#define ZIPC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes; zo_strerror gives the reference's exact English message. */
enum {
  ZO_OK = 0,
  ZO_ERR_CORRUPTED = 1,       /* "Corrupted data stream"                zipc_deflate.ml:233 */
  ZO_ERR_SIZE_EXCEEDED = 2,   /* "Expected decompression size exceeded" zipc_deflate.ml:29  */
  ZO_ERR_ZLIB_METHOD = 3,     /* "Unknown compression method (%d)"      zipc_deflate.ml:728 */
  ZO_ERR_ZLIB_WINDOW = 4,     /* "Window size too large"                zipc_deflate.ml:729 */
  ZO_ERR_ZLIB_DICT = 5,       /* "Preset dictionary unsupported"        zipc_deflate.ml:730 */
  ZO_ERR_CHECKSUM = 6,        /* "Checksum mismatch, expected %lx found %lx)"  :103-104     */
  ZO_ERR_NOMEM = 7,
  /* zipc.ml archive level */
  ZO_ERR_ZIP_ZIP64 = 20,      /* zipc.ml:290 */
  ZO_ERR_ZIP_MULTIPART = 21,  /* zipc.ml:291 */
  ZO_ERR_ZIP_EOCD = 22,       /* zipc.ml:292 */
  ZO_ERR_ZIP_NO_EOCD = 23,    /* zipc.ml:293-294 */
  ZO_ERR_ZIP_SHORT = 24,      /* zipc.ml:296 */
  ZO_ERR_ZIP_TRUNC_CD = 25,   /* zipc.ml:297 */
  ZO_ERR_ZIP_CDFH = 26,       /* zipc.ml:298 */
  ZO_ERR_ZIP_LFH = 27,        /* zipc.ml:299 */
  ZO_ERR_ZIP_COUNT = 28,      /* zipc.ml:231-232 */
  ZO_ERR_ZIP_PATH_LEN = 29,   /* zipc.ml:234-235 */
  ZO_ERR_ZIP_SIZE = 30,       /* zipc.ml:130-133 */
  ZO_ERR_ZIP_ENCRYPTED = 31,  /* zipc.ml:129 */
  ZO_ERR_ZIP_FORMAT = 32,     /* zipc.ml:128 */
  ZO_ERR_ZIP_CD_OFFSET = 33,  /* zipc.ml:550 */
  ZO_ERR_ZIP_CD_SIZE = 34     /* zipc.ml:551 */
};
const char *zo_strerror(int status);

/* crc_op  zipc_deflate.ml:210 */
enum { ZO_CRC_NOP = 0, ZO_CRC_ADLER32 = 1, ZO_CRC_CRC32 = 2 };
/* level   zipc_deflate.ml:752 */
enum { ZO_LEVEL_NONE = 0, ZO_LEVEL_FAST = 1, ZO_LEVEL_DEFAULT = 2, ZO_LEVEL_BEST = 3 };

/* Switches for the two reference behaviours that differ from the RFCs.
 * Both default to 1 = "as written in the reference". */
void zo_set_adler_signed_rem(int on);     /* zipc_deflate.ml:95,196 : Int32.rem is signed       */
void zo_set_keep_codelen_freqs(int on);   /* zipc_deflate.ml:849-854: codelen_sym_freqs not reset */

/* ---- Checksums ------------------------------------------------------------------------- */
/* Crc_32.string_update / init / finish  zipc_deflate.ml:135-156 */
uint32_t zo_crc32_init(void);
uint32_t zo_crc32_update(uint32_t c, const uint8_t *s, size_t len);
uint32_t zo_crc32_finish(uint32_t c);
uint32_t zo_crc32(const uint8_t *s, size_t len);            /* Crc_32.string  :161-163 */
/* Adler_32.string_update  zipc_deflate.ml:175-198 (int32 wrap, signed rem unless switched off) */
uint32_t zo_adler32_update(uint32_t a, const uint8_t *s, size_t len);
uint32_t zo_adler32(const uint8_t *s, size_t len);          /* Adler_32.string :203-205 */

/* ---- Inflate  zipc_deflate.ml:532-718 ---------------------------------------------------- */
/* decompressed_size < 0 means "not given" (growable output).  *out is malloc'ed (free with
 * zo_free), also on size 0.  On error *out is NULL.  *consumed (optional) receives the number
 * of source bytes the bit reader has pulled when the final block ended. */
int zo_inflate(const uint8_t *src, size_t len, int64_t decompressed_size, int crc_op,
               uint8_t **out, size_t *out_len, uint32_t *crc);
/* zlib_decompress  zipc_deflate.ml:720-740.  expect/found are filled on ZO_ERR_CHECKSUM. */
int zo_zlib_decompress(const uint8_t *src, size_t len, int64_t decompressed_size,
                       uint8_t **out, size_t *out_len, uint32_t *adler,
                       uint32_t *expect, uint32_t *found, int *method);

/* ---- Deflate  zipc_deflate.ml:742-1277 --------------------------------------------------- */
/* Per-call statistics a test can look at (block kinds chosen, etc.). */
typedef struct {
  uint32_t blocks_stored, blocks_fixed, blocks_dynamic;
  uint64_t literals, matches, match_bytes;
} zo_deflate_stats;
int zo_deflate(int level, const uint8_t *src, size_t len, int crc_op,
               uint8_t **out, size_t *out_len, uint32_t *crc, zo_deflate_stats *stats);
/* zlib_compress  zipc_deflate.ml:1262-1277 */
int zo_zlib_compress(int level, const uint8_t *src, size_t len,
                     uint8_t **out, size_t *out_len, uint32_t *adler);

void zo_free(void *p);

/* ---- ZIP archive  zipc.ml ---------------------------------------------------------------- */
/* Ptime  zipc.ml:64-125 */
void zo_ptime_to_date_time(int64_t ptime_s, int *y, int *mo, int *d, int *hh, int *mm, int *ss);
int64_t zo_ptime_of_dos(int dos_date, int dos_time);
void zo_ptime_to_dos(int64_t ptime_s, int *dos_date, int *dos_time);

/* A member as the archive encoder / decoder sees it (Member.t + File.t, zipc.ml:145-154,238-242). */
typedef struct {
  const char *path;         /* not NUL terminated; path_len bytes */
  uint32_t path_len;
  int is_dir;
  int mode;                 /* Fpath.mode */
  int64_t mtime;            /* POSIX seconds */
  /* File.t fields; ignored for directories */
  int version_made_by, version_needed, gp_flags;
  int compression;          /* method integer: 0 stored, 8 deflate ... zipc.ml:25-31 */
  const uint8_t *compressed_bytes;   /* whole buffer */
  uint64_t start;           /* offset of payload in compressed_bytes */
  uint64_t compressed_size;
  uint64_t decompressed_size;
  uint32_t crc32;
} zo_member;

/* Zipc.encoding_size zipc.ml:447-455 and Zipc.to_binary_string zipc.ml:570-588.
 * Members must be given in any order; the encoder orders them as the reference does
 * (member named `first` then byte-wise increasing path).  `first` may be NULL => "mimetype". */
uint64_t zo_zip_encoding_size(const zo_member *ms, size_t n);
int zo_zip_encode(const zo_member *ms, size_t n, const char *first,
                  uint8_t **out, size_t *out_len);
/* Zipc.of_binary_string zipc.ml:400-438.  Returns the members in CD order with duplicates
 * resolved as the reference's map does (later entry wins), then sorted by path.
 * Paths and payloads alias `s`.  *ms is malloc'ed. */
int zo_zip_decode(const uint8_t *s, size_t len, zo_member **ms, size_t *n);
/* File.to_binary_string zipc.ml:205-225 (stored / deflate extraction + CRC check).  On
 * ZO_ERR_CHECKSUM *found holds the computed CRC-32.  The reference prefixes inflate errors with
 * "deflate: " (zipc.ml:215); that is message glue and left to the caller. */
int zo_file_to_binary_string(const zo_member *m, uint8_t **out, size_t *out_len, uint32_t *found);
/* Member.make path/mode/mtime rules zipc.ml:244-255: returns malloc'ed normalised path. */
char *zo_member_make_path(const char *path, size_t path_len, int is_dir, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
