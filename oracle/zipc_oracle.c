/* zipc_oracle.c -- CPU restatement of dbuenzli/zipc's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * See zipc_oracle.h for the contract.  "ref:" comments cite /root/reference/src/zipc_deflate.ml
 * unless prefixed with zipc.ml.  The code follows the reference's *behaviour* statement by
 * statement where behaviour is observable (token stream, Huffman lengths, block choice, error
 * order) and is otherwise ordinary C.
 */
#include "zipc_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* switches                                                                                   */
/* ------------------------------------------------------------------------------------------ */
static int g_adler_signed_rem = 1;   /* ref :95,196  (mod) = Int32.rem */
static int g_keep_codelen_freqs = 1; /* ref :849-854 new_block does not clear codelen_sym_freqs */

void zo_set_adler_signed_rem(int on) { g_adler_signed_rem = on ? 1 : 0; }
void zo_set_keep_codelen_freqs(int on) { g_keep_codelen_freqs = on ? 1 : 0; }

const char *zo_strerror(int st) {
  switch (st) {
  case ZO_OK: return "";
  case ZO_ERR_CORRUPTED: return "Corrupted data stream";
  case ZO_ERR_SIZE_EXCEEDED: return "Expected decompression size exceeded";
  case ZO_ERR_ZLIB_METHOD: return "Unknown compression method (%d)";
  case ZO_ERR_ZLIB_WINDOW: return "Window size too large";
  case ZO_ERR_ZLIB_DICT: return "Preset dictionary unsupported";
  case ZO_ERR_CHECKSUM: return "Checksum mismatch, expected %lx found %lx)";
  case ZO_ERR_NOMEM: return "Out of memory";
  case ZO_ERR_ZIP_ZIP64: return "ZIP64 archives are not supported";
  case ZO_ERR_ZIP_MULTIPART: return "Multipart archives are not supported";
  case ZO_ERR_ZIP_EOCD: return "Corrupted end of central directory record";
  case ZO_ERR_ZIP_NO_EOCD:
    return "Likely not a ZIP archive: no end of central directory record found";
  case ZO_ERR_ZIP_SHORT: return "File too short to be a ZIP archive";
  case ZO_ERR_ZIP_TRUNC_CD: return "Truncated central directory";
  case ZO_ERR_ZIP_CDFH: return "Corrupted central directory file header";
  case ZO_ERR_ZIP_LFH: return "Corrupted local file header";
  case ZO_ERR_ZIP_COUNT: return "Maximum ZIP member count 65535 exceeded (%d)";
  case ZO_ERR_ZIP_PATH_LEN: return "Maximum ZIP path length 65535 exceeded (%d)";
  case ZO_ERR_ZIP_SIZE:
    return "Maximum ZIP byte size 4294967295 exceeded by compressed (%d) or decompressed (%d) "
           "file size";
  case ZO_ERR_ZIP_ENCRYPTED: return "Encrypted file not supported";
  case ZO_ERR_ZIP_FORMAT: return "Compression %a not supported";
  case ZO_ERR_ZIP_CD_OFFSET: return "Maximum ZIP central directory offset 4294967295 exceeded (%d)";
  case ZO_ERR_ZIP_CD_SIZE: return "Maximum ZIP central directory size 4294967295 exceeded (%d)";
  default: return "Unknown error";
  }
}

void zo_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* CRC-32   ref :106-164                                                                      */
/* ------------------------------------------------------------------------------------------ */
static uint32_t crc_tab[4][256];
static int crc_tab_ready = 0;

static void crc_tab_init(void) { /* ref :114-133 */
  for (unsigned i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xedb88320u ^ (c >> 1)) : (c >> 1);
    crc_tab[0][i] = c;
  }
  for (unsigned i = 0; i < 256; i++)
    for (int k = 1; k < 4; k++)
      crc_tab[k][i] = (crc_tab[k - 1][i] >> 8) ^ crc_tab[0][crc_tab[k - 1][i] & 0xff];
  crc_tab_ready = 1;
}

uint32_t zo_crc32_init(void) { return 0xFFFFFFFFu; }          /* ref :135 */
uint32_t zo_crc32_finish(uint32_t c) { return c ^ 0xFFFFFFFFu; } /* ref :136 */

uint32_t zo_crc32_update(uint32_t c, const uint8_t *s, size_t len) { /* ref :137-156 */
  if (!crc_tab_ready) crc_tab_init();
  size_t i = 0;
  while (i + 4 <= len) { /* one little-endian 32-bit word per step, ref :141-150 */
    uint32_t u = (uint32_t)s[i] | ((uint32_t)s[i + 1] << 8) | ((uint32_t)s[i + 2] << 16) |
                 ((uint32_t)s[i + 3] << 24);
    u ^= c;
    c = crc_tab[3][u & 0xff] ^ crc_tab[2][(u >> 8) & 0xff] ^ crc_tab[1][(u >> 16) & 0xff] ^
        crc_tab[0][u >> 24];
    i += 4;
  }
  for (; i < len; i++) c = (c >> 8) ^ crc_tab[0][(c ^ s[i]) & 0xff]; /* ref :151-155 */
  return c;
}

uint32_t zo_crc32(const uint8_t *s, size_t len) {
  return zo_crc32_finish(zo_crc32_update(zo_crc32_init(), s, len));
}

/* ------------------------------------------------------------------------------------------ */
/* Adler-32   ref :166-206                                                                    */
/* ------------------------------------------------------------------------------------------ */
static int32_t rem_base(int32_t v) { /* Uint32.Syntax.(mod) = Int32.rem, ref :95 */
  if (g_adler_signed_rem) return v % 65521; /* C99 % truncates like Int32.rem */
  return (int32_t)((uint32_t)v % 65521u);   /* RFC 1950 behaviour */
}

uint32_t zo_adler32_update(uint32_t a, const uint8_t *s, size_t len) { /* ref :175-198 */
  /* s1, s2 are OCaml int32: additions wrap (do them unsigned), rem is signed. */
  uint32_t s1 = a & 0xFFFFu, s2 = a >> 16;
  size_t start = 0;
  size_t block_len = len % 5552; /* first round may be empty, ref :180 */
  while (start < len) {          /* ref :181: while !start <= max */
    size_t end = start + block_len;
    for (size_t i = start; i < end; i++) { s1 += s[i]; s2 += s1; }
    s1 = (uint32_t)rem_base((int32_t)s1);
    s2 = (uint32_t)rem_base((int32_t)s2);
    start = end;
    block_len = 5552;
  }
  return (s2 << 16) + s1; /* ref :198 */
}

uint32_t zo_adler32(const uint8_t *s, size_t len) { return zo_adler32_update(1u, s, len); }

/* crc_op plumbing, ref :210-216 */
static uint32_t crc_op_init(int op) {
  return op == ZO_CRC_ADLER32 ? 1u : op == ZO_CRC_CRC32 ? 0xFFFFFFFFu : 0u;
}
static uint32_t crc_op_update(int op, uint32_t c, const uint8_t *s, size_t len) {
  if (op == ZO_CRC_ADLER32) return zo_adler32_update(c, s, len);
  if (op == ZO_CRC_CRC32) return zo_crc32_update(c, s, len);
  return c;
}
static uint32_t crc_op_finish(int op, uint32_t c) {
  return op == ZO_CRC_ADLER32 ? c : op == ZO_CRC_CRC32 ? (c ^ 0xFFFFFFFFu) : 0u;
}

/* ------------------------------------------------------------------------------------------ */
/* Deflate format tables   ref :233-313                                                       */
/* ------------------------------------------------------------------------------------------ */
enum { LITLEN_SYM_MAX = 285, EOB_SYM = 256, FIRST_LEN_SYM = 257, LEN_VALUE_MAX = 258 };
enum { DIST_SYM_MAX = 29, DIST_VALUE_MAX = 32768, CODELEN_SYM_MAX = 18 };
enum { HUFF_MAX_SYMS = 288, HUFF_MAX_BITS = 15 };

static const uint16_t len_base[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                      31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                      2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t dist_base[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                       33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6,
                                       6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t codelen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static uint16_t len_to_sym[LEN_VALUE_MAX + 1]; /* ref :260-267 */
static uint8_t dist_to_sym_tab[512];           /* ref :290-299 */
static int fmt_tabs_ready = 0;

static void fmt_tabs_init(void) {
  /* Later symbols overwrite earlier ones, so length 258 maps to 285 not 284 (ref :266). */
  for (int i = 0; i < 29; i++)
    for (int l = len_base[i]; l < len_base[i] + (1 << len_extra[i]); l++)
      if (l <= LEN_VALUE_MAX) len_to_sym[l] = (uint16_t)(257 + i);
  for (int i = 0; i < 30; i++)
    for (int d = dist_base[i]; d < dist_base[i] + (1 << dist_extra[i]); d++) {
      int k = d <= 256 ? d - 1 : 256 + ((d - 1) >> 7);
      dist_to_sym_tab[k] = (uint8_t)i;
    }
  fmt_tabs_ready = 1;
}
static inline int dist_to_sym(int d) { /* ref :301-302 */
  return dist_to_sym_tab[d <= 256 ? d - 1 : 256 + ((d - 1) >> 7)];
}

/* ------------------------------------------------------------------------------------------ */
/* Growable / fixed output buffer   ref :16-76                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint8_t *b;
  size_t cap, len;
  int fixed;
  int err; /* sticky status: ZO_ERR_SIZE_EXCEEDED / ZO_ERR_NOMEM */
} buf_t;

static int buf_make(buf_t *u, size_t sz, int fixed) { /* ref :18-20 */
  u->cap = (sz == 0 && !fixed) ? 1024 : sz;
  u->b = (uint8_t *)malloc(u->cap ? u->cap : 1);
  u->len = 0; u->fixed = fixed; u->err = 0;
  return u->b ? 0 : ZO_ERR_NOMEM;
}
static int buf_grow(buf_t *u, size_t ensure) { /* ref :27-38 */
  if (u->fixed) { u->err = ZO_ERR_SIZE_EXCEEDED; return u->err; }
  size_t n = u->cap ? u->cap : 1;
  while (n < ensure) n *= 2;
  uint8_t *nb = (uint8_t *)realloc(u->b, n);
  if (!nb) { u->err = ZO_ERR_NOMEM; return u->err; }
  u->b = nb; u->cap = n;
  return 0;
}
static inline int buf_add_u8(buf_t *u, unsigned byte) { /* ref :43-46 */
  if (u->len + 1 > u->cap && buf_grow(u, u->len + 1)) return u->err;
  u->b[u->len++] = (uint8_t)byte;
  return 0;
}
static int buf_add_bytes(buf_t *u, const uint8_t *s, size_t n) { /* ref :58-61 */
  if (u->len + n > u->cap && buf_grow(u, u->len + n)) return u->err;
  if (n) memcpy(u->b + u->len, s, n);
  u->len += n;
  return 0;
}
static int buf_recopy(buf_t *u, size_t start, size_t n) { /* ref :63-75 */
  if (u->len + n > u->cap && buf_grow(u, u->len + n)) return u->err;
  if (start + n <= u->len) memcpy(u->b + u->len, u->b + start, n);
  else for (size_t i = 0; i < n; i++) u->b[u->len + i] = u->b[start + i]; /* overlapping */
  u->len += n;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Huffman decoder   ref :324-391                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int counts[HUFF_MAX_BITS + 1];
  int symbols[HUFF_MAX_SYMS];
  int max_sym;
} hdec_t;

/* ref :355-391.  Returns 0 or ZO_ERR_CORRUPTED. */
static int hdec_init(hdec_t *t, const int *lengths, int n) {
  int offs[16];
  memset(t->counts, 0, sizeof t->counts);
  t->max_sym = -1;
  for (int i = 0; i < n; i++)
    if (lengths[i]) { t->max_sym = i; t->counts[lengths[i]]++; }
  int available = 1, num_codes = 0;
  for (int i = 0; i < 16; i++) {
    int used = t->counts[i];
    if (used > available) return ZO_ERR_CORRUPTED; /* over-subscribed, ref :371 */
    available = 2 * (available - used);
    offs[i] = num_codes;
    num_codes += used;
  }
  /* incomplete code: only tolerated when empty, or one code of length one.  ref :377-378 */
  if ((num_codes > 1 && available > 0) || (num_codes == 1 && t->counts[1] != 1))
    return ZO_ERR_CORRUPTED;
  for (int i = 0; i < n; i++)
    if (lengths[i]) t->symbols[offs[lengths[i]]++] = i;
  if (num_codes == 1) { t->counts[1] = 2; t->symbols[1] = t->max_sym + 1; } /* ref :389-390 */
  return 0;
}

static hdec_t fixed_litlen_dec, fixed_dist_dec;
static int fixed_dec_ready = 0;
static void fixed_dec_init(void) { /* ref :334-349 */
  memset(&fixed_litlen_dec, 0, sizeof fixed_litlen_dec);
  memset(&fixed_dist_dec, 0, sizeof fixed_dist_dec);
  fixed_litlen_dec.counts[7] = 24; fixed_litlen_dec.counts[8] = 152; fixed_litlen_dec.counts[9] = 112;
  for (int i = 0; i < 24; i++) fixed_litlen_dec.symbols[i] = 256 + i;
  for (int i = 24; i < 168; i++) fixed_litlen_dec.symbols[i] = i - 24;
  for (int i = 168; i < 176; i++) fixed_litlen_dec.symbols[i] = 112 + i;
  for (int i = 176; i < 288; i++) fixed_litlen_dec.symbols[i] = i - 32;
  fixed_litlen_dec.max_sym = LITLEN_SYM_MAX;
  fixed_dist_dec.counts[5] = 32;
  for (int i = 0; i < 32; i++) fixed_dist_dec.symbols[i] = i;
  fixed_dist_dec.max_sym = DIST_SYM_MAX;
  fixed_dec_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Inflate   ref :532-718                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const uint8_t *src;
  size_t src_len, src_pos;
  uint64_t bits;
  int bits_len;
  buf_t dst;
  hdec_t dyn_litlen, dyn_dist;
  int crc_op;
  uint32_t crc;
  size_t crc_next;
  int lengths[HUFF_MAX_SYMS + 32];
  int err;
} inf_t;

/* ref :564-579.  On running out of input sets err and returns 0. */
static inline unsigned inf_bits(inf_t *d, int count) {
  while (d->bits_len < count) {
    if (d->src_pos >= d->src_len) { d->err = ZO_ERR_CORRUPTED; return 0; }
    d->bits |= (uint64_t)d->src[d->src_pos++] << d->bits_len;
    d->bits_len += 8;
  }
  unsigned r = (unsigned)(d->bits & ((1ull << count) - 1));
  d->bits >>= count;
  d->bits_len -= count;
  return r;
}
static inline int inf_int(inf_t *d, int base, int bit_count) { /* ref :581-582 */
  return base + (bit_count ? (int)inf_bits(d, bit_count) : 0);
}
/* ref :584-591: canonical decode one bit at a time.  The reference indexes counts.(16) and dies
 * with Invalid_argument when the code is empty (latent hole); the oracle
 * reports "Corrupted data stream" instead. */
static int inf_symbol(inf_t *d, const hdec_t *h) {
  int len = 1, base = 0, offs = 0;
  for (;;) {
    offs = 2 * offs + (int)inf_bits(d, 1);
    if (d->err) return -1;
    if (len > HUFF_MAX_BITS) { d->err = ZO_ERR_CORRUPTED; return -1; }
    int count = h->counts[len];
    if (offs < count) return h->symbols[base + offs];
    len++; base += count; offs -= count;
  }
}

static int inf_block_symbols(inf_t *d, const hdec_t *hl, const hdec_t *hd) { /* ref :593-616 */
  for (;;) {
    int sym = inf_symbol(d, hl);
    if (d->err) return d->err;
    if (sym < EOB_SYM) {
      if (buf_add_u8(&d->dst, (unsigned)sym)) return d->err = d->dst.err;
      continue;
    }
    if (sym == EOB_SYM) return 0;
    if (sym > hl->max_sym || sym > LITLEN_SYM_MAX || hl->max_sym == -1)
      return d->err = ZO_ERR_CORRUPTED;
    int length = inf_int(d, len_base[sym - FIRST_LEN_SYM], len_extra[sym - FIRST_LEN_SYM]);
    if (d->err) return d->err;
    int dsym = inf_symbol(d, hd);
    if (d->err) return d->err;
    if (dsym > hd->max_sym || dsym > DIST_SYM_MAX) return d->err = ZO_ERR_CORRUPTED;
    int dist = inf_int(d, dist_base[dsym], dist_extra[dsym]);
    if (d->err) return d->err;
    if ((size_t)dist > d->dst.len) return d->err = ZO_ERR_CORRUPTED; /* ref :614 */
    if (buf_recopy(&d->dst, d->dst.len - (size_t)dist, (size_t)length)) return d->err = d->dst.err;
  }
}

static int inf_dynamic_block(inf_t *d) { /* ref :623-669 */
  int hlit = inf_int(d, 257, 5);
  if (d->err) return d->err;
  int hdist = inf_int(d, 1, 5);
  if (d->err) return d->err;
  if (hlit > LITLEN_SYM_MAX + 1 || hdist > DIST_SYM_MAX + 1) return d->err = ZO_ERR_CORRUPTED;
  /* code length code, ref :624-636 */
  int hclen = inf_int(d, 4, 4);
  if (d->err) return d->err;
  int *lengths = d->lengths;
  for (int i = 0; i <= CODELEN_SYM_MAX; i++) lengths[i] = 0;
  for (int i = 0; i < hclen; i++) {
    lengths[codelen_order[i]] = (int)inf_bits(d, 3);
    if (d->err) return d->err;
  }
  hdec_t *huff = &d->dyn_litlen; /* temporarily, as the reference does */
  if (hdec_init(huff, lengths, CODELEN_SYM_MAX + 1)) return d->err = ZO_ERR_CORRUPTED;
  if (huff->max_sym == -1) return d->err = ZO_ERR_CORRUPTED;
  int num = 0;
  while (num < hlit + hdist) { /* ref :647-661 */
    int sym = inf_symbol(d, huff);
    if (d->err) return d->err;
    if (sym > huff->max_sym) return d->err = ZO_ERR_CORRUPTED;
    int repeat, val;
    switch (sym) {
    case 16:
      if (num == 0) return d->err = ZO_ERR_CORRUPTED;
      repeat = inf_int(d, 3, 2); val = lengths[num - 1]; break;
    case 17: repeat = inf_int(d, 3, 3); val = 0; break;
    case 18: repeat = inf_int(d, 11, 7); val = 0; break;
    default: repeat = 1; val = sym; break;
    }
    if (d->err) return d->err;
    if (repeat > hlit + hdist - num) return d->err = ZO_ERR_CORRUPTED;
    while (repeat-- > 0) lengths[num++] = val;
  }
  if (lengths[256] == 0) return d->err = ZO_ERR_CORRUPTED; /* ref :662 */
  /* Both decoders are built from one array; copy since hdec_init of litlen may be the
   * same object as the code-length decoder (it is consumed by now). */
  if (hdec_init(&d->dyn_litlen, lengths, hlit)) return d->err = ZO_ERR_CORRUPTED;
  if (hdec_init(&d->dyn_dist, lengths + hlit, hdist)) return d->err = ZO_ERR_CORRUPTED;
  return inf_block_symbols(d, &d->dyn_litlen, &d->dyn_dist);
}

static int inf_stored_block(inf_t *d) { /* ref :671-680 */
  if (d->src_len - d->src_pos < 4) return d->err = ZO_ERR_CORRUPTED;
  unsigned length = d->src[d->src_pos] | ((unsigned)d->src[d->src_pos + 1] << 8);
  unsigned inv = d->src[d->src_pos + 2] | ((unsigned)d->src[d->src_pos + 3] << 8);
  if (length != ((~inv) & 0xFFFFu)) return d->err = ZO_ERR_CORRUPTED;
  d->src_pos += 4;
  if (d->src_len - d->src_pos < length) return d->err = ZO_ERR_CORRUPTED;
  if (buf_add_bytes(&d->dst, d->src + d->src_pos, length)) return d->err = d->dst.err;
  d->src_pos += length;
  d->bits = 0; d->bits_len = 0;
  return 0;
}

int zo_inflate(const uint8_t *src, size_t len, int64_t decompressed_size, int crc_op,
               uint8_t **out, size_t *out_len, uint32_t *crc) { /* ref :692-709 */
  if (!fixed_dec_ready) fixed_dec_init();
  inf_t *d = (inf_t *)calloc(1, sizeof *d);
  if (!d) return ZO_ERR_NOMEM;
  d->src = src; d->src_len = len; d->crc_op = crc_op; d->crc = crc_op_init(crc_op);
  int st = decompressed_size >= 0 ? buf_make(&d->dst, (size_t)decompressed_size, 1)
                                  : buf_make(&d->dst, len * 3, 0); /* ref :552-555 */
  if (st) { free(d); return st; }
  for (;;) {
    int final = (int)inf_bits(d, 1);
    if (d->err) break;
    int btype = (int)inf_bits(d, 2);
    if (d->err) break;
    if (btype == 0) inf_stored_block(d);
    else if (btype == 1) inf_block_symbols(d, &fixed_litlen_dec, &fixed_dist_dec);
    else if (btype == 2) inf_dynamic_block(d);
    else d->err = ZO_ERR_CORRUPTED;
    if (d->err) break;
    /* checksum of what this block produced, ref :682-690 */
    d->crc = crc_op_update(crc_op, d->crc, d->dst.b + d->crc_next, d->dst.len - d->crc_next);
    d->crc_next = d->dst.len;
    if (final) break;
  }
  st = d->err;
  if (st) { free(d->dst.b); if (out) *out = NULL; if (out_len) *out_len = 0; }
  else {
    if (crc) *crc = crc_op_finish(crc_op, d->crc);
    if (out_len) *out_len = d->dst.len;
    if (out) *out = d->dst.b; else free(d->dst.b);
  }
  free(d);
  return st;
}

int zo_zlib_decompress(const uint8_t *s, size_t len, int64_t decompressed_size,
                       uint8_t **out, size_t *out_len, uint32_t *adler,
                       uint32_t *expect_o, uint32_t *found_o, int *method) { /* ref :720-740 */
  if (out) *out = NULL;
  if (len < 6) return ZO_ERR_CORRUPTED;
  unsigned cmf = s[0], flg = s[1];
  if ((256 * cmf + flg) % 31 != 0) return ZO_ERR_CORRUPTED;
  unsigned cm = cmf & 0x0F;
  if (cm != 8) { if (method) *method = (int)cm; return ZO_ERR_ZLIB_METHOD; }
  if ((cmf >> 4) > 7) return ZO_ERR_ZLIB_WINDOW;
  if (flg & 0x20) return ZO_ERR_ZLIB_DICT;
  uint32_t expect = ((uint32_t)s[len - 4] << 24) | ((uint32_t)s[len - 3] << 16) |
                    ((uint32_t)s[len - 2] << 8) | s[len - 1];
  /* The reference hands inflate the range [2, 2 + len - 4): two bytes too long (it includes
   * half of the trailer); harmless because trailing input is ignored.  ref :732 */
  uint32_t found = 0;
  uint8_t *o = NULL; size_t on = 0;
  int st = zo_inflate(s + 2, len - 4, decompressed_size, ZO_CRC_ADLER32, &o, &on, &found);
  if (st) return st;
  if (expect != found) {
    free(o);
    if (expect_o) *expect_o = expect;
    if (found_o) *found_o = found;
    return ZO_ERR_CHECKSUM;
  }
  if (adler) *adler = found;
  if (out_len) *out_len = on;
  if (out) *out = o; else free(o);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Huffman encoder   ref :395-528                                                             */
/* ------------------------------------------------------------------------------------------ */
/* sym_info = (code << 5) | code_length, code stored bit-reversed.  ref :395-398 */
#define SYM_CODE(v) ((v) >> 5)
#define SYM_LEN(v) ((v)&0x1F)

static void heap_down(int64_t *h, int max, int i) { /* ref :408-416, 1-based min-heap */
  for (;;) {
    int l = 2 * i, r = l + 1;
    if (l > max) return;
    int k = r > max ? l : (h[l] < h[r] ? l : r);
    if (h[i] > h[k]) { int64_t t = h[i]; h[i] = h[k]; h[k] = t; i = k; }
    else return;
  }
}

/* ref :404-473.  e[0..max_sym] receives code lengths. */
static void huff_lengths_of_freqs(int *e, const int *freqs, int max_sym, int max_code_len) {
  int64_t heap[HUFF_MAX_SYMS * 2 + 1];
  int freq_cap = 65535;
  for (;;) {
    /* leaves, in increasing symbol order, link = max_sym + 1 + rank.  ref :421-431 */
    int max = 0;
    for (int sym = 0; sym <= max_sym; sym++) {
      int f = freqs[sym];
      if (!f) continue;
      if (f > freq_cap) f = freq_cap;
      max++;
      heap[max] = ((int64_t)f << 10) | (int64_t)(max_sym + 1 + max);
    }
    for (int i = max / 2; i >= 1; i--) heap_down(heap, max, i);
    if (max < 2) { /* ref :462-466 */
      for (int sym = 0; sym <= max_sym; sym++) e[sym] = freqs[sym] ? 1 : 0;
      return;
    }
    /* ref :432-445: merge the two least frequent nodes until one is left (index 2 = root) */
    for (int m = max; m > 1; m--) {
      int new_max = m - 1;
      int64_t p = heap[1];
      heap[1] = heap[m];
      heap_down(heap, new_max, 1);
      int64_t q = heap[1];
      int nlink = m;
      int64_t f = (p >> 10) + (q >> 10);
      heap[1] = (f << 10) | nlink;
      heap[p & 0x3FF] = nlink;
      heap[q & 0x3FF] = nlink;
      heap_down(heap, new_max, 1);
    }
    /* ref :446-460 */
    int k = 0, too_long = 0;
    for (int sym = 0; sym <= max_sym && !too_long; sym++) {
      if (!freqs[sym]) { e[sym] = 0; continue; }
      k++;
      int64_t p = heap[max_sym + 1 + k];
      int len = 1;
      while (p != 2) { len++; p = heap[p]; }
      if (len > max_code_len) too_long = 1; else e[sym] = len;
    }
    if (!too_long) return;
    freq_cap /= 2; /* flatten and retry, ref :470-473 */
  }
}

static unsigned reverse16(unsigned b) { /* ref :481-487 */
  b = ((b & 0xFF00) >> 8) | ((b & 0x00FF) << 8);
  b = ((b & 0xF0F0) >> 4) | ((b & 0x0F0F) << 4);
  b = ((b & 0xCCCC) >> 2) | ((b & 0x3333) << 2);
  b = ((b & 0xAAAA) >> 1) | ((b & 0x5555) << 1);
  return b;
}

/* ref :477-506: e holds lengths on entry, sym_infos on exit */
static void huff_init_with_lengths(int *e, int max_sym) {
  int count[16] = {0}, code[16] = {0};
  for (int s = 0; s <= max_sym; s++) count[SYM_LEN(e[s])]++;
  count[0] = 0;
  for (int l = 1; l <= 15; l++) code[l] = (code[l - 1] + count[l - 1]) << 1;
  for (int s = 0; s <= max_sym; s++) {
    int l = SYM_LEN(e[s]);
    if (!l) continue;
    unsigned bits = reverse16((unsigned)code[l]) >> (16 - l);
    e[s] = (int)((bits << 5) | (unsigned)l);
    code[l]++;
  }
}

static int fixed_litlen_enc[HUFF_MAX_SYMS], fixed_dist_enc[HUFF_MAX_SYMS];
static int fixed_enc_ready = 0;
static void fixed_enc_init(void) { /* ref :514-527 */
  for (int i = 0; i < 144; i++) fixed_litlen_enc[i] = 8;
  for (int i = 144; i < 256; i++) fixed_litlen_enc[i] = 9;
  for (int i = 256; i < 280; i++) fixed_litlen_enc[i] = 7;
  for (int i = 280; i < 288; i++) fixed_litlen_enc[i] = 8;
  huff_init_with_lengths(fixed_litlen_enc, 287);
  for (int i = 0; i < 32; i++) fixed_dist_enc[i] = 5;
  huff_init_with_lengths(fixed_dist_enc, 31);
  fixed_enc_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Deflate encoder   ref :742-1245                                                            */
/* ------------------------------------------------------------------------------------------ */
enum { WINDOW = 32768, HASH_BITS = 15, MAX_BLOCK_SRC = 65534, MIN_MATCH = 4 };

typedef struct {
  int level;
  const uint8_t *src;
  size_t src_len;
  buf_t dst;
  uint64_t dst_bits;
  int dst_bits_len;
  uint32_t *block_syms; /* (dist << 9) | len ; literal: dist 0, len = byte.  ref :766-776 */
  int block_syms_len;
  size_t block_src_start;
  int block_src_len;
  int litlen_freqs[LITLEN_SYM_MAX + 1];
  int dist_freqs[DIST_SYM_MAX + 1];
  int codelen_syms[LITLEN_SYM_MAX + DIST_SYM_MAX + 2]; /* (repeat_bits << 8) | sym */
  int codelen_syms_len;
  int codelen_freqs[CODELEN_SYM_MAX + 1];
  int good_match, max_chain;
  int64_t *head, *prev;
  int dyn_litlen[HUFF_MAX_SYMS], dyn_dist[HUFF_MAX_SYMS], dyn_codelen[HUFF_MAX_SYMS];
  int hlit, hdist, hclen;
  int crc_op;
  uint32_t crc;
  zo_deflate_stats st;
} enc_t;

static void enc_free(enc_t *e) {
  free(e->block_syms); free(e->head); free(e->prev);
}

static int enc_make(enc_t *e, int level, const uint8_t *src, size_t len, int crc_op) { /* ref :817-847 */
  if (!fmt_tabs_ready) fmt_tabs_init();
  if (!fixed_enc_ready) fixed_enc_init();
  memset(e, 0, sizeof *e);
  e->level = level; e->src = src; e->src_len = len; e->crc_op = crc_op;
  e->crc = crc_op_init(crc_op);
  switch (level) { /* ref :754-764; max_lazy and nice_length are never read */
  case ZO_LEVEL_FAST: e->good_match = 4; e->max_chain = 4; break;
  case ZO_LEVEL_DEFAULT: e->good_match = 8; e->max_chain = 128; break;
  case ZO_LEVEL_BEST: e->good_match = 32; e->max_chain = 4096; break;
  default: break;
  }
  if (buf_make(&e->dst, len, 0)) return ZO_ERR_NOMEM;
  e->block_syms = (uint32_t *)malloc(sizeof(uint32_t) * (MAX_BLOCK_SRC + 1));
  e->head = (int64_t *)malloc(sizeof(int64_t) << HASH_BITS);
  e->prev = (int64_t *)malloc(sizeof(int64_t) * WINDOW);
  if (!e->block_syms || !e->head || !e->prev) { free(e->dst.b); enc_free(e); return ZO_ERR_NOMEM; }
  for (int i = 0; i < (1 << HASH_BITS); i++) e->head[i] = -1;
  memset(e->prev, 0, sizeof(int64_t) * WINDOW);
  return 0;
}

static void enc_flush(enc_t *e) { /* ref :856-858 */
  if (e->dst_bits_len > 0) { buf_add_u8(&e->dst, (unsigned)e->dst_bits); e->dst_bits = 0; e->dst_bits_len = 0; }
}
static void enc_bits(enc_t *e, uint64_t v, int count) { /* ref :864-871 */
  e->dst_bits |= v << e->dst_bits_len;
  e->dst_bits_len += count;
  while (e->dst_bits_len >= 8) {
    buf_add_u8(&e->dst, (unsigned)(e->dst_bits & 0xff));
    e->dst_bits >>= 8;
    e->dst_bits_len -= 8;
  }
}
static void enc_u16le(enc_t *e, unsigned v) {
  buf_add_u8(&e->dst, v & 0xff); buf_add_u8(&e->dst, (v >> 8) & 0xff);
}

static void enc_write_stored(enc_t *e, int final) { /* ref :873-877 */
  unsigned len = (unsigned)e->block_src_len;
  enc_bits(e, final ? 1 : 0, 3);
  enc_flush(e);
  enc_u16le(e, len);
  enc_u16le(e, ~len);
  buf_add_bytes(&e->dst, e->src + e->block_src_start, len);
  e->st.blocks_stored++;
}

static void enc_write_syms(enc_t *e, const int *hl, const int *hd) { /* ref :879-910 */
  for (int i = 0; i < e->block_syms_len; i++) {
    uint32_t b = e->block_syms[i];
    int dist = (int)(b >> 9), len = (int)(b & 0x1FF);
    if (dist == 0) {
      enc_bits(e, (uint64_t)SYM_CODE(hl[len]), SYM_LEN(hl[len]));
    } else {
      int ls = len_to_sym[len];
      int cnt = SYM_LEN(hl[ls]);
      uint64_t bits = ((uint64_t)(len - len_base[ls - 257]) << cnt) | (uint64_t)SYM_CODE(hl[ls]);
      enc_bits(e, bits, cnt + len_extra[ls - 257]);
      int ds = dist_to_sym(dist);
      cnt = SYM_LEN(hd[ds]);
      bits = ((uint64_t)(dist - dist_base[ds]) << cnt) | (uint64_t)SYM_CODE(hd[ds]);
      enc_bits(e, bits, cnt + dist_extra[ds]);
    }
  }
}

static void enc_write_fixed(enc_t *e, int final) { /* ref :912-916 */
  enc_bits(e, final ? 3 : 2, 3);
  enc_write_syms(e, fixed_litlen_enc, fixed_dist_enc);
  e->st.blocks_fixed++;
}

static void enc_write_dynamic(enc_t *e, int final) { /* ref :918-945 */
  enc_bits(e, final ? 5 : 4, 3);
  enc_bits(e, (uint64_t)e->hlit, 5);
  enc_bits(e, (uint64_t)e->hdist, 5);
  enc_bits(e, (uint64_t)e->hclen, 4);
  for (int o = 0; o < e->hclen + 4; o++) enc_bits(e, (uint64_t)SYM_LEN(e->dyn_codelen[codelen_order[o]]), 3);
  for (int l = 0; l < e->codelen_syms_len; l++) {
    int ref = e->codelen_syms[l], sym = ref & 0xFF, info = e->dyn_codelen[sym];
    uint64_t bits = (uint64_t)SYM_CODE(info);
    int cnt = SYM_LEN(info);
    if (sym <= 15) enc_bits(e, bits, cnt);
    else {
      int rb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
      enc_bits(e, ((uint64_t)(ref >> 8) << cnt) | bits, cnt + rb);
    }
  }
  enc_write_syms(e, e->dyn_litlen, e->dyn_dist);
  e->st.blocks_dynamic++;
}

static void enc_huff_from_freqs(int *h, const int *freqs, int max_sym, int max_len) { /* ref :508-512 */
  huff_lengths_of_freqs(h, freqs, max_sym, max_len);
  huff_init_with_lengths(h, max_sym);
}

static void enc_make_dynamic_encoding(enc_t *e) { /* ref :959-1043 */
  int lengths[LITLEN_SYM_MAX + DIST_SYM_MAX + 2];
  int lc = LITLEN_SYM_MAX;
  while (lc >= 0 && SYM_LEN(e->dyn_litlen[lc]) == 0) lc--;
  int litlen_count = lc + 1;
  int dc = DIST_SYM_MAX;
  while (dc >= 0 && SYM_LEN(e->dyn_dist[dc]) == 0) dc--;
  int dist_count = dc + 1;
  if (dist_count == 0) { e->dyn_dist[0] = (0 << 5) | 1; dist_count = 1; } /* ref :974-979 */
  e->hlit = litlen_count - 257;
  e->hdist = dist_count - 1;
  for (int i = 0; i < litlen_count; i++) lengths[i] = SYM_LEN(e->dyn_litlen[i]);
  for (int i = 0; i < dist_count; i++) lengths[litlen_count + i] = SYM_LEN(e->dyn_dist[i]);
  int len_max = litlen_count + dist_count - 1;
  /* run-length encode, ref :996-1030 */
  int k = 0, i = 0;
  while (i <= len_max) {
    int v = lengths[i];
    if (v == 0) {
      int max = len_max < i + 138 - 1 ? len_max : i + 138 - 1;
      int j = i + 1;
      while (j <= max && lengths[j] == 0) j++;
      int z = j - i;
      if (z < 3) { e->codelen_syms[k++] = 0; e->codelen_freqs[0]++; i += 1; }
      else if (z <= 10) { e->codelen_syms[k++] = ((z - 3) << 8) | 17; e->codelen_freqs[17]++; i = j; }
      else { e->codelen_syms[k++] = ((z - 11) << 8) | 18; e->codelen_freqs[18]++; i = j; }
    } else {
      e->codelen_syms[k++] = v; e->codelen_freqs[v]++;
      int max = len_max < i + 6 ? len_max : i + 6;
      int j = i + 1;
      while (j <= max && lengths[j] == v) j++;
      int sc = j - i;
      if (sc <= 3) i += 1;
      else { e->codelen_syms[k++] = ((sc - 3 - 1) << 8) | 16; e->codelen_freqs[16]++; i = j; }
    }
  }
  e->codelen_syms_len = k;
  enc_huff_from_freqs(e->dyn_codelen, e->codelen_freqs, CODELEN_SYM_MAX, 7);
  int o = CODELEN_SYM_MAX; /* ref :1032-1037 */
  while (o > 0 && SYM_LEN(e->dyn_codelen[codelen_order[o]]) == 0) o--;
  e->hclen = o + 1 - 4;
}

static int64_t enc_bits_of_syms(const enc_t *e, const int *hl, const int *hd) { /* ref :1049-1064 */
  int64_t acc = 0;
  for (int s = 0; s <= LITLEN_SYM_MAX; s++) {
    int extra = s < FIRST_LEN_SYM ? 0 : len_extra[s - 257];
    acc += (int64_t)e->litlen_freqs[s] * (SYM_LEN(hl[s]) + extra);
  }
  for (int s = 0; s <= DIST_SYM_MAX; s++)
    acc += (int64_t)e->dist_freqs[s] * (SYM_LEN(hd[s]) + dist_extra[s]);
  return acc;
}

static void enc_write_block(enc_t *e, int final) { /* ref :1094-1104 */
  e->crc = crc_op_update(e->crc_op, e->crc, e->src + e->block_src_start, (size_t)e->block_src_len);
  e->block_syms[e->block_syms_len++] = EOB_SYM; /* ref :1088-1092 */
  e->litlen_freqs[EOB_SYM] = 1;
  enc_huff_from_freqs(e->dyn_litlen, e->litlen_freqs, LITLEN_SYM_MAX, 15); /* ref :953-957 */
  enc_huff_from_freqs(e->dyn_dist, e->dist_freqs, DIST_SYM_MAX, 15);
  enc_make_dynamic_encoding(e);
  int64_t nlen = 3 + (8 - ((e->dst_bits_len + 3) % 8)) + (int64_t)(4 + e->block_src_len) * 8; /* :1045-1047 */
  int64_t flen = 3 + enc_bits_of_syms(e, fixed_litlen_enc, fixed_dist_enc);                 /* :1066-1069 */
  int64_t dlen = 3 + 5 + 5 + 4 + 3 * (e->hclen + 4);                                         /* :1071-1079 */
  for (int s = 0; s <= CODELEN_SYM_MAX; s++) {
    int rb = s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0;
    dlen += (int64_t)e->codelen_freqs[s] * (SYM_LEN(e->dyn_codelen[s]) + rb);
  }
  dlen += enc_bits_of_syms(e, e->dyn_litlen, e->dyn_dist);
  if (nlen <= dlen && nlen <= flen) enc_write_stored(e, final);
  else if (flen <= dlen) enc_write_fixed(e, final);
  else enc_write_dynamic(e, final);
}

static void enc_new_block(enc_t *e) { /* ref :849-854 */
  e->block_syms_len = 0;
  e->block_src_start += (size_t)e->block_src_len;
  e->block_src_len = 0;
  memset(e->litlen_freqs, 0, sizeof e->litlen_freqs);
  memset(e->dist_freqs, 0, sizeof e->dist_freqs);
  if (!g_keep_codelen_freqs) memset(e->codelen_freqs, 0, sizeof e->codelen_freqs);
}

static inline void enc_block_sym(enc_t *e, uint32_t sym, int src_len) { /* ref :1118-1123 */
  if (e->block_src_len + src_len > MAX_BLOCK_SRC) { enc_write_block(e, 0); enc_new_block(e); }
  e->block_syms[e->block_syms_len++] = sym;
  e->block_src_len += src_len;
}
static inline void enc_lit(enc_t *e, unsigned byte) { /* ref :1125-1128 */
  enc_block_sym(e, byte, 1);
  e->litlen_freqs[byte]++;
  e->st.literals++;
}
static inline void enc_backref(enc_t *e, uint32_t bref) { /* ref :1130-1136 */
  int len = (int)(bref & 0x1FF);
  enc_block_sym(e, bref, len);
  e->litlen_freqs[len_to_sym[len]]++;
  e->dist_freqs[dist_to_sym((int)(bref >> 9))]++;
  e->st.matches++; e->st.match_bytes += (uint64_t)len;
}

static void enc_all_stored(enc_t *e) { /* ref :1106-1116 */
  int64_t src_max = (int64_t)e->src_len - 1;
  for (;;) {
    int64_t start = (int64_t)e->block_src_start;
    int64_t block_max = start + MAX_BLOCK_SRC - 1;
    if (src_max < block_max) block_max = src_max;
    int len = (int)(block_max - start + 1);
    int final = block_max == src_max;
    e->block_src_len = len;
    e->crc = crc_op_update(e->crc_op, e->crc, e->src + e->block_src_start, (size_t)len);
    enc_write_stored(e, final);
    if (final) return;
    e->block_src_start = (size_t)(start + len);
  }
}

/* Lz77, ref :1140-1245 */
static inline uint32_t lz_hash4(const uint8_t *s, int64_t i) { /* ref :1145-1148 */
  uint32_t v = (uint32_t)s[i] | ((uint32_t)s[i + 1] << 8) | ((uint32_t)s[i + 2] << 16) |
               ((uint32_t)s[i + 3] << 24);
  return (uint32_t)(v * 0x9E3779B1u) >> (32 - HASH_BITS);
}
static inline void lz_insert(enc_t *e, uint32_t h, int64_t pos) { /* ref :1150-1152 */
  e->prev[pos % WINDOW] = e->head[h];
  e->head[h] = pos;
}
/* ref :1154-1174: common prefix of s[i..] and s[j..] if strictly longer than prev_len, else 0.
 * Bytes prev_len, prev_len-1 .. 0 are compared first, then the match is extended forward. */
static inline int lz_match_len(const uint8_t *s, int64_t i, int64_t j, int prev_len, int max_len) {
  for (int t = prev_len; t >= 0; t--)
    if (s[i + t] != s[j + t]) return 0;
  int len = prev_len + 1;
  while (len < max_len && s[i + len] == s[j + len]) len++;
  return len;
}
static uint32_t lz_find_backref(enc_t *e, int64_t pos, uint32_t hash, int prev_len, int max_len) { /* ref :1176-1201 */
  if (prev_len == 0) prev_len = MIN_MATCH - 1;
  if (prev_len >= max_len) return 0;
  int steps = e->max_chain;
  if (prev_len >= e->good_match) steps /= 4; /* decided once at entry */
  int64_t i = e->head[hash], match_pos = -1;
  for (;;) {
    if (i == -1 || steps == 0 || pos - i > DIST_VALUE_MAX)
      return match_pos == -1 ? 0u : (((uint32_t)(pos - match_pos) << 9) | (uint32_t)prev_len);
    steps--;
    int len = lz_match_len(e->src, i, pos, prev_len, max_len);
    if (len == max_len) return ((uint32_t)(pos - i) << 9) | (uint32_t)len;
    if (len != 0) { match_pos = i; prev_len = len; }
    i = e->prev[i % WINDOW];
  }
}

static void lz_compress(enc_t *e) { /* ref :1203-1244 */
  if (e->level == ZO_LEVEL_NONE) { enc_all_stored(e); return; }
  const uint8_t *s = e->src;
  int64_t n = (int64_t)e->src_len, max_pos = n - MIN_MATCH, i = 0;
  uint32_t prev = 0;
  for (;;) {
    int prev_len = (int)(prev & 0x1FF);
    if (i > max_pos) {
      if (prev_len != 0) { enc_backref(e, prev); i = max_pos + prev_len; }
      for (int64_t k = i; k < n; k++) enc_lit(e, s[k]);
      enc_write_block(e, 1);
      enc_flush(e);
      return;
    }
    uint32_t h = lz_hash4(s, i);
    int max_len = n - i < LEN_VALUE_MAX ? (int)(n - i) : LEN_VALUE_MAX;
    uint32_t bref = lz_find_backref(e, i, h, prev_len, max_len);
    int match_len = (int)(bref & 0x1FF);
    lz_insert(e, h, i);
    if (prev_len != 0 && prev_len > match_len) {
      /* previous match stands: emit it, hash the positions it covers, jump.  ref :1224-1233 */
      enc_backref(e, prev);
      int64_t next = (i - 1) + prev_len;
      int64_t last = next - 1 < max_pos ? next - 1 : max_pos;
      for (int64_t j = i + 1; j <= last; j++) lz_insert(e, lz_hash4(s, j), j);
      i = next; prev = 0;
    } else if (match_len == 0) {
      enc_lit(e, s[i]); i++; prev = 0;
    } else {
      if (prev_len != 0) enc_lit(e, s[i - 1]);
      i++; prev = bref;
    }
  }
}

int zo_deflate(int level, const uint8_t *src, size_t len, int crc_op,
               uint8_t **out, size_t *out_len, uint32_t *crc, zo_deflate_stats *stats) { /* ref :1247-1251 */
  enc_t *e = (enc_t *)malloc(sizeof *e);
  if (!e) return ZO_ERR_NOMEM;
  int st = enc_make(e, level, src, len, crc_op);
  if (st) { free(e); return st; }
  lz_compress(e);
  if (e->dst.err) { st = e->dst.err; free(e->dst.b); if (out) *out = NULL; }
  else {
    if (crc) *crc = crc_op_finish(crc_op, e->crc);
    if (out_len) *out_len = e->dst.len;
    if (stats) *stats = e->st;
    if (out) *out = e->dst.b; else free(e->dst.b);
  }
  enc_free(e); free(e);
  return st;
}

int zo_zlib_compress(int level, const uint8_t *src, size_t len,
                     uint8_t **out, size_t *out_len, uint32_t *adler) { /* ref :1262-1277 */
  enc_t *e = (enc_t *)malloc(sizeof *e);
  if (!e) return ZO_ERR_NOMEM;
  int st = enc_make(e, level, src, len, ZO_CRC_ADLER32);
  if (st) { free(e); return st; }
  unsigned cmf = (7u << 4) | 8u;
  unsigned flevel = level == ZO_LEVEL_NONE ? 0 : level == ZO_LEVEL_FAST ? 1 : level == ZO_LEVEL_DEFAULT ? 2 : 3;
  unsigned header = (cmf << 8) | (flevel << 6);
  unsigned flg = (header + 31 - (header % 31)) & 0xFF;
  buf_add_u8(&e->dst, cmf);
  buf_add_u8(&e->dst, flg);
  lz_compress(e);
  uint32_t c = crc_op_finish(ZO_CRC_ADLER32, e->crc);
  buf_add_u8(&e->dst, c >> 24); buf_add_u8(&e->dst, (c >> 16) & 0xff);
  buf_add_u8(&e->dst, (c >> 8) & 0xff); buf_add_u8(&e->dst, c & 0xff);
  if (e->dst.err) { st = e->dst.err; free(e->dst.b); if (out) *out = NULL; }
  else {
    if (adler) *adler = c;
    if (out_len) *out_len = e->dst.len;
    if (out) *out = e->dst.b; else free(e->dst.b);
  }
  enc_free(e); free(e);
  return st;
}

/* ------------------------------------------------------------------------------------------ */
/* Ptime   zipc.ml:64-125                                                                     */
/* ------------------------------------------------------------------------------------------ */
#define DOS_EPOCH 315532800LL

void zo_ptime_to_date_time(int64_t t, int *y, int *mo, int *d, int *hh, int *mm, int *ss) { /* zipc.ml:67-86 */
  int64_t jd = t / 86400 + 2440588, r = t % 86400;
  *hh = (int)(r / 3600); *mm = (int)((r % 3600) / 60); *ss = (int)((r % 3600) % 60);
  int64_t a = jd + 32044, b = (4 * a + 3) / 146097, c = a - (146097 * b) / 4;
  int64_t dd = (4 * c + 3) / 1461, e = c - (1461 * dd) / 4, m = (5 * e + 2) / 153;
  *d = (int)(e - (153 * m + 2) / 5 + 1);
  *mo = (int)(m + 3 - 12 * (m / 10));
  *y = (int)(100 * b + dd - 4800 + m / 10);
}
int64_t zo_ptime_of_dos(int dos_date, int dos_time) { /* zipc.ml:96-113 */
  if (dos_date < 0x21) return DOS_EPOCH;
  int hh = dos_time >> 11, mm = (dos_time >> 5) & 0x3F, ss = (dos_time & 0x1F) * 2;
  int year = ((dos_date >> 9) & 0x7F) + 1980, month = (dos_date >> 5) & 0xF, day = dos_date & 0x1F;
  int64_t a = (14 - month) / 12, y = year + 4800 - a, m = month + 12 * a - 3;
  int64_t jd = day + (153 * m + 2) / 5 + 365 * y + y / 4 - y / 100 + y / 400 - 32045;
  return (jd - 2440588) * 86400 + hh * 3600 + mm * 60 + ss;
}
void zo_ptime_to_dos(int64_t t, int *dos_date, int *dos_time) { /* zipc.ml:115-124 */
  int y, mo, d, hh, mm, ss;
  zo_ptime_to_date_time(t, &y, &mo, &d, &hh, &mm, &ss);
  if (y < 1980) { y = 1980; mo = 1; d = 1; hh = mm = ss = 0; }
  else if (y > 2107) { y = 2107; mo = 12; d = 31; hh = 23; mm = 59; ss = 59; }
  *dos_date = d | (mo << 5) | ((y - 1980) << 9);
  *dos_time = (ss / 2) | (mm << 5) | (hh << 11);
}

/* ------------------------------------------------------------------------------------------ */
/* Member.make path rules   zipc.ml:41-45,244-255                                             */
/* ------------------------------------------------------------------------------------------ */
char *zo_member_make_path(const char *path, size_t n, int is_dir, size_t *out_len) {
  char *p = (char *)malloc(n + 3);
  if (!p) return NULL;
  for (size_t i = 0; i < n; i++) p[i] = path[i] == '\\' ? '/' : path[i];
  if (is_dir) {
    if (n == 0) { p[0] = '.'; p[1] = '/'; n = 2; }
    else if (p[n - 1] != '/') p[n++] = '/';
  }
  p[n] = 0;
  *out_len = n;
  return p;
}

/* ------------------------------------------------------------------------------------------ */
/* Archive encode   zipc.ml:442-588                                                           */
/* ------------------------------------------------------------------------------------------ */
static void put16(uint8_t *b, size_t o, unsigned v) { b[o] = v & 0xff; b[o + 1] = (v >> 8) & 0xff; }
static void put32(uint8_t *b, size_t o, uint32_t v) { put16(b, o, v & 0xffff); put16(b, o + 2, v >> 16); }
static unsigned get16(const uint8_t *b, size_t o) { return b[o] | ((unsigned)b[o + 1] << 8); }
static uint32_t get32(const uint8_t *b, size_t o) { return get16(b, o) | ((uint32_t)get16(b, o + 2) << 16); }

static int path_cmp(const zo_member *a, const zo_member *b) { /* String.compare: byte-wise */
  size_t n = a->path_len < b->path_len ? a->path_len : b->path_len;
  int c = n ? memcmp(a->path, b->path, n) : 0;
  if (c) return c;
  return a->path_len < b->path_len ? -1 : a->path_len > b->path_len ? 1 : 0;
}

/* Sort indices by path, stable, and drop all but the last of equal paths (String_map.add). */
static size_t order_members(const zo_member *ms, size_t n, size_t *idx) {
  for (size_t i = 0; i < n; i++) idx[i] = i;
  /* merge sort for stability */
  size_t *tmp = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  for (size_t w = 1; w < n; w *= 2) {
    for (size_t lo = 0; lo < n; lo += 2 * w) {
      size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      size_t a = lo, b = mid, k = lo;
      while (a < mid && b < hi) tmp[k++] = path_cmp(&ms[idx[b]], &ms[idx[a]]) < 0 ? idx[b++] : idx[a++];
      while (a < mid) tmp[k++] = idx[a++];
      while (b < hi) tmp[k++] = idx[b++];
    }
    memcpy(idx, tmp, sizeof(size_t) * n);
  }
  free(tmp);
  size_t m = 0;
  for (size_t i = 0; i < n; i++) {
    if (i + 1 < n && path_cmp(&ms[idx[i]], &ms[idx[i + 1]]) == 0) continue; /* later one wins */
    idx[m++] = idx[i];
  }
  return m;
}

uint64_t zo_zip_encoding_size(const zo_member *ms, size_t n) { /* zipc.ml:447-455 */
  size_t *idx = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  size_t m = order_members(ms, n, idx);
  uint64_t acc = 22;
  for (size_t i = 0; i < m; i++) {
    const zo_member *x = &ms[idx[i]];
    acc += 30 + x->path_len + (x->is_dir ? 0 : x->compressed_size) + 46 + x->path_len;
  }
  free(idx);
  return acc;
}

static size_t encode_lfh(uint8_t *b, size_t start, const zo_member *m) { /* zipc.ml:457-496 */
  int date, time;
  zo_ptime_to_dos(m->mtime, &date, &time);
  put32(b, start, 0x04034b50u);
  put16(b, start + 10, (unsigned)time);
  put16(b, start + 12, (unsigned)date);
  put16(b, start + 26, m->path_len);
  put16(b, start + 28, 0);
  memcpy(b + start + 30, m->path, m->path_len);
  if (m->is_dir) {
    put16(b, start + 4, 20); put16(b, start + 6, 0x800); put16(b, start + 8, 0);
    put32(b, start + 14, 0); put32(b, start + 18, 0); put32(b, start + 22, 0);
    return start + 30 + m->path_len;
  }
  put16(b, start + 4, (unsigned)m->version_needed);
  put16(b, start + 6, (unsigned)m->gp_flags & ~(1u << 3)); /* zipc.ml:442-445 */
  put16(b, start + 8, (unsigned)m->compression);
  put32(b, start + 14, m->crc32);
  put32(b, start + 18, (uint32_t)m->compressed_size);
  put32(b, start + 22, (uint32_t)m->decompressed_size);
  size_t p = start + 30 + m->path_len;
  if (m->compressed_size) memcpy(b + p, m->compressed_bytes + m->start, m->compressed_size);
  return p + m->compressed_size;
}

static size_t encode_cdfh(uint8_t *b, size_t start, size_t lfh_offset, const zo_member *m) { /* zipc.ml:498-545 */
  int date, time;
  zo_ptime_to_dos(m->mtime, &date, &time);
  unsigned hi = (m->is_dir ? 040000u : 0100000u) | ((unsigned)m->mode & 07777u);
  unsigned lo = m->is_dir ? 0x10u : 0u;
  put32(b, start, 0x02014b50u);
  put16(b, start + 12, (unsigned)time);
  put16(b, start + 14, (unsigned)date);
  put16(b, start + 28, m->path_len);
  put16(b, start + 30, 0); put16(b, start + 32, 0); put16(b, start + 34, 0); put16(b, start + 36, 0);
  put16(b, start + 38, lo);
  put16(b, start + 40, hi);
  put32(b, start + 42, (uint32_t)lfh_offset);
  memcpy(b + start + 46, m->path, m->path_len);
  if (m->is_dir) {
    put16(b, start + 4, 0x314); put16(b, start + 6, 20); put16(b, start + 8, 0x800);
    put16(b, start + 10, 0);
    put32(b, start + 16, 0); put32(b, start + 20, 0); put32(b, start + 24, 0);
  } else {
    put16(b, start + 4, (unsigned)m->version_made_by);
    put16(b, start + 6, (unsigned)m->version_needed);
    put16(b, start + 8, (unsigned)m->gp_flags & ~(1u << 3));
    put16(b, start + 10, (unsigned)m->compression);
    put32(b, start + 16, m->crc32);
    put32(b, start + 20, (uint32_t)m->compressed_size);
    put32(b, start + 24, (uint32_t)m->decompressed_size);
  }
  return start + 46 + m->path_len;
}

int zo_zip_encode(const zo_member *ms, size_t n, const char *first, uint8_t **out, size_t *out_len) { /* zipc.ml:570-588 */
  if (!first) first = "mimetype";
  size_t flen = strlen(first);
  size_t *idx = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  size_t *lfh = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
  size_t m = order_members(ms, n, idx);
  if (m > 0xFFFF) { free(idx); free(lfh); return ZO_ERR_ZIP_COUNT; }
  uint64_t total = 22;
  for (size_t i = 0; i < m; i++) {
    const zo_member *x = &ms[idx[i]];
    total += 30 + x->path_len + (x->is_dir ? 0 : x->compressed_size) + 46 + x->path_len;
  }
  uint8_t *b = (uint8_t *)malloc(total ? total : 1);
  if (!b) { free(idx); free(lfh); return ZO_ERR_NOMEM; }
  /* `first` goes first, then map order.  zipc.ml:575-580 */
  size_t fi = m;
  for (size_t i = 0; i < m; i++)
    if (ms[idx[i]].path_len == flen && memcmp(ms[idx[i]].path, first, flen) == 0) fi = i;
  if (fi < m) {
    size_t f = idx[fi];
    memmove(idx + 1, idx, sizeof(size_t) * fi);
    idx[0] = f;
  }
  size_t pos = 0;
  for (size_t i = 0; i < m; i++) { lfh[i] = pos; pos = encode_lfh(b, pos, &ms[idx[i]]); }
  size_t cd_start = pos;
  for (size_t i = 0; i < m; i++) pos = encode_cdfh(b, pos, lfh[i], &ms[idx[i]]);
  size_t cd_size = pos - cd_start;
  free(idx); free(lfh);
  if (m == 0) { cd_start = 0; cd_size = 0; }
  if (cd_start > 0xFFFFFFFFull) { free(b); return ZO_ERR_ZIP_CD_OFFSET; }
  if (cd_size > 0xFFFFFFFFull) { free(b); return ZO_ERR_ZIP_CD_SIZE; }
  put32(b, pos, 0x06054b50u); /* zipc.ml:553-566 */
  put16(b, pos + 4, 0); put16(b, pos + 6, 0);
  put16(b, pos + 8, (unsigned)m); put16(b, pos + 10, (unsigned)m);
  put32(b, pos + 12, (uint32_t)cd_size); put32(b, pos + 16, (uint32_t)cd_start);
  put16(b, pos + 20, 0);
  *out = b; *out_len = (size_t)total;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Archive decode   zipc.ml:314-438                                                           */
/* ------------------------------------------------------------------------------------------ */
int zo_zip_decode(const uint8_t *s, size_t len, zo_member **out, size_t *out_n) {
  *out = NULL; *out_n = 0;
  /* find_cd_info_in_eocd, zipc.ml:400-425 */
  if (len < 22) return ZO_ERR_ZIP_SHORT;
  int64_t start = (int64_t)len - 22, min_start = (int64_t)len - 65535 - 22;
  for (;; start--) {
    if (start < min_start || start < 0) return ZO_ERR_ZIP_NO_EOCD;
    if (get32(s, (size_t)start) == 0x06054b50u) break;
  }
  size_t e = (size_t)start;
  unsigned disk = get16(s, e + 4), disk_cd = get16(s, e + 6);
  if (disk == 0xFFFF) return ZO_ERR_ZIP_ZIP64;
  if (disk != 0 || disk_cd != 0) return ZO_ERR_ZIP_MULTIPART;
  size_t count = get16(s, e + 10);
  uint64_t cd_size = get32(s, e + 12), cd_start = get32(s, e + 16);
  if (cd_start + cd_size > len) return ZO_ERR_ZIP_EOCD;
  int64_t cd_max = (int64_t)(cd_start + cd_size) - 1;
  zo_member *ms = (zo_member *)calloc(count ? count : 1, sizeof *ms);
  if (!ms) return ZO_ERR_NOMEM;
  int64_t i = (int64_t)cd_start;
  for (size_t k = 0; k < count; k++) { /* decode_cd_members, zipc.ml:392-396 */
    if (i > cd_max) { free(ms); return ZO_ERR_ZIP_TRUNC_CD; }
    /* decode_member_of_cd, zipc.ml:344-390 */
    if (i + 46 - 1 > cd_max || get32(s, (size_t)i) != 0x02014b50u) { free(ms); return ZO_ERR_ZIP_CDFH; }
    size_t o = (size_t)i;
    unsigned path_len = get16(s, o + 28), extra = get16(s, o + 30), comment = get16(s, o + 32);
    int64_t next = i + 46 + path_len + extra + comment;
    if (next - 1 > cd_max) { free(ms); return ZO_ERR_ZIP_CDFH; }
    zo_member *m = &ms[k];
    m->path = (const char *)s + o + 46; m->path_len = path_len;
    m->mtime = zo_ptime_of_dos((int)get16(s, o + 14), (int)get16(s, o + 12));
    unsigned hi = get16(s, o + 40);
    if (hi != 0) { m->is_dir = (hi & 070000) == 040000; m->mode = (int)(hi & 07777); }
    else if (s[o + 38] & 0x10) { m->is_dir = 1; m->mode = 0755; }
    else { m->is_dir = 0; m->mode = 0644; }
    if (!m->is_dir) {
      m->compression = (int)get16(s, o + 10);
      m->version_made_by = (int)get16(s, o + 4);
      m->version_needed = (int)get16(s, o + 6);
      m->gp_flags = (int)get16(s, o + 8);
      m->compressed_size = get32(s, o + 20);
      m->decompressed_size = get32(s, o + 24);
      m->crc32 = get32(s, o + 16);
      uint64_t lfh = get32(s, o + 42);
      if (lfh >= len) { free(ms); return ZO_ERR_ZIP_CDFH; }
      /* decode_data_start_of_lfh, zipc.ml:328-336 */
      if (lfh + 30 > len || get32(s, (size_t)lfh) != 0x04034b50u) { free(ms); return ZO_ERR_ZIP_LFH; }
      uint64_t data = lfh + 30 + get16(s, (size_t)lfh + 26) + get16(s, (size_t)lfh + 28);
      if (data + m->compressed_size > len) { free(ms); return ZO_ERR_ZIP_LFH; }
      if (m->crc32 == 0) m->crc32 = get32(s, (size_t)lfh + 14); /* zipc.ml:382-385 */
      m->compressed_bytes = s; m->start = data;
    }
    i = next;
  }
  /* map semantics: sorted by path, later duplicate wins */
  size_t *idx = (size_t *)malloc(sizeof(size_t) * (count ? count : 1));
  size_t m = order_members(ms, count, idx);
  zo_member *res = (zo_member *)calloc(m ? m : 1, sizeof *res);
  for (size_t k = 0; k < m; k++) res[k] = ms[idx[k]];
  free(idx); free(ms);
  *out = res; *out_n = m;
  return 0;
}

/* File.to_binary_string, zipc.ml:205-225.  On ZO_ERR_CHECKSUM *found holds the computed CRC. */
int zo_file_to_binary_string(const zo_member *m, uint8_t **out, size_t *out_len, uint32_t *found) {
  *out = NULL; *out_len = 0;
  if (m->gp_flags & 1) return ZO_ERR_ZIP_ENCRYPTED;
  uint32_t crc = 0;
  if (m->compression == 0) {
    uint8_t *o = (uint8_t *)malloc(m->compressed_size ? m->compressed_size : 1);
    if (!o) return ZO_ERR_NOMEM;
    memcpy(o, m->compressed_bytes + m->start, m->compressed_size);
    crc = zo_crc32(o, m->compressed_size);
    *out = o; *out_len = m->compressed_size;
  } else if (m->compression == 8) {
    int st = zo_inflate(m->compressed_bytes + m->start, m->compressed_size,
                        (int64_t)m->decompressed_size, ZO_CRC_CRC32, out, out_len, &crc);
    if (st) return st;
  } else return ZO_ERR_ZIP_FORMAT;
  if (found) *found = crc;
  if (crc != m->crc32) { free(*out); *out = NULL; *out_len = 0; return ZO_ERR_CHECKSUM; }
  return 0;
}
