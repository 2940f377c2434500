/* zipc_cuda_stubs.c -- OCaml C stubs over libzipc_b200.so (include/zipc_b200.h).
 *
 * NOT COMPILED AGAINST A REAL OCAML IN THIS REPOSITORY'S IMAGE: there is no OCaml toolchain in it.  What IS checked
 * mechanically (tests/test_ocaml_binding.py): the file compiles with gcc against a minimal mock of the caml/
 * headers, so every zipc_b200_* call matches its prototype in include/zipc_b200.h, and every `external` of
 * zipc_cuda.ml names a stub defined here with the same arity.
 *
 * Rules the stubs follow (SURVEY.md 8b "Ownership"):
 *   - OCaml strings may move when the runtime lock is released, so inputs are copied into pinned staging memory
 *     (zipc_b200_host_alloc) BEFORE caml_release_runtime_system();
 *   - results are fetched into malloc'ed memory while released and turned into OCaml strings after
 *     caml_acquire_runtime_system();
 *   - one process-wide context (one device) or multi-context (device mask), created lazily BEFORE any temporary
 *     is allocated; a mutex held across call + fetch keeps a second OCaml thread off the context while the
 *     runtime lock is released (a ctx has one stream and one set of arenas);
 *   - every temporary is released on every exit path before an exception is raised.
 */
#define _POSIX_C_SOURCE 200809L /* strdup */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "zipc_b200.h"

static zipc_b200_ctx *g_ctx = NULL;   /* single-device mode */
static zipc_b200_mctx *g_mctx = NULL; /* box-wide mode (set_devices) */
static int g_device = 0;
static int g_multi = 0;
static unsigned long long g_mask = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static void drop_contexts(void) {
  if (g_ctx) { zipc_b200_ctx_destroy(g_ctx); g_ctx = NULL; }
  if (g_mctx) { zipc_b200_mctx_destroy(g_mctx); g_mctx = NULL; }
}

/* Creates the context(s) on first use.  Called with the runtime lock held and nothing allocated yet. */
static void ensure_ctx(void) {
  int st = 0;
  pthread_mutex_lock(&g_lock);
  if (g_multi) { if (!g_mctx) st = zipc_b200_mctx_create(g_mask, &g_mctx); }
  else if (!g_ctx) st = zipc_b200_ctx_create(g_device, &g_ctx);
  pthread_mutex_unlock(&g_lock);
  if (st) caml_failwith(zipc_b200_strerror(st)); /* no CPU fallback */
}

/* the single-device context a non-batch call runs on */
static zipc_b200_ctx *one_ctx(void) { return g_multi ? zipc_b200_mctx_ctx(g_mctx, 0) : g_ctx; }

CAMLprim value zipc_cuda_set_device(value d) {
  pthread_mutex_lock(&g_lock);
  drop_contexts();
  g_multi = 0;
  g_device = Int_val(d);
  pthread_mutex_unlock(&g_lock);
  return Val_unit;
}

CAMLprim value zipc_cuda_set_devices(value mask) {
  pthread_mutex_lock(&g_lock);
  drop_contexts();
  g_multi = 1;
  g_mask = (unsigned long long)Long_val(mask);
  pthread_mutex_unlock(&g_lock);
  return Val_unit;
}

CAMLprim value zipc_cuda_strerror(value st) { return caml_copy_string(zipc_b200_strerror(Int_val(st))); }

/* ---- temporaries of one batch call: one allocation, one release -------------------------------------------- */
typedef struct {
  size_t n;
  void *pinned;          /* staged inputs */
  const void **ptr;
  size_t *len, *mo, *off, *ol;
  uint32_t *ck;
  int *st;
  void *out;             /* fetched results */
} batch_tmp;

static void tmp_free(batch_tmp *t) {
  if (t->pinned) zipc_b200_host_free(t->pinned);
  free(t->ptr); free(t->len); free(t->mo); free(t->off); free(t->ol); free(t->ck); free(t->st); free(t->out);
  memset(t, 0, sizeof *t);
}

/* Allocates the arrays and copies n (string, start, len) ranges into one pinned buffer.  Returns 0 on failure
 * (everything released). */
static int tmp_stage(batch_tmp *t, value ss, value starts, value lens) {
  size_t n = Wosize_val(ss), total = 0, at = 0;
  memset(t, 0, sizeof *t);
  t->n = n;
  t->ptr = malloc(sizeof(void *) * (n + 1));
  t->len = malloc(sizeof(size_t) * (n + 1)); t->mo = malloc(sizeof(size_t) * (n + 1));
  t->off = malloc(sizeof(size_t) * (n + 1)); t->ol = malloc(sizeof(size_t) * (n + 1));
  t->ck = malloc(sizeof(uint32_t) * (n + 1));
  t->st = malloc(sizeof(int) * (n + 1));
  for (size_t i = 0; i < n; i++) total += (size_t)Long_val(Field(lens, i)) + 16;
  if (!t->ptr || !t->len || !t->mo || !t->off || !t->ol || !t->ck || !t->st ||
      zipc_b200_host_alloc(total ? total : 1, &t->pinned)) { tmp_free(t); return 0; }
  for (size_t i = 0; i < n; i++) {
    t->len[i] = (size_t)Long_val(Field(lens, i));
    t->ptr[i] = (char *)t->pinned + at;
    memcpy((char *)t->pinned + at, String_val(Field(ss, i)) + Long_val(Field(starts, i)), t->len[i]);
    at += (t->len[i] + 15) & ~(size_t)15;
    t->mo[i] = ZIPC_SIZE_UNKNOWN; t->ck[i] = 0; t->st[i] = 0; t->off[i] = 0; t->ol[i] = 0;
  }
  return 1;
}

CAMLprim value zipc_cuda_crc32_batch(value ss, value starts, value lens) {
  CAMLparam3(ss, starts, lens);
  CAMLlocal1(res);
  batch_tmp t;
  ensure_ctx();
  if (!tmp_stage(&t, ss, starts, lens)) caml_raise_out_of_memory();
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_crc32_batch(one_ctx(), t.n, t.ptr, t.len, t.ck);
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  if (rc) { tmp_free(&t); caml_failwith(zipc_b200_strerror(rc)); }
  res = caml_alloc(t.n, 0);
  for (size_t i = 0; i < t.n; i++) Store_field(res, i, caml_copy_int32((int32_t)t.ck[i]));
  tmp_free(&t);
  CAMLreturn(res);
}

CAMLprim value zipc_cuda_adler32(value s, value start, value len) {
  CAMLparam3(s, start, len);
  size_t n = (size_t)Long_val(len);
  void *buf = NULL;
  uint32_t a = 1;
  ensure_ctx();
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_adler32(one_ctx(), buf, n, ZIPC_ADLER_REF_COMPAT, &a);
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) caml_failwith(zipc_b200_strerror(rc));
  CAMLreturn(caml_copy_int32((int32_t)a));
}

/* shared by inflate / deflate: runs the call without an arena, fetches, returns [| (status, string, checksum) |] */
static value run_batch(int inflate, int ck_kind, int level, value ss, value starts, value lens, value dsz) {
  CAMLparam4(ss, starts, lens, dsz);
  CAMLlocal3(res, tup, str);
  batch_tmp t;
  size_t need = 0;
  int rc;
  ensure_ctx();
  if (!tmp_stage(&t, ss, starts, lens)) caml_raise_out_of_memory();
  if (dsz != Val_unit)
    for (size_t i = 0; i < t.n; i++)
      if (Long_val(Field(dsz, i)) >= 0) t.mo[i] = (size_t)Long_val(Field(dsz, i));
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  if (g_multi)
    rc = inflate ? zipc_b200_multi_inflate_batch(g_mctx, ck_kind, ZIPC_ADLER_REF_COMPAT, t.n, t.ptr, t.len, t.mo, NULL, 0, &need, t.off, t.ol, t.ck, t.st)
                 : zipc_b200_multi_deflate_batch(g_mctx, level, ck_kind, ZIPC_ADLER_REF_COMPAT, t.n, t.ptr, t.len, NULL, 0, &need, t.off, t.ol, t.ck, t.st);
  else
    rc = inflate ? zipc_b200_inflate_batch(g_ctx, ck_kind, ZIPC_ADLER_REF_COMPAT, t.n, t.ptr, t.len, t.mo, NULL, 0, &need, t.off, t.ol, t.ck, t.st)
                 : zipc_b200_deflate_batch(g_ctx, level, ck_kind, ZIPC_ADLER_REF_COMPAT, t.n, t.ptr, t.len, NULL, 0, &need, t.off, t.ol, t.ck, t.st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) {
    t.out = malloc(need ? need : 1);
    rc = !t.out ? ZIPC_ERR_NOMEM : g_multi ? zipc_b200_multi_fetch(g_mctx, t.out, need) : zipc_b200_fetch(g_ctx, t.out, need);
  }
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  if (rc) { tmp_free(&t); caml_failwith(zipc_b200_strerror(rc)); }
  res = caml_alloc(t.n, 0);
  for (size_t i = 0; i < t.n; i++) {
    str = caml_alloc_initialized_string(t.st[i] ? 0 : t.ol[i], t.st[i] ? "" : (char *)t.out + t.off[i]);
    tup = caml_alloc_tuple(3);
    Store_field(tup, 0, Val_int(t.st[i]));
    Store_field(tup, 1, str);
    Store_field(tup, 2, caml_copy_int32((int32_t)t.ck[i]));
    Store_field(res, i, tup);
  }
  tmp_free(&t);
  CAMLreturn(res);
}

CAMLprim value zipc_cuda_inflate_batch(value crc_op, value ss, value starts, value lens, value dsz) {
  return run_batch(1, Int_val(crc_op), 0, ss, starts, lens, dsz);
}
CAMLprim value zipc_cuda_deflate_batch(value level, value crc_op, value ss, value starts, value lens) {
  return run_batch(0, Int_val(crc_op), Int_val(level), ss, starts, lens, Val_unit);
}

CAMLprim value zipc_cuda_zlib_decompress(value s, value start, value len, value dsz) {
  CAMLparam4(s, start, len, dsz);
  CAMLlocal2(tup, str);
  size_t n = (size_t)Long_val(len), mo = Long_val(dsz) < 0 ? ZIPC_SIZE_UNKNOWN : (size_t)Long_val(dsz);
  void *buf = NULL, *out = NULL;
  size_t need = 0, off = 0, ol = 0;
  uint32_t expect = 0, found = 0;
  int st = 0;
  ensure_ctx();
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  const void *p = buf;
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_zlib_decompress_batch(one_ctx(), ZIPC_ADLER_REF_COMPAT, 1, &p, &n, &mo, NULL, 0, &need, &off, &ol,
                                           &expect, &found, &st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(need ? need : 1); rc = out ? zipc_b200_fetch(one_ctx(), out, need) : ZIPC_ERR_NOMEM; }
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  str = caml_alloc_initialized_string(st ? 0 : ol, st ? "" : (char *)out + off);
  free(out);
  tup = caml_alloc_tuple(4);
  Store_field(tup, 0, Val_int(st));
  Store_field(tup, 1, str);
  Store_field(tup, 2, caml_copy_int32((int32_t)expect));
  Store_field(tup, 3, caml_copy_int32((int32_t)found));
  CAMLreturn(tup);
}

CAMLprim value zipc_cuda_zlib_compress(value level, value s, value start, value len) {
  CAMLparam4(level, s, start, len);
  CAMLlocal2(tup, str);
  size_t n = (size_t)Long_val(len);
  void *buf = NULL, *out = NULL;
  size_t need = 0, off = 0, ol = 0;
  uint32_t adler = 1;
  int st = 0;
  ensure_ctx();
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  const void *p = buf;
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_zlib_compress_batch(one_ctx(), Int_val(level), ZIPC_ADLER_REF_COMPAT, 1, &p, &n, NULL, 0, &need, &off,
                                         &ol, &adler, &st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(need ? need : 1); rc = out ? zipc_b200_fetch(one_ctx(), out, need) : ZIPC_ERR_NOMEM; }
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  str = caml_alloc_initialized_string(st ? 0 : ol, st ? "" : (char *)out + off);
  free(out);
  tup = caml_alloc_tuple(3);
  Store_field(tup, 0, Val_int(st));
  Store_field(tup, 1, str);
  Store_field(tup, 2, caml_copy_int32((int32_t)adler));
  CAMLreturn(tup);
}

/* ---- one large stream as independent segments ------------------------------------------------------------------ */
/* -> (status, stream, flat index [| c0; u0; c1; u1; ...; ctotal; utotal |], crc32 of the input) */
CAMLprim value zipc_cuda_deflate_segmented(value level, value s, value segment_size) {
  CAMLparam3(level, s, segment_size);
  CAMLlocal3(tup, str, idx);
  const size_t n = caml_string_length(s), seg = (size_t)Long_val(segment_size);
  if (seg < 4096) caml_invalid_argument("segment_size");
  const size_t nmax = (n ? (n + seg - 1) / seg : 1) + 1;
  void *buf = NULL, *out = NULL;
  uint64_t *index = NULL;
  size_t olen = 0, nseg = 0;
  uint32_t crc = 0;
  ensure_ctx();
  index = malloc(sizeof(uint64_t) * 2 * (nmax + 1));
  if (!index || zipc_b200_host_alloc(n ? n : 1, &buf)) { free(index); caml_raise_out_of_memory(); }
  memcpy(buf, String_val(s), n);
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_deflate_segmented(one_ctx(), Int_val(level), buf, n, seg, 1, NULL, 0, &olen, index, nmax + 1, &nseg, &crc);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(olen ? olen : 1); rc = out ? zipc_b200_fetch(one_ctx(), out, olen) : ZIPC_ERR_NOMEM; }
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc == ZIPC_ERR_CUDA || rc == ZIPC_ERR_NOMEM || rc == ZIPC_ERR_NO_DEVICE || rc == ZIPC_ERR_INVALID_ARG) {
    free(out); free(index);
    caml_failwith(zipc_b200_strerror(rc));
  }
  str = caml_alloc_initialized_string(rc ? 0 : olen, rc ? "" : (char *)out);
  free(out);
  idx = caml_alloc(rc ? 0 : 2 * (nseg + 1), 0);
  if (!rc) for (size_t i = 0; i < 2 * (nseg + 1); i++) Store_field(idx, i, Val_long((long)index[i]));
  free(index);
  tup = caml_alloc_tuple(4);
  Store_field(tup, 0, Val_int(rc));
  Store_field(tup, 1, str);
  Store_field(tup, 2, idx);
  Store_field(tup, 3, caml_copy_int32((int32_t)crc));
  CAMLreturn(tup);
}

/* flat index as produced above -> (status, output, crc32 of the output) */
CAMLprim value zipc_cuda_inflate_segmented(value s, value flat) {
  CAMLparam2(s, flat);
  CAMLlocal2(tup, str);
  const size_t n = caml_string_length(s), pairs = Wosize_val(flat) / 2;
  if (pairs < 2) caml_invalid_argument("index");
  const size_t nseg = pairs - 1;
  void *buf = NULL, *out = NULL;
  uint64_t *index = NULL;
  size_t olen = 0;
  uint32_t crc = 0;
  int st = 0;
  ensure_ctx();
  index = malloc(sizeof(uint64_t) * 2 * pairs);
  if (!index || zipc_b200_host_alloc(n ? n : 1, &buf)) { free(index); caml_raise_out_of_memory(); }
  for (size_t i = 0; i < 2 * pairs; i++) index[i] = (uint64_t)Long_val(Field(flat, i));
  memcpy(buf, String_val(s), n);
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_inflate_segmented(one_ctx(), buf, n, index, nseg, NULL, 0, &olen, &crc, &st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(olen ? olen : 1); rc = out ? zipc_b200_fetch(one_ctx(), out, olen) : ZIPC_ERR_NOMEM; }
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  free(index);
  if (rc == ZIPC_ERR_INVALID_ARG) { free(out); caml_invalid_argument("index"); }
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  str = caml_alloc_initialized_string(st ? 0 : olen, st ? "" : (char *)out);
  free(out);
  tup = caml_alloc_tuple(3);
  Store_field(tup, 0, Val_int(st));
  Store_field(tup, 1, str);
  Store_field(tup, 2, caml_copy_int32((int32_t)crc));
  CAMLreturn(tup);
}

/* ---- archive: File.deflate_of_binary_string x n + Zipc.to_binary_string ------------------------------------------ */
/* -> (status, archive bytes) */
CAMLprim value zipc_cuda_archive(value level, value paths, value payloads, value first) {
  CAMLparam4(level, paths, payloads, first);
  CAMLlocal2(tup, str);
  const size_t n = Wosize_val(paths);
  if (Wosize_val(payloads) != n) caml_invalid_argument("archive_to_binary_string");
  size_t total = 0, ptotal = 0, at = 0, pat = 0, bound = 22;
  for (size_t i = 0; i < n; i++) {
    const size_t l = caml_string_length(Field(payloads, i)), pl = caml_string_length(Field(paths, i));
    total += l + 16; ptotal += pl + 1;
    bound += zipc_b200_deflate_bound(l) + 2 * pl + 30 + 46;
  }
  void *buf = NULL, *out = NULL;
  char *pbuf = NULL, *firstc = NULL;
  const char **pp = NULL;
  const void **sp = NULL;
  uint32_t *plen = NULL;
  size_t *slen = NULL, olen = 0;
  ensure_ctx();
  pbuf = malloc(ptotal + 1); pp = malloc(sizeof(char *) * (n + 1)); sp = malloc(sizeof(void *) * (n + 1));
  plen = malloc(sizeof(uint32_t) * (n + 1)); slen = malloc(sizeof(size_t) * (n + 1));
  out = malloc(bound);
  if (caml_string_length(first)) firstc = strdup(String_val(first));
  if (!pbuf || !pp || !sp || !plen || !slen || !out || zipc_b200_host_alloc(total ? total : 1, &buf)) {
    free(pbuf); free(pp); free(sp); free(plen); free(slen); free(out); free(firstc);
    caml_raise_out_of_memory();
  }
  for (size_t i = 0; i < n; i++) {
    slen[i] = caml_string_length(Field(payloads, i));
    plen[i] = (uint32_t)caml_string_length(Field(paths, i));
    sp[i] = (char *)buf + at; pp[i] = pbuf + pat;
    memcpy((char *)buf + at, String_val(Field(payloads, i)), slen[i]);
    memcpy(pbuf + pat, String_val(Field(paths, i)), plen[i]);
    at += (slen[i] + 15) & ~(size_t)15; pat += plen[i] + 1;
  }
  caml_release_runtime_system();
  pthread_mutex_lock(&g_lock);
  int rc = zipc_b200_zip_deflate_archive(one_ctx(), Int_val(level), n, pp, plen, sp, slen, NULL, NULL, firstc, out, bound, &olen);
  pthread_mutex_unlock(&g_lock);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  free(pbuf); free(pp); free(sp); free(plen); free(slen); free(firstc);
  if (rc == ZIPC_ERR_CUDA || rc == ZIPC_ERR_NOMEM || rc == ZIPC_ERR_NO_DEVICE || rc == ZIPC_ERR_INVALID_ARG) {
    free(out);
    caml_failwith(zipc_b200_strerror(rc));
  }
  str = caml_alloc_initialized_string(rc ? 0 : olen, rc ? "" : (char *)out);
  free(out);
  tup = caml_alloc_tuple(2);
  Store_field(tup, 0, Val_int(rc));
  Store_field(tup, 1, str);
  CAMLreturn(tup);
}
