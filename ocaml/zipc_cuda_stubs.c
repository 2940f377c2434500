/* zipc_cuda_stubs.c -- OCaml C stubs over libzipc_b200.so (include/zipc_b200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no OCaml toolchain (no caml/ headers) in it.
 * The stubs are deliberately mechanical.  Rules they follow (SURVEY.md 8b "Ownership"):
 *   - OCaml strings may move when the runtime lock is released, so inputs are copied into pinned
 *     staging memory (zipc_b200_host_alloc) BEFORE caml_release_runtime_system();
 *   - results are fetched into malloc'ed memory while released and turned into OCaml strings
 *     after caml_acquire_runtime_system();
 *   - one process-wide context per device, created lazily.
 */
#include <stdlib.h>
#include <string.h>

#include <caml/alloc.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "zipc_b200.h"

static zipc_b200_ctx *g_ctx = NULL;
static int g_device = 0;

static zipc_b200_ctx *ctx(void) {
  if (!g_ctx) {
    int st = zipc_b200_ctx_create(g_device, &g_ctx);
    if (st) caml_failwith(zipc_b200_strerror(st)); /* no CPU fallback */
  }
  return g_ctx;
}

CAMLprim value zipc_cuda_set_device(value d) {
  if (g_ctx) { zipc_b200_ctx_destroy(g_ctx); g_ctx = NULL; }
  g_device = Int_val(d);
  return Val_unit;
}

CAMLprim value zipc_cuda_strerror(value st) { return caml_copy_string(zipc_b200_strerror(Int_val(st))); }

/* Copies n (string, start, len) ranges into one pinned buffer; fills ptr[] / len[]. */
static void *stage(value ss, value starts, value lens, size_t n, const void **ptr, size_t *len) {
  size_t total = 0;
  for (size_t i = 0; i < n; i++) total += (size_t)Long_val(Field(lens, i)) + 16;
  void *buf = NULL;
  if (zipc_b200_host_alloc(total ? total : 1, &buf)) caml_raise_out_of_memory();
  size_t at = 0;
  for (size_t i = 0; i < n; i++) {
    len[i] = (size_t)Long_val(Field(lens, i));
    ptr[i] = (char *)buf + at;
    memcpy((char *)buf + at, String_val(Field(ss, i)) + Long_val(Field(starts, i)), len[i]);
    at += (len[i] + 15) & ~(size_t)15;
  }
  return buf;
}

CAMLprim value zipc_cuda_crc32_batch(value ss, value starts, value lens) {
  CAMLparam3(ss, starts, lens);
  CAMLlocal1(res);
  size_t n = Wosize_val(ss);
  const void **ptr = malloc(sizeof(void *) * (n + 1));
  size_t *len = malloc(sizeof(size_t) * (n + 1));
  uint32_t *crc = malloc(sizeof(uint32_t) * (n + 1));
  void *buf = stage(ss, starts, lens, n, ptr, len);
  zipc_b200_ctx *c = ctx();
  caml_release_runtime_system();
  int st = zipc_b200_crc32_batch(c, n, ptr, len, crc);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (st) { free(ptr); free(len); free(crc); caml_failwith(zipc_b200_strerror(st)); }
  res = caml_alloc(n, 0);
  for (size_t i = 0; i < n; i++) Store_field(res, i, caml_copy_int32((int32_t)crc[i]));
  free(ptr); free(len); free(crc);
  CAMLreturn(res);
}

CAMLprim value zipc_cuda_adler32(value s, value start, value len) {
  CAMLparam3(s, start, len);
  size_t n = (size_t)Long_val(len);
  void *buf = NULL;
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  uint32_t a = 1;
  zipc_b200_ctx *c = ctx();
  caml_release_runtime_system();
  int st = zipc_b200_adler32(c, buf, n, ZIPC_ADLER_REF_COMPAT, &a);
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (st) caml_failwith(zipc_b200_strerror(st));
  CAMLreturn(caml_copy_int32((int32_t)a));
}

/* shared by inflate / deflate: runs `call`, returns [| (status, string, checksum) |] */
typedef int (*batch_fn)(zipc_b200_ctx *, int, int, size_t, const void *const *, const size_t *, const size_t *,
                        void *, size_t, size_t *, size_t *, size_t *, uint32_t *, int *);

static int inflate_call(zipc_b200_ctx *c, int a, int b, size_t n, const void *const *p, const size_t *l,
                        const size_t *mo, void *d, size_t cap, size_t *need, size_t *off, size_t *len,
                        uint32_t *ck, int *st) {
  (void)b;
  return zipc_b200_inflate_batch(c, a, ZIPC_ADLER_REF_COMPAT, n, p, l, mo, d, cap, need, off, len, ck, st);
}
static int deflate_call(zipc_b200_ctx *c, int a, int level, size_t n, const void *const *p, const size_t *l,
                        const size_t *mo, void *d, size_t cap, size_t *need, size_t *off, size_t *len,
                        uint32_t *ck, int *st) {
  (void)mo;
  return zipc_b200_deflate_batch(c, level, a, ZIPC_ADLER_REF_COMPAT, n, p, l, d, cap, need, off, len, ck, st);
}

static value run_batch(batch_fn call, int a, int b, value ss, value starts, value lens, value dsz) {
  CAMLparam4(ss, starts, lens, dsz);
  CAMLlocal3(res, tup, str);
  size_t n = Wosize_val(ss);
  const void **ptr = malloc(sizeof(void *) * (n + 1));
  size_t *len = malloc(sizeof(size_t) * (n + 1)), *mo = malloc(sizeof(size_t) * (n + 1));
  size_t *off = malloc(sizeof(size_t) * (n + 1)), *ol = malloc(sizeof(size_t) * (n + 1));
  uint32_t *ck = malloc(sizeof(uint32_t) * (n + 1));
  int *st = malloc(sizeof(int) * (n + 1));
  for (size_t i = 0; i < n; i++)
    mo[i] = (dsz == Val_unit || Long_val(Field(dsz, i)) < 0) ? ZIPC_SIZE_UNKNOWN : (size_t)Long_val(Field(dsz, i));
  void *buf = stage(ss, starts, lens, n, ptr, len);
  zipc_b200_ctx *c = ctx();
  size_t need = 0;
  void *out = NULL;
  caml_release_runtime_system();
  int rc = call(c, a, b, n, ptr, len, mo, NULL, 0, &need, off, ol, ck, st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(need ? need : 1); rc = zipc_b200_fetch(c, out, need); }
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  res = caml_alloc(n, 0);
  for (size_t i = 0; i < n; i++) {
    str = caml_alloc_initialized_string(st[i] ? 0 : ol[i], st[i] ? "" : (char *)out + off[i]);
    tup = caml_alloc_tuple(3);
    Store_field(tup, 0, Val_int(st[i]));
    Store_field(tup, 1, str);
    Store_field(tup, 2, caml_copy_int32((int32_t)ck[i]));
    Store_field(res, i, tup);
  }
  free(out); free(ptr); free(len); free(mo); free(off); free(ol); free(ck); free(st);
  CAMLreturn(res);
}

CAMLprim value zipc_cuda_inflate_batch(value crc_op, value ss, value starts, value lens, value dsz) {
  return run_batch(inflate_call, Int_val(crc_op), 0, ss, starts, lens, dsz);
}
CAMLprim value zipc_cuda_deflate_batch(value level, value crc_op, value ss, value starts, value lens) {
  return run_batch(deflate_call, Int_val(crc_op), Int_val(level), ss, starts, lens, Val_unit);
}

CAMLprim value zipc_cuda_zlib_decompress(value s, value start, value len, value dsz) {
  CAMLparam4(s, start, len, dsz);
  CAMLlocal2(tup, str);
  size_t n = (size_t)Long_val(len), mo = Long_val(dsz) < 0 ? ZIPC_SIZE_UNKNOWN : (size_t)Long_val(dsz);
  void *buf = NULL, *out = NULL;
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  const void *p = buf;
  size_t need = 0, off = 0, ol = 0;
  uint32_t expect = 0, found = 0;
  int st = 0;
  zipc_b200_ctx *c = ctx();
  caml_release_runtime_system();
  int rc = zipc_b200_zlib_decompress_batch(c, ZIPC_ADLER_REF_COMPAT, 1, &p, &n, &mo, NULL, 0, &need, &off, &ol,
                                           &expect, &found, &st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(need ? need : 1); rc = zipc_b200_fetch(c, out, need); }
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  str = caml_alloc_initialized_string(st ? 0 : ol, st ? "" : (char *)out + off);
  tup = caml_alloc_tuple(4);
  Store_field(tup, 0, Val_int(st));
  Store_field(tup, 1, str);
  Store_field(tup, 2, caml_copy_int32((int32_t)expect));
  Store_field(tup, 3, caml_copy_int32((int32_t)found));
  free(out);
  CAMLreturn(tup);
}

CAMLprim value zipc_cuda_zlib_compress(value level, value s, value start, value len) {
  CAMLparam4(level, s, start, len);
  CAMLlocal2(tup, str);
  size_t n = (size_t)Long_val(len);
  void *buf = NULL, *out = NULL;
  if (zipc_b200_host_alloc(n ? n : 1, &buf)) caml_raise_out_of_memory();
  memcpy(buf, String_val(s) + Long_val(start), n);
  const void *p = buf;
  size_t need = 0, off = 0, ol = 0;
  uint32_t adler = 1;
  int st = 0;
  zipc_b200_ctx *c = ctx();
  caml_release_runtime_system();
  int rc = zipc_b200_zlib_compress_batch(c, Int_val(level), ZIPC_ADLER_REF_COMPAT, 1, &p, &n, NULL, 0, &need, &off,
                                         &ol, &adler, &st);
  if (rc == ZIPC_ERR_DST_TOO_SMALL) { out = malloc(need ? need : 1); rc = zipc_b200_fetch(c, out, need); }
  caml_acquire_runtime_system();
  zipc_b200_host_free(buf);
  if (rc) { free(out); caml_failwith(zipc_b200_strerror(rc)); }
  str = caml_alloc_initialized_string(st ? 0 : ol, st ? "" : (char *)out + off);
  tup = caml_alloc_tuple(3);
  Store_field(tup, 0, Val_int(st));
  Store_field(tup, 1, str);
  Store_field(tup, 2, caml_copy_int32((int32_t)adler));
  free(out);
  CAMLreturn(tup);
}
