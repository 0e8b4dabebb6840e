/* kpc_stubs.c -- C glue between OCaml (`external` declarations of Kpc_gpu.ml) and the C ABI of libkpopcount_gpu.so
 * (include/kpopcount.h).  One stub per `external`.  The reference (PaoloRibeca/KPop) has no FFI: these stubs are what
 * a maintainer adds next to bin/KPopCount.ml, whose KMerCounter.compute (bin/KPopCount.ml:26-63) then becomes the
 * short function in KPopCount_gpu.ml.
 *
 * Conventions: a negative return of the library becomes Failure (kpc_error ctx) -- or Invalid_argument for
 * KPC_E_ARG -- after the runtime lock has been re-acquired; blocking calls (feed, end, finish) release the OCaml
 * runtime lock; the context lives in a custom block finalised with kpc_destroy.
 * tests/test_ocaml_stubs.py compiles this file against a mock of <caml/...> (there is no OCaml in the image). */
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <caml/threads.h>

#include "kpopcount.h"

#define Ctx_val(v) (*((kpc_ctx **)Data_custom_val(v)))

static void ctx_finalize(value v) {
  if (Ctx_val(v)) kpc_destroy(Ctx_val(v));
  Ctx_val(v) = NULL;
}
static struct custom_operations ctx_ops = {"kpop.kpc_ctx",          ctx_finalize,
                                           custom_compare_default,  custom_hash_default,
                                           custom_serialize_default, custom_deserialize_default,
                                           custom_compare_ext_default, custom_fixed_length_default};

/* rc < 0: raise with the library's message (copied first: raising does not return) */
static void check(kpc_ctx *c, int rc) {
  if (rc == KPC_OK) return;
  if (rc == KPC_E_ARG) caml_invalid_argument(kpc_error(c));
  caml_failwith(kpc_error(c));
}

/* kpc_sink_fn: the spectra text, in final order, goes to a file descriptor */
static int fd_sink(void *user, const char *b, size_t n) {
  int fd = (int)(intptr_t)user;
  while (n) {
    ssize_t w = write(fd, b, n);
    if (w <= 0) return 1;
    b += w;
    n -= (size_t)w;
  }
  return 0;
}

value kpc_ml_create(value k, value content, value m, value label, value devices) {
  CAMLparam5(k, content, m, label, devices);
  CAMLlocal1(v);
  kpc_ctx *c = NULL;
  int ids[64];
  int n = (int)Wosize_val(devices);
  if (n < 1 || n > 64) caml_invalid_argument("Kpc_gpu.create: between 1 and 64 devices");
  for (int i = 0; i < n; ++i) ids[i] = Int_val(Field(devices, i));
  int rc = kpc_create(&c, Int_val(k), Int_val(content), (long long)Long_val(m), String_val(label), n, ids);
  if (rc != KPC_OK) {
    char msg[512];
    strncpy(msg, kpc_error(c), sizeof msg - 1);
    msg[sizeof msg - 1] = 0;
    if (c) kpc_destroy(c);
    if (rc == KPC_E_ARG) caml_invalid_argument(msg);
    caml_failwith(msg); /* e.g. "Invalid argument (k must be <= 30, found 31)", KMers.ml:264-267 */
  }
  v = caml_alloc_custom(&ctx_ops, sizeof(kpc_ctx *), 0, 1);
  Ctx_val(v) = c;
  CAMLreturn(v);
}

value kpc_ml_set_out(value c, value fd) {
  CAMLparam2(c, fd);
  check(Ctx_val(c), kpc_set_sink(Ctx_val(c), fd_sink, (void *)(intptr_t)Int_val(fd)));
  CAMLreturn(Val_unit);
}

value kpc_ml_staging_slots(value c) {
  CAMLparam1(c);
  CAMLreturn(Val_int(kpc_staging_slots(Ctx_val(c))));
}

value kpc_ml_staging(value c, value slot) {
  CAMLparam2(c, slot);
  size_t cap = 0;
  kpc_ctx *x = Ctx_val(c);
  int s = Int_val(slot);
  caml_release_runtime_system(); /* waits until the previous feed from this slot has left the buffer */
  void *p = kpc_staging(x, s, &cap);
  caml_acquire_runtime_system();
  if (!p) caml_failwith(kpc_error(x));
  intnat dim = (intnat)cap;
  CAMLreturn(caml_ba_alloc(CAML_BA_UINT8 | CAML_BA_C_LAYOUT | CAML_BA_EXTERNAL, 1, p, &dim));
}

value kpc_ml_begin(value c, value format) {
  CAMLparam2(c, format);
  check(Ctx_val(c), kpc_begin(Ctx_val(c), Int_val(format)));
  CAMLreturn(Val_unit);
}

value kpc_ml_feed(value c, value mate, value ba, value len, value eof) {
  CAMLparam5(c, mate, ba, len, eof);
  kpc_ctx *x = Ctx_val(c);
  void *p = Caml_ba_data_val(ba);
  size_t n = (size_t)Long_val(len);
  int m = Int_val(mate), e = Bool_val(eof);
  if ((intnat)n > Caml_ba_array_val(ba)->dim[0]) caml_invalid_argument("Kpc_gpu.feed: len exceeds the buffer");
  caml_release_runtime_system(); /* -L / spill dumps are written from inside (bin/KPopCount.ml:39-50) */
  int rc = kpc_feed(x, m, p, n, e);
  caml_acquire_runtime_system();
  check(x, rc);
  CAMLreturn(Val_unit);
}

value kpc_ml_feed_bytes(value c, value mate, value bytes, value len, value eof) {
  CAMLparam5(c, mate, bytes, len, eof);
  size_t n = (size_t)Long_val(len);
  if (n > caml_string_length(bytes)) caml_invalid_argument("Kpc_gpu.feed_bytes: len exceeds the buffer");
  /* the OCaml heap may move: no runtime release here; the library copies before it returns */
  check(Ctx_val(c), kpc_feed(Ctx_val(c), Int_val(mate), Bytes_val(bytes), n, Bool_val(eof)));
  CAMLreturn(Val_unit);
}

value kpc_ml_end(value c) {
  CAMLparam1(c);
  kpc_ctx *x = Ctx_val(c);
  caml_release_runtime_system();
  int rc = kpc_end(x);
  caml_acquire_runtime_system();
  check(x, rc);
  CAMLreturn(Val_unit);
}

value kpc_ml_finish(value c) {
  CAMLparam1(c);
  kpc_ctx *x = Ctx_val(c);
  caml_release_runtime_system();
  int rc = kpc_finish(x);
  caml_acquire_runtime_system();
  check(x, rc);
  CAMLreturn(Val_unit);
}

value kpc_ml_set_pair_limit(value c, value n) {
  CAMLparam2(c, n);
  check(Ctx_val(c), kpc_set_pair_limit(Ctx_val(c), (long long)Long_val(n)));
  CAMLreturn(Val_unit);
}

value kpc_ml_set_single_pass(value c, value on) {
  CAMLparam2(c, on);
  check(Ctx_val(c), kpc_set_single_pass(Ctx_val(c), Bool_val(on)));
  CAMLreturn(Val_unit);
}

value kpc_ml_complete_pairs(value c) {
  CAMLparam1(c);
  CAMLreturn(Val_long(kpc_complete_pairs(Ctx_val(c))));
}

value kpc_ml_kmers_counted(value c) {
  CAMLparam1(c);
  unsigned long long n = 0;
  check(Ctx_val(c), kpc_kmers_counted(Ctx_val(c), &n));
  CAMLreturn(Val_long((intnat)n));
}

value kpc_ml_reset(value c) {
  CAMLparam1(c);
  check(Ctx_val(c), kpc_reset(Ctx_val(c)));
  CAMLreturn(Val_unit);
}

value kpc_ml_backend(value unit) {
  CAMLparam1(unit);
  CAMLreturn(caml_copy_string(kpc_backend()));
}
