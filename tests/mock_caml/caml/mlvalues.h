/* tests/mock_caml/caml/mlvalues.h -- TEST INFRASTRUCTURE ONLY.
 * A minimal stand-in for the OCaml runtime headers (there is no OCaml toolchain in the image) with the documented
 * signatures of the macros and functions ocaml/kpc_stubs.c uses, so that gcc can type-check the stubs.  Nothing links
 * against it; it has no behaviour. */
#ifndef MOCK_CAML_MLVALUES_H
#define MOCK_CAML_MLVALUES_H
#include <stddef.h>
#include <stdint.h>
typedef intptr_t intnat;
typedef uintptr_t uintnat;
typedef intnat value;
typedef uintnat mlsize_t;
#define Val_long(x) ((intnat)(((uintnat)(x) << 1)) + 1)
#define Long_val(x) ((x) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(x) ((int)Long_val(x))
#define Val_unit Val_int(0)
#define Val_bool(x) Val_int((x) != 0)
#define Bool_val(x) Int_val(x)
#define Field(x, i) (((value *)(x))[i])
#define Wosize_val(v) ((mlsize_t)(((uintnat *)(v))[-1] >> 10))
#define String_val(x) ((const char *)(x))
#define Bytes_val(x) ((unsigned char *)(x))
mlsize_t caml_string_length(value s);
#endif
