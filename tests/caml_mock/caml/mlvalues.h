/* Minimal MOCK of the OCaml runtime headers -- TEST INFRASTRUCTURE ONLY (tests/test_ocaml_binding.py).
 * It exists so that gcc can type-check ocaml/zipc_cuda_stubs.c against include/zipc_b200.h in an image without an
 * OCaml toolchain.  Declarations only; nothing here is linked or run. */
#ifndef CAML_MOCK_MLVALUES_H
#define CAML_MOCK_MLVALUES_H
#include <stddef.h>
#include <stdint.h>
typedef intptr_t value;
typedef size_t mlsize_t;
#define CAMLprim
#define Val_unit ((value)1)
#define Val_int(x) ((value)(((intptr_t)(x) << 1) + 1))
#define Val_long(x) Val_int(x)
#define Int_val(v) ((int)((v) >> 1))
#define Long_val(v) ((long)((v) >> 1))
#define Field(v, i) (((value *)(v))[i])
#define String_val(v) ((const char *)(v))
mlsize_t Wosize_val(value v);
mlsize_t caml_string_length(value v);
void Store_field(value block, mlsize_t i, value v);
#endif
