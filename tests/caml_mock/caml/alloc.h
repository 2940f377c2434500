#ifndef CAML_MOCK_ALLOC_H
#define CAML_MOCK_ALLOC_H
#include "mlvalues.h"
value caml_alloc(mlsize_t wosize, int tag);
value caml_alloc_tuple(mlsize_t n);
value caml_copy_string(const char *s);
value caml_copy_int32(int32_t i);
value caml_alloc_initialized_string(mlsize_t len, const char *p);
#endif
