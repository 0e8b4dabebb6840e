/* TEST INFRASTRUCTURE ONLY: see mlvalues.h */
#ifndef MOCK_CAML_ALLOC_H
#define MOCK_CAML_ALLOC_H
#include "mlvalues.h"
value caml_copy_string(const char *s);
#endif
