/* TEST INFRASTRUCTURE ONLY: see mlvalues.h */
#ifndef MOCK_CAML_FAIL_H
#define MOCK_CAML_FAIL_H
#include "mlvalues.h"
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#endif
