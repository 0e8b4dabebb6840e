/* TEST INFRASTRUCTURE ONLY: see mlvalues.h */
#ifndef MOCK_CAML_MEMORY_H
#define MOCK_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam0() int caml__frame = 0
#define CAMLparam1(a) int caml__frame = 0; (void)(a)
#define CAMLparam2(a, b) CAMLparam1(a); (void)(b)
#define CAMLparam3(a, b, c) CAMLparam2(a, b); (void)(c)
#define CAMLparam4(a, b, c, d) CAMLparam3(a, b, c); (void)(d)
#define CAMLparam5(a, b, c, d, e) CAMLparam4(a, b, c, d); (void)(e)
#define CAMLlocal1(x) value x = Val_unit
#define CAMLreturn(x) do { (void)caml__frame; return (x); } while (0)
#endif
