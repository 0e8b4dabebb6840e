/* TEST INFRASTRUCTURE ONLY: see mlvalues.h */
#ifndef MOCK_CAML_BIGARRAY_H
#define MOCK_CAML_BIGARRAY_H
#include "mlvalues.h"
struct caml_ba_array {
  void *data;
  intnat num_dims;
  intnat flags;
  void *proxy;
  intnat dim[1];
};
enum { CAML_BA_UINT8 = 3, CAML_BA_C_LAYOUT = 0, CAML_BA_EXTERNAL = 0x200 };
#define Caml_ba_array_val(v) ((struct caml_ba_array *)Data_custom_val(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
value caml_ba_alloc(int flags, int num_dims, void *data, intnat *dim);
#endif
