// tests/emul/kpc_rt_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host-memory stand-in for kpc_rt.h so that the engine (chunk cutting, epochs, ordering, output) and the
// tile machine of kpc_tile.cuh can be run by the CPU test-suite.  It is linked into tests/emul/_build/ only;
// nothing under kpop_b200/ builds, loads or falls back to it (libkpopcount_gpu.so reports "cuda" and fails
// without a GPU).
#include <cstdlib>
#include <cstring>

#include "../../kpop_b200/csrc/kpc_rt.h"

const char *rt_backend_name() { return "emulation"; }
void rt_init(int) {}
int rt_sm_count() { return 4; }
void rt_set_device(int) {}
int rt_current_device() { return 0; }
// 256-byte aligned like cudaMalloc (the kernels take wider loads on sector-aligned launches)
void *rt_dmalloc(size_t n) {
  const size_t sz = ((n ? n + 64 : 64) + 255) & ~(size_t)255;
  void *p = aligned_alloc(256, sz);
  if (p) memset(p, 0, sz);
  return p;
}
void rt_dfree(void *p) { free(p); }
void *rt_hmalloc(size_t n) { return malloc(n ? n : 1); }
void rt_hfree(void *p) { free(p); }
void rt_h2d(void *d, const void *h, size_t n, rt_stream) { if (n) memcpy(d, h, n); }
void rt_d2h(void *h, const void *d, size_t n, rt_stream) { if (n) memcpy(h, d, n); }
void rt_d2d(void *d, const void *s, size_t n, rt_stream) { if (n) memmove(d, s, n); }
void rt_peer_copy(void *d, int, const void *s, int, size_t n, rt_stream) { if (n) memcpy(d, s, n); }
void rt_memset(void *d, int v, size_t n, rt_stream) { if (n) memset(d, v, n); }
rt_stream rt_stream_create() { return (rt_stream)malloc(1); }
void rt_stream_destroy(rt_stream s) { free(s); }
void rt_stream_sync(rt_stream) {}
void *rt_stream_native(rt_stream s) { return s; }
rt_event rt_event_create() { return (rt_event)malloc(1); }
void rt_event_destroy(rt_event e) { free(e); }
void rt_event_record(rt_event, rt_stream) {}
void rt_stream_wait(rt_stream, rt_event) {}
void rt_event_sync(rt_event) {}
float rt_event_elapsed_ms(rt_event, rt_event) { return 0.f; }
