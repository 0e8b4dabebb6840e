// kpc_rt.h -- the handful of runtime services the host engine needs (device/pinned memory, streams,
// events, copies).  kpc_rt_cuda.cu implements them with the CUDA runtime and is what ships in
// libkpopcount_gpu.so.  tests/emul/ holds a host-memory implementation that exists only so that the
// engine and the tile machine can be exercised by the CPU test-suite; it is never built into the product.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

struct KpcError {
  int code;
  std::string msg;
  KpcError(int c, const std::string &m) : code(c), msg(m) {}
};

typedef struct rt_stream_s *rt_stream;
typedef struct rt_event_s *rt_event;

const char *rt_backend_name();          // "cuda" for the product
void rt_init(int device);               // selects the device, throws KpcError(KPC_E_CUDA) if there is none
int rt_sm_count();
void rt_set_device(int device);          // make `device` current on the calling thread (every ABI entry point does)
int rt_current_device();
void *rt_dmalloc(size_t n);
void rt_dfree(void *p);
void *rt_hmalloc(size_t n);             // pinned host memory
void rt_hfree(void *p);
void rt_h2d(void *d, const void *h, size_t n, rt_stream s);
void rt_d2h(void *h, const void *d, size_t n, rt_stream s);
void rt_d2d(void *d, const void *src, size_t n, rt_stream s);
void rt_memset(void *d, int v, size_t n, rt_stream s);
// device memory of src_device -> device memory of dst_device (NVLink / PCIe peer copy), on a stream of the current device
void rt_peer_copy(void *d, int dst_device, const void *src, int src_device, size_t n, rt_stream s);
rt_stream rt_stream_create();
void rt_stream_destroy(rt_stream s);
void rt_stream_sync(rt_stream s);
void *rt_stream_native(rt_stream s);    // cudaStream_t for callers that want to time with their own events
rt_event rt_event_create();
void rt_event_destroy(rt_event e);
void rt_event_record(rt_event e, rt_stream s);
void rt_stream_wait(rt_stream s, rt_event e);
void rt_event_sync(rt_event e);
float rt_event_elapsed_ms(rt_event a, rt_event b);
