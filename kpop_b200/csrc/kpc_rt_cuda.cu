// kpc_rt_cuda.cu -- CUDA runtime implementation of kpc_rt.h (the only one that ships).
#include <cuda_runtime.h>

#include "../../include/kpopcount.h"
#include "kpc_rt.h"

#define RT_CHECK(x)                                                                                     \
  do {                                                                                                  \
    cudaError_t e_ = (x);                                                                               \
    if (e_ != cudaSuccess) {                                                                            \
      int code_ = (e_ == cudaErrorMemoryAllocation) ? KPC_E_NOMEM : KPC_E_CUDA;                         \
      throw KpcError(code_, std::string(#x) + ": " + cudaGetErrorString(e_));                           \
    }                                                                                                   \
  } while (0)

static int g_sm_count = 0;

const char *rt_backend_name() { return "cuda"; }

void rt_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    throw KpcError(KPC_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                   "); libkpopcount_gpu has no CPU fallback");
  if (device < 0 || device >= n) throw KpcError(KPC_E_ARG, "device index out of range");
  RT_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  RT_CHECK(cudaGetDeviceProperties(&prop, device));
  g_sm_count = prop.multiProcessorCount;
}
int rt_sm_count() { return g_sm_count ? g_sm_count : 148; }
void rt_set_device(int device) { RT_CHECK(cudaSetDevice(device)); }
int rt_current_device() {
  int d = 0;
  RT_CHECK(cudaGetDevice(&d));
  return d;
}
void *rt_dmalloc(size_t n) {
  void *p = nullptr;
  RT_CHECK(cudaMalloc(&p, n ? n : 1));
  return p;
}
void rt_dfree(void *p) {
  if (p) cudaFree(p);
}
void *rt_hmalloc(size_t n) {
  void *p = nullptr;
  RT_CHECK(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault));
  return p;
}
void rt_hfree(void *p) {
  if (p) cudaFreeHost(p);
}
static inline cudaStream_t cs(rt_stream s) { return (cudaStream_t)s; }
void rt_h2d(void *d, const void *h, size_t n, rt_stream s) {
  if (n) RT_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, cs(s)));
}
void rt_d2h(void *h, const void *d, size_t n, rt_stream s) {
  if (n) RT_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, cs(s)));
}
void rt_d2d(void *d, const void *src, size_t n, rt_stream s) {
  if (n) RT_CHECK(cudaMemcpyAsync(d, src, n, cudaMemcpyDeviceToDevice, cs(s)));
}
void rt_peer_copy(void *d, int dst_device, const void *src, int src_device, size_t n, rt_stream s) {
  if (n) RT_CHECK(cudaMemcpyPeerAsync(d, dst_device, src, src_device, n, cs(s)));
}
void rt_memset(void *d, int v, size_t n, rt_stream s) {
  if (n) RT_CHECK(cudaMemsetAsync(d, v, n, cs(s)));
}
rt_stream rt_stream_create() {
  cudaStream_t s;
  RT_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return (rt_stream)s;
}
void rt_stream_destroy(rt_stream s) {
  if (s) cudaStreamDestroy(cs(s));
}
void rt_stream_sync(rt_stream s) { RT_CHECK(cudaStreamSynchronize(cs(s))); }
void *rt_stream_native(rt_stream s) { return (void *)s; }
rt_event rt_event_create() {
  cudaEvent_t e;
  RT_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return (rt_event)e;
}
void rt_event_destroy(rt_event e) {
  if (e) cudaEventDestroy((cudaEvent_t)e);
}
void rt_event_record(rt_event e, rt_stream s) { RT_CHECK(cudaEventRecord((cudaEvent_t)e, cs(s))); }
void rt_stream_wait(rt_stream s, rt_event e) { RT_CHECK(cudaStreamWaitEvent(cs(s), (cudaEvent_t)e, 0)); }
void rt_event_sync(rt_event e) { RT_CHECK(cudaEventSynchronize((cudaEvent_t)e)); }
float rt_event_elapsed_ms(rt_event, rt_event) { return 0.f; }
