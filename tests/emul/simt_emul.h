// tests/emul/simt_emul.h -- TEST INFRASTRUCTURE ONLY.
// A small SIMT emulator: runs a CUDA-style kernel body (threads, warps, CTAs, barriers, warp collectives, shared
// memory, atomics) on ONE host thread with cooperative fibers (ucontext), so that the warp-level kernels of
// kpop_b200/csrc (kpc_partition.cuh, ...) can be compiled by g++ and fuzzed against the oracle without a GPU.
// Scheduling is deterministic: CTAs round-robin, warps in order, lanes in order; a fiber runs until it reaches a
// barrier, a warp collective or a spin-wait.  Nothing under kpop_b200/ builds, loads or falls back to this.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <functional>

namespace simt {

struct Dim3 { unsigned x, y, z; };
extern Dim3 threadIdx, blockIdx, blockDim, gridDim;
extern uint8_t *smem;  // dynamic shared memory of the current CTA

// run body() once per thread of a grid x block launch (block a multiple of 32); returns when every thread has exited
void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body);

void syncthreads();
int syncthreads_or(int pred);
void syncwarp();
void spin_yield();  // inside a wait loop on memory written by another CTA / warp

// warp collectives (all non-exited lanes of the warp must take part)
uint64_t shfl_idx64(uint64_t v, int src);
uint64_t shfl_up64(uint64_t v, unsigned d);
uint64_t shfl_down64(uint64_t v, unsigned d);
uint64_t shfl_xor64(uint64_t v, unsigned m);
uint32_t ballot(int pred);
uint32_t reduce_add(uint32_t v);
uint32_t reduce_max(uint32_t v);
uint32_t reduce_min(uint32_t v);

}  // namespace simt

// ---- CUDA spellings ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
using simt::blockDim;
using simt::blockIdx;
using simt::gridDim;
using simt::threadIdx;

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

static inline void __syncthreads() { simt::syncthreads(); }
static inline int __syncthreads_or(int p) { return simt::syncthreads_or(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::syncwarp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  uint64_t r = simt::shfl_idx64((uint64_t)v, src);
  return (T)r;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) { return (T)simt::shfl_up64((uint64_t)v, d); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) { return (T)simt::shfl_down64((uint64_t)v, d); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, unsigned m) { return (T)simt::shfl_xor64((uint64_t)v, m); }
static inline unsigned __ballot_sync(unsigned, int p) { return simt::ballot(p); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return simt::reduce_add(v); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return simt::reduce_max(v); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return simt::reduce_min(v); }

static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __brev(uint32_t v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(v);
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {  // top word of (hi:lo) << (s & 31)
  s &= 31u;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {  // low word of (hi:lo) >> (s & 31)
  s &= 31u;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {  // PRMT, default mode
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = (sel >> (4 * i)) & 0xFu;
    uint32_t byte = (uint32_t)(src >> (8 * (s & 7u))) & 0xFFu;
    if (s & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;  // sign replication
    r |= byte << (8 * i);
  }
  return r;
}
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {  // unsigned 8-bit dot product + c
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFFu) * ((b >> (8 * i)) & 0xFFu);
  return c;
}
template <class T> static inline T __ldg(const T *p) { return *p; }

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = (T)(o | v); return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
