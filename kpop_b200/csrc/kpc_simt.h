// kpc_simt.h -- the few device primitives the warp-level kernels use beyond plain CUDA C++ (PTX loads with cache
// hints, L2 policies, bulk copies completing on an mbarrier, shared-memory atomics by address), each with a plain
// C++ rendition for the test-only SIMT emulator (tests/emul/simt_emul.h, selected by KPC_SIMT_EMUL).  The product
// (libkpopcount_gpu.so) is always built without KPC_SIMT_EMUL; the emulator exists so that the kernel sources can be
// fuzzed against the oracle on a machine without a GPU.
#pragma once
#include <stdint.h>

#ifdef KPC_SIMT_EMUL
#include <string.h>

#include "simt_emul.h"
#define KP_SPIN_YIELD() simt::spin_yield()
#define KP_DEV static inline
#else
#include <cuda_runtime.h>
#define KP_SPIN_YIELD() do { } while (0)
#define KP_DEV __device__ __forceinline__
#endif

KP_DEV uint32_t kp_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
KP_DEV uint32_t kp_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

// ---- global memory -------------------------------------------------------------------------------------------------
KP_DEV unsigned long long kp_ld_relaxed_u64(const unsigned long long *p) {
#ifdef KPC_SIMT_EMUL
  return *(const volatile unsigned long long *)p;
#else
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
KP_DEV void kp_st_relaxed_u64(unsigned long long *p, unsigned long long v) {
#ifdef KPC_SIMT_EMUL
  *(volatile unsigned long long *)p = v;
#else
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
// 128-bit streaming load (read once: no L1 allocation)
KP_DEV uint4 kp_ldg_stream(const uint8_t *p) {
#ifdef KPC_SIMT_EMUL
  uint4 x;
  memcpy(&x, p, 16);
  return x;
#else
  uint4 x;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
               : "l"(p));
  return x;
#endif
}
// the same load with an L2 eviction policy (see kp_l2_policy_*)
KP_DEV uint4 kp_ldg_stream_hint(const uint8_t *p, unsigned long long policy) {
#ifdef KPC_SIMT_EMUL
  (void)policy;
  uint4 x;
  memcpy(&x, p, 16);
  return x;
#else
  if (!policy) return kp_ldg_stream(p);
  uint4 x;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
               : "l"(p), "l"(policy));
  return x;
#endif
}
// 256-bit store (one full 32-byte sector per thread; dst 32-byte aligned).  SASS: STG.E.ENL2.256
KP_DEV void kp_stg_256(void *dst, const uint4 &a, const uint4 &b) {
#ifdef KPC_SIMT_EMUL
  memcpy(dst, &a, 16);
  memcpy((uint8_t *)dst + 16, &b, 16);
#else
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#endif
}
// 256-bit streaming load (32-byte aligned: one whole sector per thread) with an L2 eviction policy.  SASS: LDG.E.NA.ENL2.256
KP_DEV void kp_ldg_stream_hint_256(const uint8_t *p, unsigned long long policy, uint4 &a, uint4 &b) {
#ifdef KPC_SIMT_EMUL
  (void)policy;
  memcpy(&a, p, 16);
  memcpy(&b, p + 16, 16);
#else
  if (policy)
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p), "l"(policy));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
#endif
}
// L2 eviction policies: a line read with evict_last stays until it is read with evict_first (its last use)
KP_DEV unsigned long long kp_l2_policy_evict_last() {
#ifdef KPC_SIMT_EMUL
  return 1;
#else
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#endif
}
KP_DEV unsigned long long kp_l2_policy_evict_first() {
#ifdef KPC_SIMT_EMUL
  return 2;
#else
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#endif
}
// pull [p, p + n) into L2 (n a multiple of 16)
KP_DEV void kp_prefetch_l2(const uint8_t *p, uint32_t n) {
#ifdef KPC_SIMT_EMUL
  (void)p; (void)n;
#else
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(n) : "memory");
#endif
}

// ---- shared memory by 32-bit address ---------------------------------------------------------------------------------
// On the device these take addresses in the shared window (cvta'd once per kernel); the emulator keeps an offset from
// the CTA's shared-memory base in the same 32 bits.
KP_DEV uint32_t kp_smem_addr(const void *p) {
#ifdef KPC_SIMT_EMUL
  return (uint32_t)((const uint8_t *)p - simt::smem);
#else
  return (uint32_t)__cvta_generic_to_shared(p);
#endif
}
KP_DEV uint32_t kp_atoms_add(uint32_t addr, uint32_t v) {
#ifdef KPC_SIMT_EMUL
  uint32_t *q = (uint32_t *)(simt::smem + addr);
  const uint32_t o = *q;
  *q = o + v;
  return o;
#else
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
#endif
}
KP_DEV void kp_sts_u16(uint32_t addr, uint32_t v) {
#ifdef KPC_SIMT_EMUL
  *(uint16_t *)(simt::smem + addr) = (uint16_t)v;
#else
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
#endif
}

// ---- bulk copy global -> shared, completion on an mbarrier (TMA, SASS: UBLKCP + SYNCS) ------------------------------------
KP_DEV void kp_mbar_init(unsigned long long *bar, uint32_t count) {
#ifdef KPC_SIMT_EMUL
  *bar = 0; (void)count;
#else
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kp_smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
// one thread: announce `bytes` and start the copy; dst, src and bytes are multiples of 16
KP_DEV void kp_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar, unsigned long long policy) {
#ifdef KPC_SIMT_EMUL
  (void)policy;
  memcpy(dst, src, bytes);
  *bar += 1;  // completed phases
#else
  const uint32_t b = kp_smem_addr(bar), d = kp_smem_addr(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  if (policy)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b), "l"(policy) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
#endif
}
// all threads: wait until phase number `phase` (0, 1, 2, ...) of the barrier has completed
KP_DEV void kp_mbar_wait(unsigned long long *bar, uint32_t phase) {
#ifdef KPC_SIMT_EMUL
  while (*(volatile unsigned long long *)bar <= phase) KP_SPIN_YIELD();
#else
  const uint32_t b = kp_smem_addr(bar), parity = phase & 1u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "KP_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra KP_MBAR_WAIT;\n\t"
      "}" ::"r"(b), "r"(parity) : "memory");
#endif
}
// generic-proxy writes to shared memory that a later bulk copy will overwrite / that the async proxy must observe
KP_DEV void kp_fence_proxy_async() {
#ifndef KPC_SIMT_EMUL
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
