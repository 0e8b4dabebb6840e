// kpc_bucketsort.cuh -- the second half of the sort path for large-k samples.
//
// After the counting pass of the framing kernel and the streaming scatter (KpcBucketCountSink / KpcBucketScatterSink, kpc_tile.cuh)
// the (key, rank) pairs of all windows of a sample sit grouped by coarse bucket -- the top bits of (key mod B), B =
// the bucket count of OCaml's Hashtbl (BiOCamLib/lib/Better.ml:741, KMers.ml:99-113) -- in arbitrary order inside a
// group.  This file turns every group into what KIHF.iter prints (bin/KPopCount.ml:60):
//   * equal keys merge: count = multiplicity (IntHashFrequencies.add, KMers.ml:107-111), rank = first occurrence;
//   * entries come out by ascending (key mod B) and, inside one OCaml bucket, newest first (Hashtbl.add conses at the
//     bucket head, Hashtbl.iter walks head to tail: SURVEY.md App. A.4).
// Groups are tiny for real genomes (0.3 keys per OCaml bucket at k = 30): one thread sorts a group of <= 24 pairs in
// local memory; a CTA sorts larger ones (<= 2048) in shared memory by ranking; anything beyond sets a flag and the
// engine redoes the sample on the hash-table path.
//
// Compiled by nvcc (product) and by g++ under the SIMT emulator (KPC_SIMT_EMUL, tests only).
#pragma once
#include "kpc_kernels.h"
#include "kpc_simt.h"

#ifndef KPC_BS_SMALL_CFG
#define KPC_BS_SMALL_CFG 24
#endif
#ifndef KPC_BS_HEAVY_CFG
#define KPC_BS_HEAVY_CFG 2048
#endif
constexpr int KPC_BS_SMALL = KPC_BS_SMALL_CFG;  // pairs one thread sorts
constexpr int KPC_BS_HEAVY = KPC_BS_HEAVY_CFG;  // pairs one CTA sorts (the emulation build uses tiny values)

// order 1: (key mod B, key, rank) -- equal keys adjacent, first occurrence first
KP_DEV bool kpc_bs_less1(unsigned long long ka, unsigned long long ra, unsigned long long kb, unsigned long long rb,
                         unsigned long long bmask) {
  const unsigned long long ba = ka & bmask, bb = kb & bmask;
  if (ba != bb) return ba < bb;
  if (ka != kb) return ka < kb;
  return ra < rb;
}
// order 2 (final): (key mod B) ascending, newest (largest first-occurrence rank) first
KP_DEV bool kpc_bs_less2(unsigned long long ka, unsigned long long ra, unsigned long long kb, unsigned long long rb,
                         unsigned long long bmask) {
  const unsigned long long ba = ka & bmask, bb = kb & bmask;
  if (ba != bb) return ba < bb;
  return ra > rb;
}

// one thread per coarse bucket
KP_DEV void kpc_bucket_finalize_small(const KpcBucketFinalize &F, uint32_t b, uint32_t &distinct) {
  const uint32_t lo = F.offsets[b], hi = F.offsets[b + 1], n = hi - lo;
  if (n == 0) return;
  if (n > (uint32_t)KPC_BS_SMALL) {
    const uint32_t i = (uint32_t)atomicAdd(F.stats + 2, 1ull);
    if (i < F.heavy_cap) F.heavy_list[i] = b; else atomicAdd(F.stats + 1, 1ull);
    return;
  }
  unsigned long long k[KPC_BS_SMALL], r[KPC_BS_SMALL], c[KPC_BS_SMALL];
  for (uint32_t i = 0; i < n; ++i) {  // insertion sort, order 1
    const KpcPair pr = F.pairs[lo + i];
    const unsigned long long ki = pr.key, ri = pr.rank;
    uint32_t j = i;
    while (j > 0 && kpc_bs_less1(ki, ri, k[j - 1], r[j - 1], F.bmask)) { k[j] = k[j - 1]; r[j] = r[j - 1]; --j; }
    k[j] = ki; r[j] = ri;
  }
  uint32_t m = 0;  // merge runs of equal keys
  for (uint32_t i = 0; i < n; ++i) {
    if (m && k[m - 1] == k[i]) { c[m - 1] += 1ull; continue; }
    k[m] = k[i]; r[m] = r[i]; c[m] = 1ull; ++m;
  }
  for (uint32_t i = 1; i < m; ++i) {  // insertion sort, order 2
    const unsigned long long ki = k[i], ri = r[i], ci = c[i];
    uint32_t j = i;
    while (j > 0 && kpc_bs_less2(ki, ri, k[j - 1], r[j - 1], F.bmask)) { k[j] = k[j - 1]; r[j] = r[j - 1]; c[j] = c[j - 1]; --j; }
    k[j] = ki; r[j] = ri; c[j] = ci;
  }
  for (uint32_t i = 0; i < n; ++i) {
    F.keys[lo + i] = i < m ? k[i] : ~0ull;
    F.ranks[lo + i] = i < m ? r[i] : ~0ull;
    F.counts[lo + i] = i < m ? c[i] : 0ull;
  }
  distinct += m;
}

template <int NT>
KP_DEV void kpc_bucket_finalize_small_body(const KpcBucketFinalize &F) {
  uint32_t distinct = 0;
  for (uint32_t b = blockIdx.x * (uint32_t)NT + threadIdx.x; b < F.nb; b += gridDim.x * (uint32_t)NT)
    kpc_bucket_finalize_small(F, b, distinct);
  distinct = __reduce_add_sync(0xffffffffu, distinct);
  if ((threadIdx.x & 31) == 0 && distinct) atomicAdd(F.stats, (unsigned long long)distinct);
}

// a CTA per heavy group: sort by ranking (position = number of smaller elements; ranks are unique, so every order
// here is strict), merge, rank again in the final order
struct KpcBsHeavySmem {
  unsigned long long k[KPC_BS_HEAVY], r[KPC_BS_HEAVY], k2[KPC_BS_HEAVY], r2[KPC_BS_HEAVY];
  uint32_t c2[KPC_BS_HEAVY];
  uint32_t head[KPC_BS_HEAVY];  // index of the entry (run of equal keys) an element belongs to
  uint32_t m;
};
template <int NT>
KP_DEV void kpc_bucket_finalize_heavy_body(const KpcBucketFinalize &F, uint8_t *smem_raw) {
  KpcBsHeavySmem &S = *reinterpret_cast<KpcBsHeavySmem *>(smem_raw);
  const uint32_t tid = threadIdx.x;
  unsigned long long nh = F.stats[2];
  if (nh > F.heavy_cap) nh = F.heavy_cap;
  for (uint32_t h = blockIdx.x; h < (uint32_t)nh; h += gridDim.x) {
    const uint32_t b = F.heavy_list[h];
    const uint32_t lo = F.offsets[b], n = F.offsets[b + 1] - lo;
    if (n > (uint32_t)KPC_BS_HEAVY) {
      if (tid == 0) atomicAdd(F.stats + 1, 1ull);
      continue;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NT) { const KpcPair pr = F.pairs[lo + i]; S.k[i] = pr.key; S.r[i] = pr.rank; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NT) {  // order 1 by ranking
      const unsigned long long ki = S.k[i], ri = S.r[i];
      uint32_t pos = 0;
      for (uint32_t j = 0; j < n; ++j) pos += kpc_bs_less1(S.k[j], S.r[j], ki, ri, F.bmask) ? 1u : 0u;
      S.k2[pos] = ki; S.r2[pos] = ri;
    }
    __syncthreads();
    // runs of equal keys: thread 0 numbers them (n <= 2048: a few microseconds, and heavy groups are rare)
    if (tid == 0) {
      uint32_t m = 0;
      for (uint32_t i = 0; i < n; ++i) {
        if (i == 0 || S.k2[i] != S.k2[i - 1]) { S.k[m] = S.k2[i]; S.r[m] = S.r2[i]; S.c2[m] = 0; ++m; }
        S.c2[m - 1] += 1u;
      }
      S.m = m;
    }
    __syncthreads();
    const uint32_t m = S.m;
    for (uint32_t i = tid; i < n; i += NT) { F.keys[lo + i] = ~0ull; F.ranks[lo + i] = ~0ull; F.counts[lo + i] = 0ull; }
    __syncthreads();
    for (uint32_t i = tid; i < m; i += NT) {  // order 2 by ranking, straight to the output
      const unsigned long long ki = S.k[i], ri = S.r[i];
      uint32_t pos = 0;
      for (uint32_t j = 0; j < m; ++j) pos += kpc_bs_less2(S.k[j], S.r[j], ki, ri, F.bmask) ? 1u : 0u;
      F.keys[lo + pos] = ki; F.ranks[lo + pos] = ri; F.counts[lo + pos] = (unsigned long long)S.c2[i];
    }
    if (tid == 0) atomicAdd(F.stats, (unsigned long long)m);
  }
}
