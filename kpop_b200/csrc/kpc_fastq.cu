// kpc_fastq.cu -- fast FASTQ -> dense 4^k table pipeline for sm_100a (see kpc_fastq.h for the outline).
//
// Reference semantics restated here (paths relative to the KPop tree):
//   Files.FASTQ.iter_se        BiOCamLib/lib/Files.ml:201-221   4-line records, '@' / '+' checks, only line 2 is sequence
//   Sequences.Lint.dnaize      BiOCamLib/lib/Sequences.ml:41-67  [ACGTacgt] are bases, every other byte breaks k-mers
//   DNAHash*.iteri / iterc     BiOCamLib/lib/KMers.ml:319-349, 357-389   first base most significant, key = min f rc
//   IntHashFrequencies.add     BiOCamLib/lib/KMers.ml:107-111   count[key] += 1
//
// Geometry: a CTA of FQ_NT threads owns one tile of FQ_TB bytes at a time (claimed in stream order); DESIGN.md section 5
// has the measurements behind every choice.
//   1. 128-bit loads (L2 evict_first) -> shared memory; the newline census is taken from the registers on the way:
//      three integer operations per word flag the line feeds, dot products (IDP.4A) gather the flags into a 16-bit mask
//   2. warp scans over packed counts + one cross-warp scan number the line feeds; the number of line feeds BEFORE the
//      tile comes from a decoupled look-back over one word per tile.  Every CTA counts its NEXT tile one tile ahead (a
//      second read, L2 evict_last, of a tile that was bulk-prefetched into L2 two grids earlier) and publishes the count
//      then, so that the look-back never waits for a predecessor
//   3. every line gets a slot (start, end, line index mod 4): lines 0 and 2 of a record are checked for '@' / '+',
//      line 1 is a row; rows are cut into units of FQ_W window-end positions, numbered by a division when all rows
//      have (about) the same length and by a block scan + unit table otherwise
//   4. one thread per unit: 12 + 16 bytes -> 2-bit codes + "not a base" flags with SIMD-in-register arithmetic and
//      dot-product gathers, reverse complement of the 28 bases from two BREVs
//   5. per window: forward and reverse-complement k-mers funnel-shifted to the top of a word, unsigned min, slice and
//      queue entry by shift / PRMT, one shared-memory atomic + one predicated 16-bit store into the slice's bucket --
//      15 SASS instructions per k-mer, no branch (fq_append_k12).  Software write-combining: after every round the
//      thread that owns a slice reserves whole 32-byte chunks of its bucket in the slice's queue in HBM (one global
//      atomic, issued a round before its result is needed) and copies them out with 128-bit accesses.  The slice is
//      taken from the MIDDLE bits of the key: min(f, rc) skews the top and the bottom bases of a canonical k-mer but
//      leaves the central ones uniform, so the buckets fill evenly; a bucket that overflows anyway (skewed input)
//      sends its keys to the global table with RED.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include <string>
#include <vector>

#include "../../include/kpopcount.h"
#include "kpc_fastq.h"

#define FQ_CUDA_CHECK(x)                                                                                      \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      throw KpcError(KPC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +        \
                                     std::to_string(__LINE__) + ")");                                         \
  } while (0)

// dynamic shared memory of both kernels; at file scope so that its (constant) shared-space address can be named in PTX
extern __shared__ __align__(16) uint8_t fq_smem_raw[];

namespace {

#ifndef FQ_NT_CFG
#define FQ_NT_CFG 512
#endif
constexpr int FQ_NT = FQ_NT_CFG;                  // threads per CTA (a multiple of 32, at least FQ_MAXSLICES)
constexpr int FQ_NW = FQ_NT / 32;
constexpr int FQ_W = 16;                          // window-end positions per unit (one thread)
constexpr int FQ_CTX = 12;                        // context bytes loaded before a unit (>= k - 1)
constexpr int FQ_PIECES = 4;                      // 16-byte vectors per thread
constexpr int FQ_TB = 16 * FQ_PIECES * FQ_NT;     // tile bytes (32 KiB with 512 threads)
constexpr int FQ_HALO = 16;
constexpr int FQ_MAXROWS = FQ_NT;                 // sequence lines per batch
constexpr int FQ_MAXSLOTS = 4 * FQ_NT;            // lines per batch
constexpr int FQ_MAXUNITS = FQ_TB / FQ_W + FQ_MAXROWS + 8;  // units per batch: sum of ceil(len / W) over its rows
constexpr int FQ_BIGROW = 32;                     // rows with more units are expanded by the whole CTA
constexpr int FQ_MAXBIG = FQ_TB / (FQ_BIGROW * FQ_W) + 2;
constexpr int FQ_MAXSLICES = 512;
constexpr int FQ_BIAS = 17;                       // positions are stored + FQ_BIAS (they start at -17)
constexpr int FQ_BUCKET_ENTRIES = 24576;          // shared-memory bucket space (u16 entries) shared by all slices
constexpr int FQ_BPAD = 8;                        // entries between two buckets: a stride of cap + 8 entries (28 or 100 words) keeps the owners' 128-bit accesses free of bank conflicts
constexpr int FQ_CHUNK = 16;                      // entries per copy-out chunk (32 bytes: one L2 sector)
constexpr uint16_t FQ_PAD = 0xFFFFu;              // queue entry that pads the last chunk of a CTA (skipped by fq_count)
static_assert(FQ_PIECES == 4, "the census packs four newline counts into two scans");
static_assert(FQ_TB + FQ_BIAS < 65536 && FQ_NT % 32 == 0, "positions are 16-bit");
static_assert(FQ_MAXSLICES <= FQ_NT, "one thread owns one slice");

struct FqBigRow { uint32_t ub, info, n; };
struct FqSmem {
  alignas(16) uint8_t raw[FQ_HALO + FQ_TB + 48];  // raw[16 + i] = tile byte i
  alignas(16) uint16_t bucket[FQ_BUCKET_ENTRIES + FQ_BPAD * FQ_MAXSLICES + 2 * FQ_CHUNK];  // slice s owns [s * (cap + FQ_BPAD), + cap)
  uint32_t fill[FQ_MAXSLICES];                    // entries in the bucket (may run past cap while appending)
  uint32_t qb16[FQ_MAXSLICES];                    // queue base of the slice / FQ_CHUNK
  uint32_t qcap[FQ_MAXSLICES];
  uint32_t dummy[32];                             // fill[FQ_MAXSLICES + lane]: where the appends of invalid windows count
  uint32_t uinfo[FQ_MAXUNITS];                    // unit -> first window end (low half) | end of its line (high half)
  uint32_t wtot_a[FQ_PIECES * FQ_NW];
  uint32_t wtot_b[32];
  FqBigRow big[FQ_MAXBIG];
  unsigned long long G;                           // number of '\n' in the stream before the tile
  uint32_t tileq[2];
  uint32_t nbig;
  uint32_t cnt_next;                              // newline count of the next tile (fq_count_tile)
  uint32_t umax, usum;                            // longest row of the batch (units) and the sum over its rows
  int head;                                       // position of the last '\n' before the tile (-1 .. -16), or -17
  uint16_t nlpos[FQ_MAXSLOTS + 2];
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream(const uint8_t *p) {
  uint4 x;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
               : "l"(p));
  return x;
}
// the same load with an L2 eviction policy: the count-ahead pass wants its tile to stay in L2 until the tile is
// processed (evict_last), the processing pass reads it for the last time (evict_first)
__device__ __forceinline__ uint4 ldg_stream_hint(const uint8_t *p, unsigned long long policy) {
  uint4 x;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
               : "l"(p), "l"(policy));
  return x;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// exclusive prefix over the CTA (thread order), one barrier; wtot must not be in use by a slower warp
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *wtot, uint32_t &total, int lane, int w) {
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  uint32_t t = lane < FQ_NW ? wtot[lane] : 0u;
  uint32_t tinc = warp_incl_scan(t, lane);
  total = __shfl_sync(0xffffffffu, tinc, 31);
  uint32_t wex = __shfl_sync(0xffffffffu, tinc - t, w);
  return wex + inc - v;
}
// 0x80 in every byte of w that equals '\n' (exact: three operations; the two-constant logic ops are written as LOP3 so
// that ptxas keeps one constant in a uniform register instead of splitting the operation)
__device__ __forceinline__ uint32_t nl_mask(uint32_t w) {
  uint32_t t, z;
  asm("lop3.b32 %0, %1, 0x0A0A0A0A, 0x7F7F7F7F, 0x28;" : "=r"(t) : "r"(w));   // (w ^ 0x0A..) & 0x7F..
  t += 0x7F7F7F7Fu;
  asm("lop3.b32 %0, %1, %2, 0x80808080, 0x02;" : "=r"(z) : "r"(t), "r"(w));    // ~(t | w) & 0x80..
  return z;
}
// bit i of the result <=> byte i of the 16-byte vector is '\n': the 0x80 flags are gathered with dot products
__device__ __forceinline__ uint32_t nl_mask16(const uint4 &x) {
  uint32_t lo = __dp4a(nl_mask(x.x), 0x08040201u, 0u);
  lo = __dp4a(nl_mask(x.y), 0x80402010u, lo);
  uint32_t hi = __dp4a(nl_mask(x.z), 0x08040201u, 0u);
  hi = __dp4a(nl_mask(x.w), 0x80402010u, hi);
  return (lo + (hi << 8)) >> 7;
}

// number of '\n' in the 16-byte vector
__device__ __forceinline__ uint32_t nl_count16(const uint4 &x) {  // 0x80 * count, summed with dot products
  uint32_t c = __dp4a(nl_mask(x.x), 0x01010101u, 0u);
  c = __dp4a(nl_mask(x.y), 0x01010101u, c);
  c = __dp4a(nl_mask(x.z), 0x01010101u, c);
  return __dp4a(nl_mask(x.w), 0x01010101u, c) >> 7;
}
// newline count of a whole tile, one tile ahead of its processing: threads [first, FQ_NT) add their share to *acc
__device__ __forceinline__ void fq_count_tile(const KpcFqLaunch &p, uint32_t tile, int tid, int first, uint32_t *acc) {
  const uint64_t t0 = (uint64_t)tile * FQ_TB;
  const int len = (int)((p.n - t0) < (uint64_t)FQ_TB ? (p.n - t0) : (uint64_t)FQ_TB);
  uint32_t c = 0;
  constexpr int NV = FQ_TB / 16;
  const unsigned long long pol_keep = l2_policy_evict_last();
  if (len == FQ_TB && first == 32) {  // full tile, 15 warps: all the loads of a thread are in flight together
    constexpr int ROUNDS = (NV + FQ_NT - 32 - 1) / (FQ_NT - 32);
    uint4 x[ROUNDS];
#pragma unroll
    for (int u = 0; u < ROUNDS; ++u) {
      const int i = tid - 32 + u * (FQ_NT - 32);
      x[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < NV) x[u] = ldg_stream_hint(p.data + t0 + 16 * i, pol_keep);
    }
#pragma unroll
    for (int u = 0; u < ROUNDS; ++u) c += nl_count16(x[u]);
  } else {
    for (int i = tid - first; 16 * i < len; i += FQ_NT - first) {
      uint4 x = ldg_stream(p.data + t0 + 16 * i);
      if (16 * i + 16 > len) {  // bytes past the end read as 0
        uint32_t *xw = reinterpret_cast<uint32_t *>(&x);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int rem = len - (16 * i + 4 * m);
          if (rem <= 0) xw[m] = 0u;
          else if (rem < 4) xw[m] &= (1u << (8 * rem)) - 1u;
        }
      }
      c += nl_count16(x);
    }
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((tid & 31) == 0 && c) atomicAdd(acc, c);
}

// slice / bin of a key and back: slice = key bits [lo, lo + sb), bin = the other bits packed together
__device__ __forceinline__ uint32_t fq_slice_of(uint32_t key, int lo, uint32_t smask) { return (key >> lo) & smask; }
__device__ __forceinline__ uint32_t fq_bin_of(uint32_t key, int lo, int sb, uint32_t lomask) {
  return ((key >> (lo + sb)) << lo) | (key & lomask);
}
__device__ __forceinline__ uint32_t fq_key_of(uint32_t slice, uint32_t bin, int lo, int sb, uint32_t lomask) {
  return ((bin >> lo) << (lo + sb)) | (slice << lo) | (bin & lomask);
}

// number of '\n' before the tile: decoupled look-back over one word per tile (2 status bits + 62 value bits).
// All CTAs work on neighbouring tiles at the same time, so the nearest tile with an inclusive count is usually about
// one grid (2 x 148 tiles) back: the warp loads FQ_LB_ROUNDS x 32 states in one go (the loads overlap) before it looks
// at any of them, which makes the common case one L2 round trip.
constexpr unsigned long long FQ_ST_AGG = 1ull << 62, FQ_ST_INC = 2ull << 62, FQ_VAL = (1ull << 62) - 1ull;
constexpr int FQ_LB_ROUNDS = 10;
// the tile's own count is published one tile ahead of its processing (fq_count_tile), so that by the time a tile
// looks back every predecessor has published at least its aggregate: nobody waits for anybody
__device__ __forceinline__ void lookback_publish(unsigned long long *state, uint32_t tile, uint32_t total,
                                                 unsigned long long g_in) {
  st_relaxed_u64(state + tile, tile == 0 ? (FQ_ST_INC | (g_in + total)) : (FQ_ST_AGG | (unsigned long long)total));
}
__device__ __forceinline__ unsigned long long lookback(unsigned long long *state, uint32_t tile, uint32_t total,
                                                       unsigned long long g_in, int lane) {
  if (tile == 0) return g_in;
  unsigned long long acc = 0;
  long long j0 = (long long)tile - 1;
  for (;;) {
    unsigned long long v[FQ_LB_ROUNDS];
#pragma unroll
    for (int i = 0; i < FQ_LB_ROUNDS; ++i) {
      const long long j = j0 - 32 * i - lane;
      if (j >= 0) v[i] = ld_relaxed_u64(state + j);
      else if (j == -1) v[i] = FQ_ST_INC | g_in;
      else v[i] = FQ_ST_AGG;  // never used: lies behind the inclusive entry at j == -1
    }
    bool done = false, stale = false;
    unsigned long long part = 0;
#pragma unroll
    for (int i = 0; i < FQ_LB_ROUNDS; ++i) {
      if (done || stale) continue;
      const unsigned inc_mask = __ballot_sync(0xffffffffu, (v[i] >> 62) == 2ull);
      const unsigned inv_mask = __ballot_sync(0xffffffffu, (v[i] >> 62) == 0ull);
      const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
      const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
      if (inv_mask & need) { stale = true; continue; }  // a predecessor has not published yet: reload from here
      part += ((need >> lane) & 1u) ? (v[i] & FQ_VAL) : 0ull;
      if (first < 32) done = true; else j0 -= 32;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    acc += part;
    if (done) break;
  }
  if (lane == 0) st_relaxed_u64(state + tile, FQ_ST_INC | (acc + total));
  return acc;
}

// ---- bucket copy-out (step 5) ------------------------------------------------------------------------------------
// Thread s owns slice s.  reserve: called after the barrier that follows a round of appends; whole chunks of the
// bucket are reserved in the slice's queue with one global atomic whose result is not needed before fq_flush_copy
// (one round of k-mer arithmetic later).  copy: the owner moves the reserved chunks with 128-bit accesses and the
// bucket's remainder (< one chunk) to the front; a barrier must follow before the buckets are appended to again.
__device__ __forceinline__ void fq_flush_reserve(FqSmem &S, const KpcFqLaunch &p, int tid, uint32_t NS, uint32_t cap,
                                                 bool final, uint32_t &my_n, uint32_t &my_g) {
  my_n = 0; my_g = 0;
  if ((uint32_t)tid < NS) {
    uint32_t f = S.fill[tid];
    if (f > cap) f = cap;
    uint32_t n = f & ~(uint32_t)(FQ_CHUNK - 1);
    if (final && n < f) {  // the CTA is leaving: pad the last chunk
      for (uint32_t i = f; i < n + FQ_CHUNK; ++i) S.bucket[tid * (cap + FQ_BPAD) + i] = FQ_PAD;
      n += FQ_CHUNK;
      f = n;
    }
    S.fill[tid] = f - n;
    my_n = n;
    if (n) my_g = atomicAdd(p.qcursor + tid, n);
  }
}
__device__ __forceinline__ void fq_flush_copy(FqSmem &S, const KpcFqLaunch &p, int tid, uint32_t cap, uint32_t my_n,
                                              uint32_t my_g, int lo, int sb, uint32_t lomask) {
  if (!my_n) return;
  const uint32_t qc = S.qcap[tid];
  uint4 *src = reinterpret_cast<uint4 *>(S.bucket + tid * (cap + FQ_BPAD));
  uint4 *dst = reinterpret_cast<uint4 *>(p.queue + (unsigned long long)S.qb16[tid] * FQ_CHUNK + my_g);
  for (uint32_t c = 0; c < my_n; c += FQ_CHUNK) {
    const uint4 v0 = src[c >> 3], v1 = src[(c >> 3) + 1];
    if (my_g + c + FQ_CHUNK <= qc) {
      dst[c >> 3] = v0;
      dst[(c >> 3) + 1] = v1;
    } else {  // queue full: count in place
      const uint32_t ww[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t e = (ww[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        if (e != FQ_PAD) atomicAdd(p.table + fq_key_of(tid, e, lo, sb, lomask), 1u);
      }
    }
  }
  if (my_n < cap) {
    const uint4 r0 = src[my_n >> 3], r1 = src[(my_n >> 3) + 1];
    src[0] = r0;
    src[1] = r1;
  }
}

// ---- one k-mer: slice, bucket append (step 5) --------------------------------------------------------------------
// kk holds the canonical k-mer in its TOP 2k bits (the bits below are ignored).  `pending & bit` says whether the
// window is valid; the bit is cleared once the key sits in its bucket.  k = 12: slice = key bits [8, 17), the queue
// entry is key[0, 8) | key[17, 24) << 8 (one PRMT); everything is predicated, there is no branch.
__device__ __forceinline__ void fq_append_k12(uint32_t kk, uint32_t &pending, uint32_t bit, uint32_t s_fill,
                                              uint32_t s_bucket, uint32_t dummy_off, uint32_t ok) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .u32 t, o, os, en, pos, ad;\n\t"
      ".reg .u16 e16;\n\t"
      "and.b32 t, %6, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "shr.u32 o, %1, 14;\n\t"
      "and.b32 o, o, 0x7FC;\n\t"
      "selp.u32 os, o, %5, p;\n\t"
      "add.u32 os, os, %3;\n\t"
      "atom.shared.add.u32 pos, [os], 1;\n\t"
      "setp.lt.and.u32 q, pos, 48, p;\n\t"
      "shr.u32 t, %1, 1;\n\t"
      "prmt.b32 en, %1, t, 0x0071;\n\t"
      "cvt.u16.u32 e16, en;\n\t"
      "mad.lo.u32 ad, o, 28, %4;\n\t"             // bucket of the slice: (48 + FQ_BPAD) entries = 112 bytes apart
      "shl.b32 t, pos, 1;\n\t"
      "add.u32 ad, ad, t;\n\t"
      "@q st.shared.u16 [ad], e16;\n\t"
      "@q xor.b32 %0, %0, %2;\n\t"
      "}"
      : "+r"(pending)
      : "r"(kk), "r"(bit), "r"(s_fill), "r"(s_bucket), "r"(dummy_off), "r"(ok)
      : "memory");
}
// top 32 bits of (hi:lo) << s, 0 <= s < 64
__device__ __forceinline__ uint32_t fq_top_word(uint32_t hi, uint32_t lo, int s) {
  if (s >= 32) return lo << (s - 32);
  if (s == 0) return hi;
  return __funnelshift_l(lo, hi, s);
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}

#ifdef FQ_PROFILE
#define FQ_T(i) do { if (FQ_PMAP(i) >= 0) { const uint32_t t_ = (uint32_t)clock(); prof[FQ_PMAP(i) < 0 ? 0 : FQ_PMAP(i)] += t_ - tlast; tlast = t_; } } while (0)
// profile slots: 0 load..sync1 | 1 sync1..sync2 (scan, look-back, newline positions) | 2 sync2..sync4 (rows, units) |
//                3 classify | 4 copy-out + barrier | 5 append + barrier + reserve
#if !defined(FQ_PSET) || FQ_PSET == 0
#define FQ_PMAP(i) ((i) == 1 ? 0 : (i) == 3 ? 1 : (i) == 4 ? 2 : (i) == 5 ? 3 : (i) == 7 ? 4 : (i) == 9 ? 5 : -1)
#else  // inside "rows": 0 everything up to sync2 | 1 malformed check | 2 row info | 3 reductions | 4 barrier (3) | 5 the rest of the tile
#define FQ_PMAP(i) ((i) == 3 ? 0 : (i) == 11 ? 1 : (i) == 12 ? 2 : (i) == 15 ? 3 : (i) == 13 ? 4 : (i) == 9 ? 5 : -1)
#endif
#else
#define FQ_T(i) do { } while (0)
#endif
template <bool DS, int KT>
__global__ void __launch_bounds__(FQ_NT, 2) fq_partition_kernel(const KpcFqLaunch p) {
#ifdef FQ_PROFILE
  uint32_t prof[6] = {0, 0, 0, 0, 0, 0};
  uint32_t tlast = (uint32_t)clock();
#endif
  FqSmem &S = *reinterpret_cast<FqSmem *>(fq_smem_raw);
  uint32_t s_base;
  asm("mov.u32 %0, fq_smem_raw;" : "=r"(s_base));
  const uint32_t s_fill = s_base + (uint32_t)offsetof(FqSmem, fill);
  const uint32_t s_bucket = s_base + (uint32_t)offsetof(FqSmem, bucket);
  const uint32_t dummy_off = (uint32_t)(offsetof(FqSmem, dummy) - offsetof(FqSmem, fill)) + 4u * (threadIdx.x & 31u);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k = KT ? KT : p.k;
  const int slo = KT == 12 ? 8 : p.lo_bits;           // slice = key bits [slo, slo + sb)
  const int sb = KT == 12 ? 9 : p.slice_bits;
  const uint32_t NS = KT == 12 ? 512u : p.n_slices;
  const uint32_t smask = NS - 1u, lomask = (1u << slo) - 1u;
  const uint32_t cap = (uint32_t)FQ_BUCKET_ENTRIES / NS;  // bucket capacity per slice: a multiple of FQ_CHUNK
  uint32_t my_n = 0, my_g = 0;                           // copy-out in flight for slice tid
  bool flush_pending = false;
  uint32_t next_tile = 0;                                // thread 0: the tile claimed for the next iteration

  if (tid < FQ_MAXSLICES) S.fill[tid] = 0;
  if ((uint32_t)tid < NS) { S.qb16[tid] = (uint32_t)(__ldg(p.qbase + tid) / FQ_CHUNK); S.qcap[tid] = __ldg(p.qcap + tid); }
  if (tid < 48) S.raw[FQ_HALO + FQ_TB + tid] = 0;
  if (tid == 0) { S.tileq[0] = atomicAdd(p.counters, 1u); S.nbig = 0; S.umax = 0; S.usum = 0; S.cnt_next = 0; }
  const unsigned long long g_in = p.carry_in->s1.count;  // lines before the launch
  const unsigned long long pol_last_use = l2_policy_evict_first();
  __syncthreads();
  {  // the first tile's count is published here, every later one while the tile before it is processed
    const uint32_t first_tile = S.tileq[0];
    if (first_tile < p.n_tiles) fq_count_tile(p, first_tile, tid, 0, &S.cnt_next);
    __syncthreads();
    if (tid == 0) {
      if (first_tile < p.n_tiles) lookback_publish(p.tile_state, first_tile, S.cnt_next, g_in);
      S.cnt_next = 0;
      S.tileq[1] = atomicAdd(p.counters, 1u);
    }
    __syncthreads();
  }

  for (uint32_t it = 0;; ++it) {
    const uint32_t tile = S.tileq[it & 1], tile_next = S.tileq[(it + 1) & 1];
    if (tile >= p.n_tiles) break;
    const uint64_t t0 = (uint64_t)tile * FQ_TB;
    const int len = (int)((p.n - t0) < (uint64_t)FQ_TB ? (p.n - t0) : (uint64_t)FQ_TB);

    // the tile two grids ahead is pulled into L2 now: it is counted one tile time from now and processed after two
    if (tid == 0) {
      const uint64_t pt = (uint64_t)tile + 2u * gridDim.x;
      if (pt < p.n_tiles) {
        const uint64_t pb = pt * FQ_TB;
        const uint32_t pn = (uint32_t)((p.n - pb) < (uint64_t)FQ_TB ? (p.n - pb) : (uint64_t)FQ_TB) & ~15u;
        if (pn) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.data + pb), "r"(pn) : "memory");
      }
    }
    // ---- 1. load + newline census from the registers ------------------------------------------------------
    uint32_t m16[FQ_PIECES], cnt[FQ_PIECES];
#pragma unroll
    for (int q = 0; q < FQ_PIECES; ++q) {
      const int off = 16 * (q * FQ_NT + tid);
      uint4 x = make_uint4(0u, 0u, 0u, 0u);
      if (off < len) x = ldg_stream_hint(p.data + t0 + off, pol_last_use);
      if (len < FQ_TB && off + 16 > len) {  // last tile: bytes past the end read as 0
        uint32_t *xw = reinterpret_cast<uint32_t *>(&x);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int rem = len - (off + 4 * m);
          if (rem <= 0) xw[m] = 0u;
          else if (rem < 4) xw[m] &= (1u << (8 * rem)) - 1u;
        }
      }
      *reinterpret_cast<uint4 *>(S.raw + FQ_HALO + off) = x;
      m16[q] = nl_mask16(x);
      cnt[q] = __popc(m16[q]);
    }
    if (tid == FQ_NT - 1) {
      uint4 h = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);  // a launch without halo starts a line
      if (t0 > 0 || p.halo_ok) h = ldg_stream(p.data + t0 - 16);
      *reinterpret_cast<uint4 *>(S.raw) = h;
    }
    // two scans over packed pairs of counts (a field holds at most 16 * 32)
    uint32_t inc[FQ_PIECES];
    {
      const uint32_t ia = warp_incl_scan(cnt[0] | (cnt[1] << 16), lane);
      const uint32_t ib = warp_incl_scan(cnt[2] | (cnt[3] << 16), lane);
      inc[0] = ia & 0xFFFFu; inc[1] = ia >> 16; inc[2] = ib & 0xFFFFu; inc[3] = ib >> 16;
    }
    if (lane == 31) {
#pragma unroll
      for (int q = 0; q < FQ_PIECES; ++q) S.wtot_a[q * FQ_NW + w] = inc[q];
    }
    FQ_T(0);
    __syncthreads();  // (1) raw[] and wtot_a[] are complete
    FQ_T(1);
    bool claimed = false;

    // newline index of the first newline of every piece (pieces are ordered (q, thread))
    uint32_t base[FQ_PIECES];
    uint32_t N;
    {
      constexpr int VPL = (FQ_PIECES * FQ_NW + 31) / 32;
      uint32_t a[VPL], s = 0;
#pragma unroll
      for (int i = 0; i < VPL; ++i) { a[i] = VPL * lane + i < FQ_PIECES * FQ_NW ? S.wtot_a[VPL * lane + i] : 0u; s += a[i]; }
      const uint32_t sinc = warp_incl_scan(s, lane);
      N = __shfl_sync(0xffffffffu, sinc, 31);
      const uint32_t sex = sinc - s;
#pragma unroll
      for (int q = 0; q < FQ_PIECES; ++q) {
        const int idx = q * FQ_NW + w;
        uint32_t e = __shfl_sync(0xffffffffu, sex, idx / VPL);
#pragma unroll
        for (int i = 0; i < VPL - 1; ++i) {
          const uint32_t ai = __shfl_sync(0xffffffffu, a[i], idx / VPL);
          if (i < idx % VPL) e += ai;
        }
        base[q] = e + inc[q] - cnt[q];
      }
    }
    FQ_T(10);
    if (w == 0) {
      // the line the tile starts in: where did it begin?
      int head = -17;
      if (lane < 16 && S.raw[15 - lane] == '\n') head = -(lane + 1);
      const unsigned m = __ballot_sync(0xffffffffu, head != -17);
      if (m) head = -(__ffs(m));
      const unsigned long long g = lookback(p.tile_state, tile, N, g_in, lane);
      if (lane == 0) {
        S.G = g;
        S.head = head;
        S.nlpos[0] = (uint16_t)(head + FQ_BIAS);
        if (N + 1u <= (uint32_t)FQ_MAXSLOTS) S.nlpos[N + 1u] = (uint16_t)(len + FQ_BIAS);  // the last line ends with the tile
      }
    }

    // nlpos[j] = position of newline (lo - 1 + j), j = 0 .. FQ_MAXSLOTS; first batch straight from the registers
#pragma unroll
    for (int q = 0; q < FQ_PIECES; ++q) {
      uint32_t v = m16[q];
      uint32_t idx = base[q] + 1u;
      const uint32_t pos0 = (uint32_t)(16 * (q * FQ_NT + tid) + FQ_BIAS - 1);
      while (v) {
        const uint32_t b = (uint32_t)__ffs(v);
        v &= v - 1u;
        if (idx <= (uint32_t)FQ_MAXSLOTS) S.nlpos[idx] = (uint16_t)(pos0 + b);
        ++idx;
      }
    }

    // while warp 0 looks back, the other warps count the newlines of the NEXT tile of this CTA (its bytes are in L2)
    if (w != 0 && tile_next < p.n_tiles) fq_count_tile(p, tile_next, tid, 32, &S.cnt_next);

    // ---- 2./3. lines -> slots -> rows -> units, in batches of FQ_MAXSLOTS lines ------------------------------
    for (uint32_t lo = 0; lo < N + 1u; lo += FQ_MAXSLOTS) {
      if (lo) {
        __syncthreads();  // the previous batch is done with nlpos[] / uinfo[]
        // tiles with more than FQ_MAXSLOTS lines (rare): the newline masks are recomputed from shared memory
#pragma unroll
        for (int q = 0; q < FQ_PIECES; ++q) {
          if (cnt[q]) {
            const uint4 x = *reinterpret_cast<const uint4 *>(S.raw + FQ_HALO + 16 * (q * FQ_NT + tid));
            uint32_t v = nl_mask16(x);
            uint32_t idx = base[q];
            while (v) {
              const uint32_t b = (uint32_t)__ffs(v) - 1u;
              v &= v - 1u;
              const uint32_t j = idx + 1u - lo;  // wraps for idx + 1 < lo: rejected by the range test
              if (j <= (uint32_t)FQ_MAXSLOTS) S.nlpos[j] = (uint16_t)(16 * (q * FQ_NT + tid) + b + FQ_BIAS);
              ++idx;
            }
          }
        }
        if (tid == 0 && N + 1u - lo <= (uint32_t)FQ_MAXSLOTS) S.nlpos[N + 1u - lo] = (uint16_t)(len + FQ_BIAS);
        if (tid == 0) { S.umax = 0; S.usum = 0; }
      }
      FQ_T(2);
      __syncthreads();  // (2) nlpos[] of the batch, G and head are visible
      FQ_T(3);
      const unsigned long long G = S.G;
      if (tid == 0) {
        S.nbig = 0;
        if (lo == 0) {  // the next tile's count goes out one tile ahead of its processing
          if (tile_next < p.n_tiles) lookback_publish(p.tile_state, tile_next, S.cnt_next, g_in);
          S.cnt_next = 0;
        }
      }
      const uint32_t hi = (N + 1u < lo + FQ_MAXSLOTS) ? N + 1u : lo + FQ_MAXSLOTS;  // slots [lo, hi)
      const uint32_t jrow0 = (uint32_t)((1ull - (G + lo)) & 3ull);                  // first slot (batch relative) on phase 1
      const uint32_t NR = (hi - lo > jrow0) ? (hi - lo - jrow0 + 3u) / 4u : 0u;
      // tag.[0] <> '@' || tmp.[0] <> '+' (Files.ml:213); an empty tag / '+' line raises as well
      for (uint32_t j = tid; j < hi - lo; j += FQ_NT) {
        const unsigned long long L = G + lo + j;
        const uint32_t ph = (uint32_t)L & 3u;
        if ((ph & 1u) == 0u && L < p.max_lines) {
          const int s = (int)S.nlpos[j] - FQ_BIAS + 1;
          if (s >= 0 && s < len) {
            const uint8_t ch = S.raw[FQ_HALO + s];
            if (ch != (ph == 0 ? '@' : '+')) atomicMin(p.err_line, L);
          }
        }
      }
      FQ_T(11);
      // rows below jlim are inside max_lines (incomplete last record, -p cap)
      const unsigned long long Lb = G + lo;
      const uint32_t jlim = p.max_lines <= Lb ? 0u : (p.max_lines - Lb < (unsigned long long)(hi - lo) ? (uint32_t)(p.max_lines - Lb) : hi - lo);
      uint32_t nunits = 0, rinfo = 0;
      if ((uint32_t)tid < NR) {
        const uint32_t j = jrow0 + 4u * tid;
        const int pm = (int)S.nlpos[j] - FQ_BIAS, e = (int)S.nlpos[j + 1] - FQ_BIAS;
        const int a = pm + k > 0 ? pm + k : 0;  // first window end: line start + k - 1, inside the tile
        if (j < jlim && e > a) {
          nunits = (uint32_t)(e - a + FQ_W - 1) / FQ_W;
          rinfo = (uint32_t)a | ((uint32_t)e << 16);
        }
      }
      FQ_T(12);
      // Units are numbered row by row.  Reads of one length (the usual FASTQ) take the short way: every row gets
      // UPR = (longest row) unit numbers, unit q belongs to row q / UPR, and nothing is scanned or stored.
      if ((uint32_t)tid < ((NR + 31u) & ~31u)) {
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, nunits), wsum = __reduce_add_sync(0xffffffffu, nunits);
        if (lane == 0 && wsum) { atomicMax(&S.umax, wmax); atomicAdd(&S.usum, wsum); }
      }
      FQ_T(15);
      __syncthreads();  // (3)
      FQ_T(13);
      const uint32_t UPR = S.umax, usum = S.usum;
      const bool uniform = NR * UPR <= usum + (usum >> 2) + 64u;
      const uint32_t recip = UPR > 1u ? 0xFFFFFFFFu / UPR + 1u : 0u;  // q / UPR = umulhi(q, recip) for q < 2^16, UPR > 1
      uint32_t U = NR * UPR;
      if (!uniform) {
        const uint32_t ub = block_excl_scan(nunits, S.wtot_b, U, lane, w);
        if (nunits) {
          if (nunits <= (uint32_t)FQ_BIGROW) {
            for (uint32_t u = 0; u < nunits; ++u) S.uinfo[ub + u] = rinfo + u * FQ_W;
          } else {
            const uint32_t i = atomicAdd(&S.nbig, 1u);
            S.big[i].ub = ub; S.big[i].info = rinfo; S.big[i].n = nunits;
          }
        }
        FQ_T(14);
        __syncthreads();  // (4)
        const uint32_t nb = S.nbig;
        if (nb) {  // long lines: every thread fills its share
          for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t bub = S.big[b].ub, bi = S.big[b].info, bn = S.big[b].n;
            for (uint32_t u = tid; u < bn; u += FQ_NT) S.uinfo[bub + u] = bi + u * FQ_W;
          }
          __syncthreads();
        }
      }
      FQ_T(4);

      // the state the next launch starts from (only the tile that ends the launch)
      if (tid == 0 && tile == p.n_tiles - 1 && hi == N + 1u) {
        KpcStreamCarry co;
        co.s1.count = G + N;
        co.s1.last_hdr = 0;
        const int plast = (int)S.nlpos[N - lo] - FQ_BIAS;  // newline N - 1, or the head entry
        co.s1.last_nl = N ? p.abs_base + t0 + (uint64_t)plast + 1u : p.carry_in->s1.last_nl;
        co.kc.syms = 0; co.kc.n = 0; co.kc.closed = 1;
        const unsigned long long L = G + N;
        if ((L & 3ull) == 1ull && L < p.max_lines) {
          for (int pos = len - 1; pos > plast && pos >= -FQ_HALO && co.kc.n < (uint32_t)(k - 1); --pos) {
            const uint8_t sym = kpc_classify_dna(S.raw[FQ_HALO + pos]);
            if (sym == KPC_CLS_BREAK) break;
            co.kc.syms |= (uint64_t)sym << (2 * co.kc.n);
            co.kc.n++;
          }
        }
        co.last_byte = S.raw[FQ_HALO + len - 1];
        co.pad = 0;
        *p.carry_out = co;
      }

      // ---- 4./5. rounds of FQ_NT units ------------------------------------------------------------------------
#ifdef FQ_X_ROUND_REPS  // experiment: the rounds of every tile run several times (counts are multiplied): marginal cost of a round
      for (int rep = 0; rep < FQ_X_ROUND_REPS; ++rep)
#endif
      for (uint32_t q0 = 0; q0 < U; q0 += FQ_NT) {
        uint32_t ok = 0;
        uint32_t hi24 = 0, lo32 = 0, rlo = 0, rhi = 0;
        const uint32_t q = q0 + tid;
        uint32_t info = 0xFFFFu;  // p0 = 0xFFFF, e = 0: no unit
        if (q < U) {
          if (uniform) {
            const uint32_t r = UPR == 1u ? q : __umulhi(q, recip), u = q - r * UPR, j = jrow0 + 4u * r;
            const int pm = (int)S.nlpos[j] - FQ_BIAS, e = (int)S.nlpos[j + 1] - FQ_BIAS;
            const int a = (pm + k > 0 ? pm + k : 0) + (int)(u * FQ_W);
            if (j < jlim && a < e) info = (uint32_t)a | ((uint32_t)e << 16);
          } else {
            info = S.uinfo[q];
          }
        }
        if ((info & 0xFFFFu) < (info >> 16)) {
          const int p0 = (int)(info & 0xFFFFu), e = (int)(info >> 16);
          const int nvalid = e - p0 < FQ_W ? e - p0 : FQ_W;
          // bytes [p0 - 12, p0 + 16): 8 aligned words, funnel-shifted to 7
          const int A = FQ_HALO + p0 - FQ_CTX;
          const uint32_t *rw = reinterpret_cast<const uint32_t *>(S.raw) + (A >> 2);
          const uint32_t sh = (uint32_t)(A & 3) * 8u;
          uint32_t xw[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) xw[m] = rw[m];
          // Sequences.ml:52-58 + KMers.ml:272-277: x = (byte >> 1) & 3 maps A C T G (either case) to 0 1 2 3; a byte is a
          // base iff it equals the letter its x stands for (case bit ignored).  The 2-bit fields and the 0x80
          // "not a base" flags of four bytes are gathered with one dot product each.
          uint32_t xh = 0, xl = 0, iA = 0, iB = 0, iC = 0, iD = 0;
#pragma unroll
          for (int m = 0; m < 7; ++m) {
            const uint32_t wd = __funnelshift_r(xw[m], xw[m + 1], sh);
            const uint32_t x = (wd >> 1) & 0x03030303u;
            const uint32_t t = x | (x >> 4);
            const uint32_t sel = __byte_perm(t, 0u, 0x4420u);
            const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);  // 0->'A' 1->'C' 2->'T' 3->'G'
            const uint32_t a = ((wd & 0x5F5F5F5Fu) ^ expect) + 0x7F7F7F7Fu;
            const uint32_t nz = (a | wd) & 0x80808080u;                // 0x80 <=> the byte breaks k-mers
            if (m < 3) xh = __dp4a(x, 0x01041040u, xh << 8); else xl = __dp4a(x, 0x01041040u, xl << 8);
            const uint32_t wt = (m & 1) ? 0x01020408u : 0x10204080u;
            if (m < 2) iA = __dp4a(nz, wt, iA);
            else if (m < 4) iB = __dp4a(nz, wt, iB);
            else if (m < 6) iC = __dp4a(nz, wt, iC);
            else iD = __dp4a(nz, 0x01020408u, iD);
          }
          // A C T G -> A C G T on the packed fields (first base in the most significant bits)
          hi24 = xh ^ ((xh >> 1) & 0x00555555u);
          lo32 = xl ^ ((xl >> 1) & 0x55555555u);
          // validity bit (27 - i) for byte i; the dot products carry a factor 0x80
          const uint32_t inv28 = ((((iA << 8) + iB) >> 7) << 12) | (((iC << 4) + iD) >> 7);
          // bit b of okm <=> the k bases whose validity bits are b .. b + k - 1 are all valid
          uint32_t okm = ~inv28 & 0x0FFFFFFFu;
          {
            int have = 1;  // bit b of okm = AND of validity bits b .. b + have - 1
            while (2 * have <= k) { okm &= okm >> have; have *= 2; }
            if (have < k) okm &= okm >> (k - have);
          }
          ok = okm & 0xFFFFu;
          if (nvalid < FQ_W) ok &= ~((1u << (FQ_W - nvalid)) - 1u);
          // reverse complement of the 28 bases: base i (complemented) at bits 2i+1 : 2i
          if (DS) {
            const uint32_t nh = __brev(lo32), nlw = __brev(hi24);  // 64-bit reversal of hi24:lo32
            const uint32_t xl2 = __funnelshift_r(nlw, nh, 8), xh2 = nh >> 8;
            rlo = ~(((xl2 >> 1) & 0x55555555u) | ((xl2 & 0x55555555u) << 1));
            rhi = ~(((xh2 >> 1) & 0x55555555u) | ((xh2 & 0x55555555u) << 1)) & 0x00FFFFFFu;
          }
        }
        // the next tile is claimed as late as possible (tiles are published in claim order: an early claim makes every
        // later tile wait for this CTA), but early enough for the atomic to return before the tile ends
        if (!claimed && hi == N + 1u && q0 + FQ_NT >= U) {
          if (tid == 0) next_tile = atomicAdd(p.counters, 1u);
          claimed = true;
        }
        FQ_T(5);
        if (flush_pending) {  // the copy-out reserved after the previous round (its atomic has had time to return)
          fq_flush_copy(S, p, tid, cap, my_n, my_g, slo, sb, lomask);
          FQ_T(6);
          __syncthreads();    // (B) the buckets may be appended to again
          FQ_T(7);
        }
        // ---- 5. canonical keys, appended to the buckets as they are made ----------------------------------------
        // window end jw: forward k-mer = bits of hi24:lo32, reverse complement = bits of rhi:rlo, both moved to the top
        // of a 32-bit word (KMers.ml:364-368); min (KMers.ml:388) ignores the bits below because they only matter
        // when the k-mers are equal
        if (ok) {
          uint32_t pending = ok;
#pragma unroll
          for (int jw = 0; jw < FQ_W; ++jw) {
            uint32_t kk = fq_top_word(hi24, lo32, 34 + 2 * jw - 2 * k);
            if (DS) kk = min(kk, fq_top_word(rhi, rlo, 38 - 2 * jw));
            const uint32_t bit = 1u << (FQ_W - 1 - jw);
            if (KT == 12) {
              fq_append_k12(kk, pending, bit, s_fill, s_bucket, dummy_off, ok);
            } else if (pending & bit) {
              const uint32_t key = kk >> (32 - 2 * k);
              const uint32_t sl = fq_slice_of(key, slo, smask);
              const uint32_t pos = atoms_add(s_fill + 4u * sl, 1u);
              if (pos < cap) {
                sts_u16(s_bucket + 2u * (sl * (cap + FQ_BPAD) + pos), fq_bin_of(key, slo, sb, lomask));
                pending ^= bit;
              }
            }
          }
          // a slice whose bucket is full (skewed input): the k-mers that did not fit are counted in place
          while (pending) {
            const int b = 31 - __clz(pending);
            pending ^= 1u << b;
            const int jw = FQ_W - 1 - b;
            uint32_t kk = fq_top_word(hi24, lo32, 34 + 2 * jw - 2 * k);
            if (DS) kk = min(kk, fq_top_word(rhi, rlo, 38 - 2 * jw));
            atomicAdd(p.table + (kk >> (32 - 2 * k)), 1u);
          }
        }
        FQ_T(8);
        __syncthreads();  // (A) the appends of the round are complete
        FQ_T(9);
        fq_flush_reserve(S, p, tid, NS, cap, false, my_n, my_g);
        flush_pending = true;
      }
    }
    if (!claimed && tid == 0) next_tile = atomicAdd(p.counters, 1u);
    if (tid == 0) { S.tileq[it & 1] = next_tile; S.umax = 0; S.usum = 0; }  // slot of the tile just finished
    __syncthreads();  // the next tile overwrites raw[], nlpos[] and the scan scratch
  }
#ifdef FQ_PROFILE
  if ((threadIdx.x == 32 || threadIdx.x == 480) && blockIdx.x == 100) printf("FQPROF cta %d thread %d: load %u front %u rows %u classify %u copy %u append %u\n", blockIdx.x, threadIdx.x, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5]);
#endif
  // the CTA leaves: everything still in the buckets goes out, the last chunk of every slice padded
  if (flush_pending) fq_flush_copy(S, p, tid, cap, my_n, my_g, slo, sb, lomask);
  __syncthreads();
  fq_flush_reserve(S, p, tid, NS, cap, true, my_n, my_g);
  fq_flush_copy(S, p, tid, cap, my_n, my_g, slo, sb, lomask);
}

// one CTA per slice at a time: shared-memory histogram of the slice's queue, then RED of the non-zero bins
constexpr int FQ_CNT_NT = 1024;
__device__ __forceinline__ void fq_count2(uint32_t *tbl, uint32_t x) {
  const uint32_t a = x & 0xFFFFu, b = x >> 16;
  if (a != FQ_PAD) atomicAdd(&tbl[a], 1u);
  if (b != FQ_PAD) atomicAdd(&tbl[b], 1u);
}
__global__ void __launch_bounds__(FQ_CNT_NT, 1) fq_count_kernel(const KpcFqLaunch p) {
  uint32_t *tbl = reinterpret_cast<uint32_t *>(fq_smem_raw);
  __shared__ uint32_t s_slice;
  const int tid = threadIdx.x;
  const int lb = p.log_bins, slo = p.lo_bits, sb = p.slice_bits;
  const uint32_t lomask = (1u << slo) - 1u;
  const uint32_t nbins = 1u << lb;
  for (uint32_t i = tid; i < nbins; i += FQ_CNT_NT) tbl[i] = 0;
  // work items: whole slices while they fill complete waves of the grid; the slices of the last, partial wave are cut
  // into `parts` pieces so that it keeps every SM busy as well (512 slices on 148 SMs: 444 whole + 68 x 2 halves)
  const uint32_t G = gridDim.x, rem = p.n_slices % G;
  const uint32_t parts = rem ? (G / rem < 4u ? G / rem : 4u) : 1u;
  const uint32_t whole = p.n_slices - rem, n_items = whole + rem * parts;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_slice = atomicAdd(p.counters + 1, 1u);
    __syncthreads();
    const uint32_t item = s_slice;
    if (item >= n_items) break;
    const uint32_t b = item < whole ? item : whole + (item - whole) / parts;
    const uint32_t part = item < whole ? 0u : (item - whole) % parts, nparts = item < whole ? 1u : parts;
    uint32_t cn = p.qcursor[b];
    const uint32_t cap = p.qcap[b];
    if (cn > cap) cn = cap;  // both are multiples of FQ_CHUNK
    if (!cn) continue;
    const uint4 *sv = reinterpret_cast<const uint4 *>(p.queue + p.qbase[b]);
    const uint32_t nvec_all = cn >> 3;
    const uint32_t nvec = (uint32_t)((unsigned long long)nvec_all * (part + 1u) / nparts);
    uint32_t v = (uint32_t)((unsigned long long)nvec_all * part / nparts) + tid;
    for (; v + 3u * FQ_CNT_NT < nvec; v += 4u * FQ_CNT_NT) {
      uint4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = ldg_stream(reinterpret_cast<const uint8_t *>(sv + v + u * FQ_CNT_NT));
#pragma unroll
      for (int u = 0; u < 4; ++u) { fq_count2(tbl, x[u].x); fq_count2(tbl, x[u].y); fq_count2(tbl, x[u].z); fq_count2(tbl, x[u].w); }
    }
    for (; v < nvec; v += FQ_CNT_NT) {
      const uint4 x = ldg_stream(reinterpret_cast<const uint8_t *>(sv + v));
      fq_count2(tbl, x.x); fq_count2(tbl, x.y); fq_count2(tbl, x.z); fq_count2(tbl, x.w);
    }
    __syncthreads();
    for (uint32_t i = tid; i < nbins; i += FQ_CNT_NT) {
      const uint32_t cv = tbl[i];
      if (cv) { atomicAdd(p.table + fq_key_of(b, i, slo, sb, lomask), cv); tbl[i] = 0; }
    }
  }
}

}  // namespace

uint32_t kpc_fq_tile_bytes() { return FQ_TB; }
// bins per slice: at most 2^15 (one u16 queue entry, a 128 KiB shared-memory table) and at least 128 slices so
// that the counting kernel has enough CTAs; the slice index is cut out of the middle of the key
int kpc_fq_log_bins(int k) { return 2 * k - 7 < 15 ? 2 * k - 7 : 15; }
int kpc_fq_lo_bits(int k) { return k == 12 ? 8 : kpc_fq_log_bins(k) / 2; }  // k = 12: byte-aligned fields (fq_append_k12)
uint32_t kpc_fq_queue_slack() { return FQ_CHUNK * 512u; }  // padding entries per slice: < one chunk per CTA (<= 2 per SM)
bool kpc_fq_supported(int k, int content) {
  return (content == KPC_CONTENT_DNA_SS || content == KPC_CONTENT_DNA_DS) && k >= 4 && k <= 12;
}

static cudaStream_t fq_cs(rt_stream s) { return (cudaStream_t)rt_stream_native(s); }

// ---- optional timing: one event triple per launch (before partition, between, after count) -------------------------
namespace {
struct FqTiming {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // 3 per launch
  size_t used = 0;              // launches recorded since the last read
  unsigned long long bytes = 0;
} g_fqt;
cudaEvent_t fq_timing_event(size_t i) {
  while (g_fqt.ev.size() <= i) {
    cudaEvent_t e;
    FQ_CUDA_CHECK(cudaEventCreate(&e));
    g_fqt.ev.push_back(e);
  }
  return g_fqt.ev[i];
}
}  // namespace
void kpc_fq_timing_enable(bool on) { g_fqt.on = on; g_fqt.used = 0; g_fqt.bytes = 0; }
void kpc_fq_timing_read(double *partition_ms, double *count_ms, unsigned long long *launches, unsigned long long *bytes) {
  double a = 0, b = 0;
  for (size_t i = 0; i < g_fqt.used; ++i) {
    float t = 0;
    FQ_CUDA_CHECK(cudaEventSynchronize(g_fqt.ev[3 * i + 2]));
    FQ_CUDA_CHECK(cudaEventElapsedTime(&t, g_fqt.ev[3 * i], g_fqt.ev[3 * i + 1])); a += t;
    FQ_CUDA_CHECK(cudaEventElapsedTime(&t, g_fqt.ev[3 * i + 1], g_fqt.ev[3 * i + 2])); b += t;
  }
  *partition_ms = a; *count_ms = b; *launches = g_fqt.used; *bytes = g_fqt.bytes;
  g_fqt.used = 0; g_fqt.bytes = 0;
}

template <bool DS, int KT>
static void launch_partition(const KpcFqLaunch &L, rt_stream s) {
  auto kern = fq_partition_kernel<DS, KT>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    FQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FqSmem)));
    FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, FQ_NT, sizeof(FqSmem)));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (const char *e = getenv("KPC_FQ_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < blocks_per_sm) blocks_per_sm = v; }  // experiments
  }
  long long grid = (long long)rt_sm_count() * blocks_per_sm;
  if (grid > (long long)L.n_tiles) grid = L.n_tiles;
  if (grid < 1) return;
  kern<<<(unsigned)grid, FQ_NT, sizeof(FqSmem), fq_cs(s)>>>(L);
  FQ_CUDA_CHECK(cudaGetLastError());
}

void kpc_fq_partition(const KpcFqLaunch &L, rt_stream s) {
  if (!kpc_fq_supported(L.k, L.content) || L.n_slices > (uint32_t)FQ_MAXSLICES || L.n_slices < 1 ||
      (L.n_slices & (L.n_slices - 1)) || (FQ_BUCKET_ENTRIES / L.n_slices) % FQ_CHUNK || FQ_BUCKET_ENTRIES / L.n_slices < 2 * FQ_CHUNK ||
      (1u << L.slice_bits) != L.n_slices || L.lo_bits + L.slice_bits > 2 * L.k)
    throw KpcError(KPC_E_STATE, "internal: fast FASTQ path asked for an unsupported configuration");
  const bool ds = L.content == KPC_CONTENT_DNA_DS;
  if (g_fqt.on) FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used), fq_cs(s)));
  if (L.k == 12) { if (ds) launch_partition<true, 12>(L, s); else launch_partition<false, 12>(L, s); }
  else { if (ds) launch_partition<true, 0>(L, s); else launch_partition<false, 0>(L, s); }
  if (g_fqt.on) FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used + 1), fq_cs(s)));
}

void kpc_fq_count(const KpcFqLaunch &L, rt_stream s) {
  static bool attr = false;
  if (!attr) {
    FQ_CUDA_CHECK(cudaFuncSetAttribute(fq_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << 15));
    attr = true;
  }
  long long grid = rt_sm_count();
  if (grid > (long long)L.n_slices) grid = L.n_slices;
  fq_count_kernel<<<(unsigned)grid, FQ_CNT_NT, (size_t)4 << L.log_bins, fq_cs(s)>>>(L);
  FQ_CUDA_CHECK(cudaGetLastError());
  if (g_fqt.on) {
    FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used + 2), fq_cs(s)));
    g_fqt.used++;
    g_fqt.bytes += L.n;
  }
}
