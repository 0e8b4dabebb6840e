// kpc_partition.cuh -- the partition kernel of the dense-table pipeline (reads -> canonical k-mers -> per-slice queues).
//
// Reference semantics restated here (paths relative to the KPop tree):
//   Files.FASTQ.iter_se        BiOCamLib/lib/Files.ml:201-221   4-line records, '@' / '+' checks, only line 2 is sequence
//   Sequences.Lint.dnaize      BiOCamLib/lib/Sequences.ml:41-67  [ACGTacgt] are bases, every other byte breaks k-mers
//   DNAHash*.iteri / iterc     BiOCamLib/lib/KMers.ml:319-349, 357-389   first base most significant, key = min f rc
//   IntHashFrequencies.add     BiOCamLib/lib/KMers.ml:107-111   count[key] += 1 (done by fq_count_kernel from the queues)
//
// One CTA of NT threads owns one tile of TB = 16 * VPT * NT bytes at a time (claimed in stream order):
//   0. the tile (and the 16 bytes before it) arrives in shared memory by ONE bulk copy (cp.async.bulk, completion on an
//      mbarrier) that thread 0 started as soon as the previous tile's bytes had been read for the last time
//   1. census, taken one tile AHEAD straight from L2: every thread owns 16 * VPT consecutive bytes of the tile; the loads
//      are 256 bits wide and laid out so that a warp's load instruction covers whole 128-byte lines (lane L takes sector
//      L, the per-sector masks reach their owners by shuffles): what such scattered accesses cost is the number of
//      sector requests, not the bytes.  Three integer operations per word flag the line feeds, dot products gather the
//      flags into a bit mask that stays in registers; one warp scan + one 16-entry scan rank them
//   2. warp 0 publishes the tile's line-feed count and gets the number of line feeds before the tile by a decoupled
//      look-back over one word per tile (the other warps use the time for the bucket copy-out that is still pending)
//   3. every line feed knows its line number: the thread that holds it writes the start / the end of the sequence
//      line next to it straight into the row table and checks the first byte of header / '+' lines (Files.ml:213)
//   4. rows are cut into units of W window-end positions, numbered by a division when all rows have (about) the same
//      length and by a block scan + unit table otherwise
//   5. one thread per unit: 12 + 16 bytes -> 2-bit codes + "not a base" flags with SIMD-in-register arithmetic and
//      dot-product gathers, reverse complement of the 28 bases from two BREVs; per window the forward and the
//      reverse-complement k-mer are funnel-shifted to the top of a word, unsigned min is the canonical key, its middle
//      bits name a slice, and the key is appended to the slice's bucket in shared memory (one predicated ATOMS + one
//      predicated STS.U16, no branch)
//   6. software write-combining: after every round the thread that owns a slice reserves whole 32-byte chunks of its
//      bucket in the slice's queue in HBM (one global atomic, issued a round before its result is needed) and copies
//      them out, one 256-bit store (a whole sector) per chunk.  A bucket or queue that overflows (skewed input) sends its keys to the global
//      table with RED, so the result is exact for any input.
//
// The source is compiled by nvcc for sm_100a (the product) and by g++ against tests/emul/simt_emul.h (KPC_SIMT_EMUL,
// test infrastructure: small geometries, fuzzed against the oracle on the CPU).
#pragma once
#include <stddef.h>

#include "kpc_fastq.h"
#include "kpc_simt.h"

// tuning switches (defaults = what measured best on a B200; tools/gpu_define_variants.sh rebuilds with -D overrides)
#ifndef FQ_APPEND_GROUP
#define FQ_APPEND_GROUP 1   // shared-memory atomics issued back to back per group of windows (1, 2 or 4)
#endif
#ifndef FQ_PADFILL
#define FQ_PADFILL 0        // 1: windows that are not valid take a bucket position that already reads as padding; 0: they
#endif                      //    count into per-lane dummy counters instead (one SEL per window, no padding in the queues)
#ifndef FQ_STG256
#define FQ_STG256 1         // bucket chunks leave with one 256-bit store (a whole 32-byte sector) instead of two 128-bit ones
#endif
#ifndef FQ_LDG256
#define FQ_LDG256 1         // census loads of 32 bytes (a whole sector) per thread instead of 16
#endif
#ifndef FQ_L2HINTS
#define FQ_L2HINTS 1        // census loads evict_last, the tile's bulk copy evict_first
#endif
// measurement-only builds (never the product): FQ_PHASE_CLOCKS adds clock() probes at the phase boundaries
// (tools/gpu_phase_clocks.sh); FQ_X_NOATOM / FQ_X_NOSTG drop the queue atomics / the queue stores to see what they cost --
// the results are then WRONG, only the timing means something
#ifdef FQ_PHASE_CLOCKS
__device__ unsigned long long g_fq_phase[16];
#define FQ_PROBE(i) do { if (tid == 32) { const uint32_t c_ = (uint32_t)clock(); S.dbg[i] += c_ - S.dbg_last; S.dbg_last = c_; } } while (0)
#else
#define FQ_PROBE(i) do { } while (0)
#endif
constexpr int FQ_W = 16;                          // window-end positions per unit (one thread)
constexpr int FQ_CTX = 12;                        // context bytes loaded before a unit (>= k - 1)
constexpr int FQ_HALO = 16;
constexpr int FQ_BIGROW = 32;                     // rows with more units are expanded by the whole CTA
constexpr int FQ_MAXSLICES = 512;
#ifndef FQ_BUCKET_ENTRIES_V
#define FQ_BUCKET_ENTRIES_V 24576
#endif
constexpr int FQ_BUCKET_ENTRIES = FQ_BUCKET_ENTRIES_V;  // shared-memory bucket space (u16 entries) shared by all slices
constexpr int FQ_BPAD = 8;                        // entries between two buckets: a stride of cap + 8 entries (28 or 100 words) keeps the owners' 128-bit accesses free of bank conflicts
constexpr uint32_t FQ_CAP12 = FQ_BUCKET_ENTRIES / FQ_MAXSLICES;   // k = 12 always runs with FQ_MAXSLICES slices: bucket capacity
constexpr uint32_t FQ_STRIDE12 = (FQ_CAP12 + FQ_BPAD) / 2u;       // and bucket stride in 32-bit words (sl4 = 4 * slice)
constexpr int FQ_CHUNK = 16;                      // entries per copy-out chunk (32 bytes: one L2 sector)
constexpr uint16_t FQ_PAD = 0xFFFFu;              // queue entry that pads the last chunk of a CTA (skipped by fq_count)
constexpr int FQ_PBIAS = 16;                      // row starts are stored + FQ_PBIAS (they begin at -16)

template <int NT_, int VPT_>
struct FqGeom {
  static constexpr int NT = NT_, VPT = VPT_, NW = NT_ / 32;
  static constexpr int SEG = 16 * VPT_;           // bytes per thread in the census
  static constexpr int TB = SEG * NT_;            // tile bytes
  static constexpr int MAXROWS = NT_;             // sequence lines per batch
  static constexpr int MAXUNITS = TB / FQ_W + MAXROWS + 8;  // units per batch: sum of ceil(len / W) over its rows
  static constexpr int MAXBIG = TB / (FQ_BIGROW * FQ_W) + 2;
  static constexpr int SPT = (FQ_MAXSLICES + NT_ - 1) / NT_;  // slices per owner thread
  static_assert(NT_ % 32 == 0 && NT_ <= 1024, "whole warps");
  static_assert(VPT_ == 1 || VPT_ == 2 || VPT_ == 4, "the census mask is 64 bits");
  static_assert(TB + FQ_PBIAS < 65536, "positions are 16-bit");
};

struct FqBigRow { uint32_t ub, info, n; };
template <class G>
struct FqSmemT {
  alignas(128) uint8_t raw[FQ_HALO + G::TB + 48];  // raw[16 + i] = tile byte i
  alignas(16) uint16_t bucket[FQ_BUCKET_ENTRIES + FQ_BPAD * FQ_MAXSLICES + 2 * FQ_CHUNK];  // slice s owns [s * (cap + FQ_BPAD), + cap)
  uint32_t fill[FQ_MAXSLICES];                    // entries in the bucket (may run past cap while appending)
#ifdef FQ_PHASE_CLOCKS
  unsigned long long dbg[16]; uint32_t dbg_last;  // experiment: cycles per phase, as seen by one thread of warp 1
#endif
  uint32_t dummy[32];                             // FQ_PADFILL == 0: where the appends of windows that are not valid count
  uint32_t uinfo[G::MAXUNITS];                    // unit -> first window end (low half) | end of its line (high half)
  uint16_t rowS[G::MAXROWS + 2], rowE[G::MAXROWS + 2];  // first byte / line feed of every sequence line, + FQ_PBIAS
  uint32_t wtot[32], wtot2[64];                   // warp totals: block scans / census of the next tile (two slots)
  FqBigRow big[G::MAXBIG];
  struct alignas(16) { int head; uint32_t jrow0, NRt, jlim; } ti;  // framing constants of the tile (warp 0 -> everybody)
  unsigned long long G_;                          // number of '\n' in the stream before the tile
  unsigned long long mbar;                        // completion of the tile's bulk copy
  uint32_t tileq[2];
  uint32_t nbig;
  uint32_t umax, usum;                            // longest row of the batch (units) and the sum over its rows
  uint32_t recip[65];                             // 2^32 / u + 1 for u = 2 .. 64 (q / u = umulhi(q, recip[u]), q < 2^16)
};

KP_DEV uint32_t fq_warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// exclusive prefix over the CTA (thread order), one barrier; wtot must not be in use by a slower warp
template <int NW>
KP_DEV uint32_t fq_block_excl_scan(uint32_t v, uint32_t *wtot, uint32_t &total, int lane, int w) {
  uint32_t inc = fq_warp_incl_scan(v, lane);
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  uint32_t t = lane < NW ? wtot[lane] : 0u;
  uint32_t tinc = fq_warp_incl_scan(t, lane);
  total = __shfl_sync(0xffffffffu, tinc, 31);
  uint32_t wex = __shfl_sync(0xffffffffu, tinc - t, w);
  return wex + inc - v;
}
// 0x80 in every byte of w that equals '\n' (exact, three operations)
KP_DEV uint32_t fq_nl_mask(uint32_t w) {
#ifdef KPC_SIMT_EMUL
  uint32_t t = (w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
  t += 0x7F7F7F7Fu;
  return ~(t | w) & 0x80808080u;
#else
  // the two-constant logic operations are written as LOP3 so that ptxas keeps one constant in a uniform register
  uint32_t t, z;
  asm("lop3.b32 %0, %1, 0x0A0A0A0A, 0x7F7F7F7F, 0x28;" : "=r"(t) : "r"(w));   // (w ^ 0x0A..) & 0x7F..
  t += 0x7F7F7F7Fu;
  asm("lop3.b32 %0, %1, %2, 0x80808080, 0x02;" : "=r"(z) : "r"(t), "r"(w));    // ~(t | w) & 0x80..
  return z;
#endif
}
// bit i of the result <=> byte i of the 16-byte vector is '\n': the 0x80 flags are gathered with dot products
KP_DEV uint32_t fq_nl_mask16(const uint4 &x) {
  uint32_t lo = __dp4a(fq_nl_mask(x.x), 0x08040201u, 0u);
  lo = __dp4a(fq_nl_mask(x.y), 0x80402010u, lo);
  uint32_t hi = __dp4a(fq_nl_mask(x.z), 0x08040201u, 0u);
  hi = __dp4a(fq_nl_mask(x.w), 0x80402010u, hi);
  return (lo + (hi << 8)) >> 7;
}

// slice / bin of a key and back: slice = key bits [lo, lo + sb), bin = the other bits packed together
KP_DEV uint32_t fq_slice_of(uint32_t key, int lo, uint32_t smask) { return (key >> lo) & smask; }
KP_DEV uint32_t fq_bin_of(uint32_t key, int lo, int sb, uint32_t lomask) {
  return ((key >> (lo + sb)) << lo) | (key & lomask);
}
KP_DEV uint32_t fq_key_of(uint32_t slice, uint32_t bin, int lo, int sb, uint32_t lomask) {
  return ((bin >> lo) << (lo + sb)) | (slice << lo) | (bin & lomask);
}

// ---- number of '\n' before the tile: decoupled look-back over one word per tile (2 status bits + 62 value bits) -------
// All CTAs work on neighbouring tiles at the same time, so the nearest tile with an inclusive count is usually about
// one grid back: the warp loads FQ_LB_ROUNDS x 32 states in one go (the loads overlap) before it looks at any of them,
// which makes the common case one L2 round trip.
constexpr unsigned long long FQ_ST_AGG = 1ull << 62, FQ_ST_INC = 2ull << 62, FQ_VAL = (1ull << 62) - 1ull;
constexpr int FQ_LB_ROUNDS = 10;
KP_DEV void fq_lookback_publish(unsigned long long *state, uint32_t tile, uint32_t total, unsigned long long g_in) {
  kp_st_relaxed_u64(state + tile, tile == 0 ? (FQ_ST_INC | (g_in + total)) : (FQ_ST_AGG | (unsigned long long)total));
}
KP_DEV unsigned long long fq_lookback(unsigned long long *state, uint32_t tile, uint32_t total, unsigned long long g_in,
                                      int lane) {
  if (tile == 0) return g_in;
  unsigned long long acc = 0;
  long long j0 = (long long)tile - 1;
  for (;;) {
    unsigned long long v[FQ_LB_ROUNDS];
#pragma unroll
    for (int i = 0; i < FQ_LB_ROUNDS; ++i) {
      const long long j = j0 - 32 * i - lane;
      if (j >= 0) v[i] = kp_ld_relaxed_u64(state + j);
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
      const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
      const unsigned need = first_inc >= 31 ? 0xffffffffu : ((2u << first_inc) - 1u);
      if (inv_mask & need) { stale = true; continue; }  // a predecessor has not published yet: reload from here
      part += ((need >> lane) & 1u) ? (v[i] & FQ_VAL) : 0ull;
      if (first_inc < 32) done = true; else j0 -= 32;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    acc += part;
    if (done) break;
    if (stale) KP_SPIN_YIELD();
  }
  if (lane == 0) kp_st_relaxed_u64(state + tile, FQ_ST_INC | (acc + total));
  return acc;
}

// ---- bucket copy-out (step 6) ------------------------------------------------------------------------------------------
// Thread t owns slices t, t + NT, ...  reserve: called after the barrier that follows a round of appends; whole chunks
// of the bucket are reserved in the slice's queue with one global atomic whose result is not needed before the copy
// (one round of k-mer arithmetic later).  copy: the owner moves the reserved chunks with 128-bit accesses and the
// bucket's remainder (< one chunk) to the front; a barrier must follow before the buckets are appended to again.
template <class G>
struct FqOwner {
  uint32_t n[G::SPT], g[G::SPT];  // copy-out in flight: entries reserved, position in the queue
  uint32_t qb16[G::SPT], qcap[G::SPT];
};
template <class G>
KP_DEV void fq_flush_reserve(FqSmemT<G> &S, const KpcFqLaunch &p, FqOwner<G> &own, int tid, uint32_t NS, uint32_t cap,
                             bool final) {
#pragma unroll
  for (int i = 0; i < G::SPT; ++i) {
    const uint32_t s = (uint32_t)tid + (uint32_t)(i * G::NT);
    own.n[i] = 0; own.g[i] = 0;
    if (s < NS) {
      uint32_t f = S.fill[s];
      if (f > cap) f = cap;
      // whole chunks; the CTA is leaving: the last chunk goes out as it is (the entries behind the fill are FQ_PAD)
      const uint32_t n = final ? (f + FQ_CHUNK - 1u) & ~(uint32_t)(FQ_CHUNK - 1) : f & ~(uint32_t)(FQ_CHUNK - 1);
#if !FQ_PADFILL
      if (final) for (uint32_t e = f; e < n; ++e) S.bucket[s * (cap + FQ_BPAD) + e] = FQ_PAD;
#endif
      S.fill[s] = f > n ? f - n : 0u;
      own.n[i] = n;
#ifdef FQ_X_NOATOM
      if (n) own.g[i] = 0;
#else
      if (n) own.g[i] = atomicAdd(p.qcursor + s, n);
#endif
    }
  }
}
// Invariant between rounds: every bucket entry at or behind the slice's fill is FQ_PAD (windows that are not valid take
// a position but store nothing, so their positions must already read as padding).
template <class G>
KP_DEV void fq_flush_copy(FqSmemT<G> &S, const KpcFqLaunch &p, const FqOwner<G> &own, int tid, uint32_t cap, int lo, int sb,
                          uint32_t lomask) {
#pragma unroll
  for (int i = 0; i < G::SPT; ++i) {
    const uint32_t my_n = own.n[i], my_g = own.g[i];
    if (!my_n) continue;
    const uint32_t s = (uint32_t)tid + (uint32_t)(i * G::NT);
    const uint32_t qc = own.qcap[i];
    uint4 *src = reinterpret_cast<uint4 *>(S.bucket + s * (cap + FQ_BPAD));
    uint4 *dst = reinterpret_cast<uint4 *>(p.queue + (unsigned long long)own.qb16[i] * FQ_CHUNK + my_g);
    const uint4 pad4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    // the remainder (< one chunk) moves to the front
    uint4 r0 = pad4, r1 = pad4;
    if (my_n < cap) { r0 = src[my_n >> 3]; r1 = src[(my_n >> 3) + 1]; }
    for (uint32_t c = 0; c < my_n; c += FQ_CHUNK) {
      const uint4 v0 = src[c >> 3], v1 = src[(c >> 3) + 1];
      if (my_g + c + FQ_CHUNK <= qc) {
#ifndef FQ_X_NOSTG
#if FQ_STG256
        kp_stg_256(dst + (c >> 3), v0, v1);  // one whole sector per store
#else
        dst[c >> 3] = v0;
        dst[(c >> 3) + 1] = v1;
#endif
#else
        if (v0.x == 0x12345u && v1.y == 0x54321u) dst[0] = v0;
#endif
      } else {  // queue full: count in place
        const uint32_t ww[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const uint32_t en = (ww[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
          if (en != FQ_PAD) atomicAdd(p.table + fq_key_of(s, en, lo, sb, lomask), 1u);
        }
      }
#if FQ_PADFILL
      src[c >> 3] = pad4;  // what has been copied reads as padding again
      src[(c >> 3) + 1] = pad4;
#endif
    }
#if FQ_PADFILL
    if (my_n < cap) { src[my_n >> 3] = pad4; src[(my_n >> 3) + 1] = pad4; }
#endif
    src[0] = r0;
    src[1] = r1;
  }
}

// ---- one k-mer: slice, bucket append (step 5) ----------------------------------------------------------------------
// kk holds the canonical k-mer in its TOP 2k bits (the bits below are ignored).  `ok & bit` says whether the window is
// valid; the bit is cleared in `pending` once the key sits in its bucket.  k = 12: slice = key bits [8, 17), the queue
// entry is key[0, 8) | key[17, 24) << 8 (one PRMT); everything is predicated, there is no branch.
// (moved below fq_top_word)
// top 32 bits of (hi:lo) << s, 0 <= s < 64
KP_DEV uint32_t fq_top_word(uint32_t hi, uint32_t lo, int s) {
  if (s >= 32) return lo << (s - 32);
  if (s == 0) return hi;
  return __funnelshift_l(lo, hi, s);
}

// Four windows jw0 .. jw0 + 3 of a unit, k = 12: the four shared-memory atomics are issued together so that their
// latencies overlap.  kk holds the canonical k-mer in its TOP 24 bits (the bits below are ignored): slice = key bits
// [8, 17), the queue entry is key[0, 8) | key[17, 24) << 8 (one PRMT).  A window that is not valid takes a position in
// some bucket like any other but stores nothing there: the position keeps its FQ_PAD (see fq_flush_copy).  The bit of
// a valid window is cleared in `pending` once the key sits in its bucket.
template <bool DS, class SmemT>
KP_DEV void fq_append4_k12(SmemT &S, uint32_t hi24, uint32_t lo32, uint32_t rhi, uint32_t rlo, int jw0, uint32_t ok,
                           uint32_t &pending, uint32_t dummy_off) {
  (void)dummy_off;
  constexpr int GR = FQ_APPEND_GROUP;
  uint32_t kk[GR], sl4[GR], pos[GR];
#pragma unroll
  for (int i = 0; i < GR; ++i) {
    const int jw = jw0 + i;
    kk[i] = fq_top_word(hi24, lo32, 10 + 2 * jw);
    if (DS) kk[i] = kp_umin(kk[i], fq_top_word(rhi, rlo, 38 - 2 * jw));
    sl4[i] = (kk[i] >> 14) & 0x7FCu;                              // 4 * slice
  }
#pragma unroll
  for (int i = 0; i < GR; ++i) {
#if FQ_PADFILL
    pos[i] = atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(S.fill) + sl4[i]), 1u);
#else
    const bool valid = (ok & (1u << (FQ_W - 1 - (jw0 + i)))) != 0u;
    pos[i] = atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(S.fill) + (valid ? sl4[i] : dummy_off)), 1u);
#endif
  }
#pragma unroll
  for (int i = 0; i < GR; ++i) {
    const uint32_t bit = 1u << (FQ_W - 1 - (jw0 + i));
    const uint32_t en = __byte_perm(kk[i], kk[i] >> 1, 0x0071u);  // key[0, 8) | key[17, 24) << 8
    if ((ok & bit) != 0u && pos[i] < FQ_CAP12) {
      *reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(S.bucket) + sl4[i] * FQ_STRIDE12 + pos[i] * 2u) = (uint16_t)en;
      pending ^= bit;
    }
  }
}

// Sequences.ml:52-58 + KMers.ml:272-277 on four bytes at once: x = (byte >> 1) & 3 maps A C T G (either case) to
// 0 1 2 3; a byte is a base iff, with the case bit and the two code bits masked out, it reads 0x41 (A, C, G) or 0x50
// (T: the bytes whose code bits are 1,0).  Returns the four 2-bit codes (one per byte) and sets 0x80 in `nz` for
// every byte that breaks k-mers.
KP_DEV uint32_t fq_classify4(uint32_t wd, uint32_t &nz) {
  const uint32_t s1 = wd >> 1;
  const uint32_t x = s1 & 0x03030303u;
  const uint32_t g = (wd >> 2) & ~s1 & 0x01010101u;       // T / t
  const uint32_t target = g * 0x0Fu + 0x41414141u;         // 0x41 or 0x50 per byte
  const uint32_t e1 = (wd & 0x59595959u) ^ target;         // 0 <=> the byte is a base (bit 7 is handled below)
  const uint32_t a = e1 + 0x7F7F7F7Fu;
  nz = (a | wd) & 0x80808080u;
  return x;
}

// ---- newline census of one thread's SEG bytes of a tile, straight from global memory (L2) ------------------------------
// Taken one tile AHEAD of the tile's processing: the masks wait in registers, the tile's line-feed count is published
// right away, and by the time the tile looks back every predecessor has published its own (nobody waits for anybody).
KP_DEV uint32_t fq_keep_bytes(uint32_t w, int n) {  // the first n bytes of the word, the others zero
  return n <= 0 ? 0u : (n < 4 ? w & ((1u << (8 * n)) - 1u) : w);
}
template <class G>
struct FqCensus {
  uint32_t mlo, mhi;  // bit b <=> byte SEG * tid + b of the tile is '\n'
  uint32_t rank0;     // line feeds of the tile before this thread's bytes
};
// Returns true when x holds the warp-interleaved layout: lane L has bytes [32 L, 32 L + 32) of the warp's first and of
// its second 1024 bytes (every load instruction covers whole 128-byte lines) instead of its own SEG bytes.
template <class G>
KP_DEV bool fq_census_load(const KpcFqLaunch &p, uint32_t tile, int tid, unsigned long long policy, uint4 (&x)[G::VPT]) {
  const uint64_t t0 = (uint64_t)tile * G::TB;
  const int len = (int)((p.n - t0) < (uint64_t)G::TB ? (p.n - t0) : (uint64_t)G::TB);
#if FQ_LDG256
  // whole tile, launch on a sector boundary (device inputs only promise 16 bytes): 32 bytes (one sector) per load
  if (G::VPT >= 2 && len == G::TB && ((unsigned long long)(size_t)p.data & 31ull) == 0ull) {
    if (G::VPT == 4) {
      const uint8_t *wb = p.data + t0 + (size_t)(G::SEG * 32) * (size_t)(tid >> 5) + 32 * (tid & 31);
      kp_ldg_stream_hint_256(wb, policy, x[0], x[1]);
      kp_ldg_stream_hint_256(wb + 1024, policy, x[G::VPT == 4 ? 2 : 0], x[G::VPT == 4 ? 3 : 0]);
      return true;
    } else {
#pragma unroll
      for (int c = 0; c + 1 < G::VPT; c += 2)
        kp_ldg_stream_hint_256(p.data + t0 + G::SEG * tid + 16 * c, policy, x[c], x[c + 1 < G::VPT ? c + 1 : c]);
      return false;
    }
  }
#endif
#pragma unroll
  for (int c = 0; c < G::VPT; ++c) {
    const int off = G::SEG * tid + 16 * c;
    x[c] = make_uint4(0u, 0u, 0u, 0u);
    if (off < len) x[c] = kp_ldg_stream_hint(p.data + t0 + off, policy);
    if (len < G::TB && off + 16 > len) {  // last tile: bytes past the end read as 0
      x[c].x = fq_keep_bytes(x[c].x, len - off);
      x[c].y = fq_keep_bytes(x[c].y, len - off - 4);
      x[c].z = fq_keep_bytes(x[c].z, len - off - 8);
      x[c].w = fq_keep_bytes(x[c].w, len - off - 12);
    }
  }
  return false;
}
// masks + warp scan; lane 31 leaves the warp's total in wtot[w]; returns the inclusive count of the thread
template <class G>
KP_DEV uint32_t fq_census_masks(const uint4 (&x)[G::VPT], bool interleaved, FqCensus<G> &c, uint32_t *wtot, int lane, int w) {
  uint32_t m16[G::VPT];
#pragma unroll
  for (int i = 0; i < G::VPT; ++i) m16[i] = fq_nl_mask16(x[i]);
  c.mlo = m16[0]; c.mhi = 0;
  if (G::VPT >= 2) c.mlo |= m16[G::VPT >= 2 ? 1 : 0] << 16;
  if (G::VPT == 4) c.mhi = m16[G::VPT == 4 ? 2 : 0] | (m16[G::VPT == 4 ? 3 : 0] << 16);
  if (G::VPT == 4 && interleaved) {
    // c.mlo / c.mhi are the masks of sector `lane` of the warp's first / second 1024 bytes; this thread's 64 bytes are
    // sectors 2 lane and 2 lane + 1 of the warp's 2048 bytes
    const int s0 = (2 * lane) & 31, s1 = (2 * lane + 1) & 31;
    const uint32_t a0 = __shfl_sync(0xffffffffu, c.mlo, s0), a1 = __shfl_sync(0xffffffffu, c.mhi, s0);
    const uint32_t b0 = __shfl_sync(0xffffffffu, c.mlo, s1), b1 = __shfl_sync(0xffffffffu, c.mhi, s1);
    c.mlo = lane < 16 ? a0 : a1;
    c.mhi = lane < 16 ? b0 : b1;
  }
  const uint32_t cnt = (uint32_t)__popc(c.mlo) + (uint32_t)__popc(c.mhi);
  const uint32_t inc = fq_warp_incl_scan(cnt, lane);
  if (lane == 31) wtot[w] = inc;
  return inc - cnt;
}
// after a barrier: exclusive base of the warp and the tile's total from wtot[]
template <class G>
KP_DEV uint32_t fq_census_total(const uint32_t *wtot, uint32_t excl_in_warp, FqCensus<G> &c, int lane, int w) {
  const uint32_t t = lane < G::NW ? wtot[lane] : 0u;
  const uint32_t tinc = fq_warp_incl_scan(t, lane);
  c.rank0 = __shfl_sync(0xffffffffu, tinc - t, w) + excl_in_warp;
  return __shfl_sync(0xffffffffu, tinc, 31);
}

template <class G, bool DS, int KT>
KP_DEV void fq_partition_body(const KpcFqLaunch &p, uint8_t *smem_raw) {
  typedef FqSmemT<G> Smem;
  constexpr int NT = G::NT, TB = G::TB, SEG = G::SEG, VPT = G::VPT;
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const uint32_t s_fill = kp_smem_addr(S.fill), s_bucket = kp_smem_addr(S.bucket);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k = KT ? KT : p.k;
  const int slo = KT == 12 ? 8 : p.lo_bits;           // slice = key bits [slo, slo + sb)
  const int sb = KT == 12 ? 9 : p.slice_bits;
  const uint32_t NS = KT == 12 ? 512u : p.n_slices;
  const uint32_t smask = NS - 1u, lomask = (1u << slo) - 1u;
  const uint32_t cap = (uint32_t)FQ_BUCKET_ENTRIES / NS;  // bucket capacity per slice: a multiple of FQ_CHUNK
  const uint32_t dummy_off = (uint32_t)(offsetof(Smem, dummy) - offsetof(Smem, fill)) + 4u * (uint32_t)lane;
  FqOwner<G> own;
  bool flush_pending = false;
  uint32_t next_tile = 0;                                // thread 0: the tile claimed last

#pragma unroll
  for (int i = 0; i < G::SPT; ++i) {
    const uint32_t s = (uint32_t)tid + (uint32_t)(i * NT);
    own.n[i] = 0; own.g[i] = 0; own.qb16[i] = 0; own.qcap[i] = 0;
    if (s < (uint32_t)FQ_MAXSLICES) S.fill[s] = 0;
    if (s < NS) { own.qb16[i] = (uint32_t)(__ldg(p.qbase + s) / FQ_CHUNK); own.qcap[i] = __ldg(p.qcap + s); }
  }
  for (int i = tid; i < 48; i += NT) S.raw[FQ_HALO + TB + i] = 0;
  if (tid < 32) S.dummy[tid] = 0x40000000u;              // far above any bucket capacity
  for (uint32_t i = tid; i < sizeof(S.bucket) / 16; i += NT)
    reinterpret_cast<uint4 *>(S.bucket)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);  // FQ_PAD everywhere
  for (uint32_t u = 2u + (uint32_t)tid; u <= 64u; u += NT) S.recip[u] = 0xFFFFFFFFu / u + 1u;
  if (tid == 0) {
    S.tileq[0] = atomicAdd(p.counters, 1u);
    S.nbig = 0; S.umax = 0; S.usum = 0;
    kp_mbar_init(&S.mbar, 1);
  }
  const unsigned long long g_in = p.carry_in->s1.count;  // lines before the launch
#ifdef FQ_PHASE_CLOCKS
  if (tid == 32) { for (int i = 0; i < 16; ++i) S.dbg[i] = 0; S.dbg_last = (uint32_t)clock(); }
#endif
  __syncthreads();

  // one bulk copy per tile: [t0 - 16, t0 + len rounded up to 16); a launch without halo starts a line
  // the census reads a tile first and wants it to stay in L2 (evict_last); the bulk copy reads it for the last time
#if FQ_L2HINTS
  const unsigned long long pol_keep = kp_l2_policy_evict_last(), pol_last_use = kp_l2_policy_evict_first();
#else
  const unsigned long long pol_keep = 0, pol_last_use = 0;
#endif
  auto start_tile_load = [&](uint32_t tile) {
    const uint64_t t0 = (uint64_t)tile * TB;
    const uint32_t len = (uint32_t)((p.n - t0) < (uint64_t)TB ? (p.n - t0) : (uint64_t)TB);
    const uint32_t body = (len + 15u) & ~15u;
    if (t0 > 0 || p.halo_ok) {
      kp_bulk_load(S.raw, p.data + t0 - FQ_HALO, body + FQ_HALO, &S.mbar, pol_last_use);
    } else {
      kp_bulk_load(S.raw + FQ_HALO, p.data + t0, body, &S.mbar, pol_last_use);
    }
  };

  // ---- prologue: census of the first tile, claim of the second -----------------------------------------------------------
  FqCensus<G> cur;        // the tile being processed
  uint32_t N = 0;         // its line feeds
  cur.mlo = 0; cur.mhi = 0; cur.rank0 = 0;
  {
    const uint32_t tile0 = S.tileq[0];
    uint32_t excl = 0;
    if (tile0 == 0 && !p.halo_ok && tid < 4) {  // a launch without halo starts a line
      reinterpret_cast<uint32_t *>(S.raw)[tid] = 0x0A0A0A0Au;
      kp_fence_proxy_async();  // a later bulk copy writes these bytes again
    }
    if (tile0 < p.n_tiles) {
      uint4 x[VPT];
      const bool il = fq_census_load<G>(p, tile0, tid, pol_keep, x);
      excl = fq_census_masks<G>(x, il, cur, S.wtot, lane, w);
    }
    if (tid == 0) {
      S.tileq[1] = tile0 < p.n_tiles ? atomicAdd(p.counters, 1u) : 0xffffffffu;
      if (tile0 < p.n_tiles) start_tile_load(tile0);
    }
    __syncthreads();
    if (tile0 < p.n_tiles) {
      N = fq_census_total<G>(S.wtot, excl, cur, lane, w);
      if (tid == 0) fq_lookback_publish(p.tile_state, tile0, N, g_in);
    }
  }

  for (uint32_t it = 0;; ++it) {
    const uint32_t tile = S.tileq[it & 1], tile_next = S.tileq[(it + 1) & 1];
    if (tile >= p.n_tiles) break;
    const uint64_t t0 = (uint64_t)tile * TB;
    const int len = (int)((p.n - t0) < (uint64_t)TB ? (p.n - t0) : (uint64_t)TB);
    const bool have_next = tile_next < p.n_tiles;
    uint32_t *wtot_next = S.wtot2 + 32 * (it & 1);
    bool claimed = false, loaded_next = false;
    FQ_PROBE(0);   // prologue / end of the previous tile

    // the tile two grids ahead is pulled into L2 now: its census is taken one tile time from now
    if (tid == 0) {
      const uint64_t pt = (uint64_t)tile + 2u * gridDim.x;
      if (pt < p.n_tiles) {
        const uint64_t pb = pt * TB;
        const uint32_t pn = (uint32_t)((p.n - pb) < (uint64_t)TB ? (p.n - pb) : (uint64_t)TB) & ~15u;
        if (pn) kp_prefetch_l2(p.data + pb, pn);
      }
    }
    // ---- 1. census of the NEXT tile: the loads (L2 hits) overlap the pending copy-out and the wait for this tile's bytes --
    FqCensus<G> nxt;
    nxt.mlo = 0; nxt.mhi = 0; nxt.rank0 = 0;
    uint32_t nxt_excl = 0, N_next = 0;
    // Warp 0 looks back first (every predecessor published a tile ago, so this is one L2 round trip): its latency runs in
    // parallel with the other warps' copy-out; warp 0's own slices are copied out behind barrier (2) instead.
    unsigned long long g_tile = 0;
    {
      uint4 nx[VPT];
      bool il = false;
      if (have_next) il = fq_census_load<G>(p, tile_next, tid, pol_keep, nx);
      FQ_PROBE(13);
      if (w == 0) g_tile = fq_lookback(p.tile_state, tile, N, g_in, lane);
      else if (flush_pending) fq_flush_copy<G>(S, p, own, tid, cap, slo, sb, lomask);
      FQ_PROBE(14);
      if (have_next) nxt_excl = fq_census_masks<G>(nx, il, nxt, wtot_next, lane, w);
    }
    FQ_PROBE(1);   // census of the next tile, look-back or copy-out
    kp_mbar_wait(&S.mbar, it);
    FQ_PROBE(2);   // wait for the tile's bytes

    // ---- 2. the tile's framing constants ---------------------------------------------------------------------------------
    if (w == 0) {
      // the line the tile starts in: where did it begin?
      int head = -17;
      if (lane < 16 && S.raw[15 - lane] == '\n') head = -(lane + 1);
      const unsigned hm = __ballot_sync(0xffffffffu, head != -17);
      if (hm) head = -(__ffs(hm));
      if (lane == 0) {
        const unsigned long long g = g_tile;
        const uint32_t g4 = (uint32_t)g & 3u;
        const uint32_t jr = (1u - g4) & 3u;                      // first line of the tile (tile relative) on phase 1
        S.G_ = g;
        S.ti.head = head;
        S.ti.jrow0 = jr;
        S.ti.NRt = N >= jr ? (N - jr) / 4u + 1u : 0u;            // sequence lines that touch the tile
        // tile lines below jlim are inside max_lines (incomplete last record, -p cap)
        S.ti.jlim = p.max_lines <= g ? 0u : (p.max_lines - g < 0x7fffffffull ? (uint32_t)(p.max_lines - g) : 0x7fffffffu);
      }
    }
    __syncthreads();  // (2) the framing constants, wtot_next[]
    FQ_PROBE(3);
    if (w == 0 && flush_pending) fq_flush_copy<G>(S, p, own, tid, cap, slo, sb, lomask);  // (barrier (3) comes before any append)
    flush_pending = false;
    if (have_next) {
      N_next = fq_census_total<G>(wtot_next, nxt_excl, nxt, lane, w);
      if (tid == 0) fq_lookback_publish(p.tile_state, tile_next, N_next, g_in);
    }

    const uint4 ti = *reinterpret_cast<const uint4 *>(&S.ti);
    const int head = (int)ti.x;
    const uint32_t jrow0 = ti.y, NRt = ti.z, jlim = ti.w;
    const uint32_t G4 = (1u - jrow0) & 3u;

    // ---- 3./4. lines -> rows -> units, in batches of MAXROWS rows -------------------------------------------------------
    for (uint32_t rb = 0; rb == 0 || rb < NRt; rb += G::MAXROWS) {
      if (rb) __syncthreads();  // the previous batch is done with rowS[] / rowE[] / uinfo[]
      const uint32_t nrows = NRt - rb < (uint32_t)G::MAXROWS ? NRt - rb : (uint32_t)G::MAXROWS;
      if (tid == 0) {
        S.nbig = 0; S.umax = 0; S.usum = 0;
        if (rb == 0) {
          // a line that starts exactly with the tile: tag.[0] <> '@' || tmp.[0] <> '+' (Files.ml:213)
          if (head == -1 && len > 0 && (G4 & 1u) == 0u && jlim > 0u) {
            if (S.raw[FQ_HALO] != (G4 == 0u ? '@' : '+')) atomicMin(p.err_line, S.G_);
          }
          if (jrow0 == 0u && NRt) S.rowS[0] = (uint16_t)(head + 1 + FQ_PBIAS);  // the tile starts inside a sequence line
        }
        // the last line of the tile has no line feed in it: if it is a sequence line it ends with the tile
        if (N >= jrow0 && ((N - jrow0) & 3u) == 0u) {
          const uint32_t r = (N - jrow0) / 4u - rb;
          if (r < (uint32_t)G::MAXROWS) S.rowE[r] = (uint16_t)(len + FQ_PBIAS);
        }
      }
      {
        uint32_t a = cur.mlo, b = cur.mhi, j = cur.rank0;
        while (a | b) {
          uint32_t bitpos;
          if (a) { bitpos = (uint32_t)__ffs(a) - 1u; a &= a - 1u; }
          else { bitpos = 32u + (uint32_t)__ffs(b) - 1u; b &= b - 1u; }
          const int pos = SEG * tid + (int)bitpos;           // line feed at the end of tile line j
          const uint32_t ph1 = (G4 + j + 1u) & 3u;           // phase of the line that starts after it
          if ((ph1 & 1u) == 0u) {
            // header / '+' line: an empty one raises as well (the byte is then '\n')
            if (rb == 0 && j + 1u < jlim && pos + 1 < len) {
              if (S.raw[FQ_HALO + pos + 1] != (ph1 == 0u ? '@' : '+')) atomicMin(p.err_line, S.G_ + j + 1ull);
            }
            if (ph1 == 2u) {                                 // line j is a sequence line: it ends here
              const uint32_t r = (j - jrow0) / 4u - rb;
              if (r < (uint32_t)G::MAXROWS) S.rowE[r] = (uint16_t)(pos + FQ_PBIAS);
            }
          } else if (ph1 == 1u) {                            // line j + 1 is a sequence line: it starts after the line feed
            const uint32_t r = (j + 1u - jrow0) / 4u - rb;
            if (r < (uint32_t)G::MAXROWS) S.rowS[r] = (uint16_t)(pos + 1 + FQ_PBIAS);
          }
          ++j;
        }
      }
      FQ_PROBE(4);   // totals, publish, newline walk
      __syncthreads();  // (3) rowS[] / rowE[] of the batch
      FQ_PROBE(5);

      uint32_t nunits = 0, rinfo = 0;
      if ((uint32_t)tid < nrows && NRt) {
        const int st = (int)S.rowS[tid] - FQ_PBIAS, e = (int)S.rowE[tid] - FQ_PBIAS;
        // first window end: line start + k - 1, inside the tile, rounded down to a word boundary (the windows this
        // adds hold the line feed before the row: they are not valid)
        const int a = st + k - 1 > 0 ? (st + k - 1) & ~3 : 0;
        if (jrow0 + 4u * (rb + (uint32_t)tid) < jlim && e > a) {
          nunits = (uint32_t)(e - a + FQ_W - 1) / FQ_W;
          rinfo = (uint32_t)a | ((uint32_t)e << 16);
        }
      }
      // Units are numbered row by row.  Reads of one length (the usual FASTQ) take the short way: every row gets
      // UPR = (longest row) unit numbers, unit q belongs to row q / UPR, and nothing is scanned or stored.
      if (NRt && (uint32_t)tid < ((nrows + 31u) & ~31u)) {
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, nunits), wsum = __reduce_add_sync(0xffffffffu, nunits);
        if (lane == 0 && wsum) { atomicMax(&S.umax, wmax); atomicAdd(&S.usum, wsum); }
      }
      __syncthreads();  // (4)
      FQ_PROBE(6);   // units per row + barrier
      const uint32_t NR = NRt ? nrows : 0u;
      const uint32_t UPR = S.umax, usum = S.usum;
      const bool uniform = NR * UPR <= usum + (usum >> 2) + 64u;
      // q / UPR = umulhi(q, recip) for q < 2^16, UPR > 1
      const uint32_t recip = UPR <= 1u ? 0u : (UPR <= 64u ? S.recip[UPR] : 0xFFFFFFFFu / UPR + 1u);
      uint32_t U = NR * UPR;
      if (!uniform) {
        const uint32_t ub = fq_block_excl_scan<G::NW>(nunits, S.wtot, U, lane, w);
        if (nunits) {
          if (nunits <= (uint32_t)FQ_BIGROW) {
            for (uint32_t u = 0; u < nunits; ++u) S.uinfo[ub + u] = rinfo + u * FQ_W;
          } else {
            const uint32_t i = atomicAdd(&S.nbig, 1u);
            S.big[i].ub = ub; S.big[i].info = rinfo; S.big[i].n = nunits;
          }
        }
        __syncthreads();
        const uint32_t nb = S.nbig;
        if (nb) {  // long lines: every thread fills its share
          for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t bub = S.big[b].ub, bi = S.big[b].info, bn = S.big[b].n;
            for (uint32_t u = tid; u < bn; u += NT) S.uinfo[bub + u] = bi + u * FQ_W;
          }
          __syncthreads();
        }
      }
      const bool last_batch = rb + (uint32_t)G::MAXROWS >= NRt;

      // the state the next launch starts from (only the tile that ends the launch)
      if (tid == 0 && tile == p.n_tiles - 1 && last_batch) {
        KpcStreamCarry co;
        co.s1.count = S.G_ + N;
        co.s1.last_hdr = 0;
        int plast = head;  // position of the last line feed of the tile (or of the halo)
        if (N) {
          plast = len - 1;
          while (plast >= 0 && S.raw[FQ_HALO + plast] != '\n') --plast;
        }
        co.s1.last_nl = N ? p.abs_base + t0 + (uint64_t)plast + 1u : p.carry_in->s1.last_nl;
        co.kc.syms = 0; co.kc.n = 0; co.kc.closed = 1;
        if (((G4 + N) & 3u) == 1u && N < jlim) {
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

      // ---- 5. rounds of NT units ------------------------------------------------------------------------------------------
      for (uint32_t q0 = 0; q0 < U; q0 += NT) {
        const bool last_round = last_batch && q0 + NT >= U;
        FQ_PROBE(7);   // unit numbering, carry-out / the reserve of the previous round
        // the tile after the next one is claimed late (tiles publish in claim order: an early claim would make every
        // later tile wait for this CTA), but early enough for the atomic to return before the tile ends
        if (last_round && !claimed) {
          if (tid == 0) next_tile = have_next ? atomicAdd(p.counters, 1u) : 0xffffffffu;
          claimed = true;
        }
        uint32_t ok = 0;
        uint32_t hi24 = 0, lo32 = 0, rlo = 0, rhi = 0;
        const uint32_t q = q0 + tid;
        uint32_t info = 0xFFFFu;  // p0 = 0xFFFF, e = 0: no unit
        if (q < U) {
          if (uniform) {
            const uint32_t r = UPR == 1u ? q : __umulhi(q, recip), u = q - r * UPR;
            const int st = (int)S.rowS[r] - FQ_PBIAS, e = (int)S.rowE[r] - FQ_PBIAS;
            const int a = (st + k - 1 > 0 ? (st + k - 1) & ~3 : 0) + (int)(u * FQ_W);
            if (jrow0 + 4u * (rb + r) < jlim && a < e) info = (uint32_t)a | ((uint32_t)e << 16);
          } else {
            info = S.uinfo[q];
          }
        }
        if ((info & 0xFFFFu) < (info >> 16)) {
          const int p0 = (int)(info & 0xFFFFu), e = (int)(info >> 16);
          const int nvalid = e - p0 < FQ_W ? e - p0 : FQ_W;
          // bytes [p0 - 12, p0 + 16): 7 words (units start on word boundaries)
          const uint32_t *rw = reinterpret_cast<const uint32_t *>(S.raw) + ((FQ_HALO + p0 - FQ_CTX) >> 2);
          uint32_t xw[7];
#pragma unroll
          for (int m = 0; m < 7; ++m) xw[m] = rw[m];
          // the 2-bit fields and the 0x80 "not a base" flags of four bytes are gathered with one dot product each
          uint32_t xh = 0, xl = 0, iA = 0, iB = 0, iC = 0, iD = 0;
#pragma unroll
          for (int m = 0; m < 7; ++m) {
            const uint32_t wd = xw[m];
            uint32_t nz;
            const uint32_t x = fq_classify4(wd, nz);
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
        FQ_PROBE(8);   // classification
        if (flush_pending) {  // the copy-out reserved after the previous round (its atomic has had time to return)
          fq_flush_copy<G>(S, p, own, tid, cap, slo, sb, lomask);
          flush_pending = false;
        }
        FQ_PROBE(9);   // copy-out
        __syncthreads();      // (B) the buckets may be appended to again; raw[] has been read for the last time in this round
        if (last_round) {     // nothing reads the tile bytes any more: the next tile may land on them
          if (tid == 0) {
            S.tileq[it & 1] = next_tile;
            if (have_next) start_tile_load(tile_next);
          }
          loaded_next = true;
        }
        // ---- canonical keys, appended to the buckets as they are made ---------------------------------------------------
        // window end jw: forward k-mer = bits of hi24:lo32, reverse complement = bits of rhi:rlo, both moved to the top
        // of a 32-bit word (KMers.ml:364-368); min (KMers.ml:388) ignores the bits below because they only matter
        // when the k-mers are equal
        FQ_PROBE(10);  // barrier (B)
        if (ok) {
          uint32_t pending = ok;
          if (KT == 12) {
#pragma unroll
            for (int jw = 0; jw < FQ_W; jw += FQ_APPEND_GROUP) fq_append4_k12<DS>(S, hi24, lo32, rhi, rlo, jw, ok, pending, dummy_off);
          }
#pragma unroll
          for (int jw = 0; jw < FQ_W; ++jw) {
            if (KT == 12) break;
            uint32_t kk = fq_top_word(hi24, lo32, 34 + 2 * jw - 2 * k);
            if (DS) kk = kp_umin(kk, fq_top_word(rhi, rlo, 38 - 2 * jw));
            const uint32_t bit = 1u << (FQ_W - 1 - jw);
            if (pending & bit) {
              const uint32_t key = kk >> (32 - 2 * k);
              const uint32_t sl = fq_slice_of(key, slo, smask);
              const uint32_t pos = kp_atoms_add(s_fill + 4u * sl, 1u);
              if (pos < cap) {
                kp_sts_u16(s_bucket + 2u * (sl * (cap + FQ_BPAD) + pos), fq_bin_of(key, slo, sb, lomask));
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
            if (DS) kk = kp_umin(kk, fq_top_word(rhi, rlo, 38 - 2 * jw));
            atomicAdd(p.table + (kk >> (32 - 2 * k)), 1u);
          }
        }
        FQ_PROBE(11);  // appends
        __syncthreads();  // (A) the appends of the round are complete
        FQ_PROBE(12);  // barrier (A)
        fq_flush_reserve<G>(S, p, own, tid, NS, cap, false);
        flush_pending = true;
      }
    }
    // a tile without a single unit in its last batch: claim and load here
    if (!loaded_next) {
      __syncthreads();  // every thread is done with raw[]
      if (tid == 0) {
        if (!claimed) next_tile = have_next ? atomicAdd(p.counters, 1u) : 0xffffffffu;
        S.tileq[it & 1] = next_tile;
        if (have_next) start_tile_load(tile_next);
      }
    }
    cur = nxt;
    N = N_next;
    __syncthreads();  // tileq[], and rowS[] / rowE[] / the scan scratch are free for the next tile
  }
#ifdef FQ_PHASE_CLOCKS
  if (tid == 32) for (int i = 0; i < 16; ++i) atomicAdd(&g_fq_phase[i], S.dbg[i]);
#endif
  // the CTA leaves: everything still in the buckets goes out, the last chunk of every slice padded
  if (flush_pending) fq_flush_copy<G>(S, p, own, tid, cap, slo, sb, lomask);
  __syncthreads();
  fq_flush_reserve<G>(S, p, own, tid, NS, cap, true);
  fq_flush_copy<G>(S, p, own, tid, cap, slo, sb, lomask);
}

// ---- fq_count: one CTA per slice at a time ------------------------------------------------------------------------------
// The slice's 2^log_bins u32 bins live in shared memory, queue entries are counted with shared-memory atomics, non-zero
// bins are added to the global table (IntHashFrequencies.add, KMers.ml:107-111).
KP_DEV void fq_count2(uint32_t *tbl, uint32_t x) {
  const uint32_t a = x & 0xFFFFu, b = x >> 16;
  if (a != FQ_PAD) atomicAdd(&tbl[a], 1u);
  if (b != FQ_PAD) atomicAdd(&tbl[b], 1u);
}
template <int CNT_NT>
KP_DEV void fq_count_body(const KpcFqLaunch &p, uint8_t *smem_raw) {
  uint32_t *tbl = reinterpret_cast<uint32_t *>(smem_raw);
  const int tid = threadIdx.x;
  const int lb = p.log_bins, slo = p.lo_bits, sb = p.slice_bits;
  const uint32_t lomask = (1u << slo) - 1u;
  const uint32_t nbins = 1u << lb;
  uint32_t *s_slice = tbl + nbins;  // the item claimed by thread 0
  for (uint32_t i = tid; i < nbins; i += CNT_NT) tbl[i] = 0;
  // work items: whole slices while they fill complete waves of the grid; the slices of the last, partial wave are cut
  // into `parts` pieces so that it keeps every SM busy as well (512 slices on 148 SMs: 444 whole + 68 x 2 halves)
  const uint32_t GR = gridDim.x, rem = p.n_slices % GR;
  const uint32_t parts = rem ? (GR / rem < 4u ? GR / rem : 4u) : 1u;
  const uint32_t whole = p.n_slices - rem, n_items = whole + rem * parts;
  for (;;) {
    __syncthreads();
    if (tid == 0) *s_slice = atomicAdd(p.counters + 1, 1u);
    __syncthreads();
    const uint32_t item = *s_slice;
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
    for (; v + 3u * CNT_NT < nvec; v += 4u * CNT_NT) {
      uint4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = kp_ldg_stream(reinterpret_cast<const uint8_t *>(sv + v + u * CNT_NT));
#pragma unroll
      for (int u = 0; u < 4; ++u) { fq_count2(tbl, x[u].x); fq_count2(tbl, x[u].y); fq_count2(tbl, x[u].z); fq_count2(tbl, x[u].w); }
    }
    for (; v < nvec; v += CNT_NT) {
      const uint4 x = kp_ldg_stream(reinterpret_cast<const uint8_t *>(sv + v));
      fq_count2(tbl, x.x); fq_count2(tbl, x.y); fq_count2(tbl, x.z); fq_count2(tbl, x.w);
    }
    __syncthreads();
    for (uint32_t i = tid; i < nbins; i += CNT_NT) {
      const uint32_t cv = tbl[i];
      if (cv) { atomicAdd(p.table + fq_key_of(b, i, slo, sb, lomask), cv); tbl[i] = 0; }
    }
  }
}
