// kpc_fastq.cu -- fast FASTQ -> dense 4^k table pipeline for sm_100a (see kpc_fastq.h for the outline).
//
// Reference semantics restated here (paths relative to the KPop tree):
//   Files.FASTQ.iter_se        BiOCamLib/lib/Files.ml:201-221   4-line records, '@' / '+' checks, only line 2 is sequence
//   Sequences.Lint.dnaize      BiOCamLib/lib/Sequences.ml:41-67  [ACGTacgt] are bases, every other byte breaks k-mers
//   DNAHash*.iteri / iterc     BiOCamLib/lib/KMers.ml:319-349, 357-389   first base most significant, key = min f rc
//   IntHashFrequencies.add     BiOCamLib/lib/KMers.ml:107-111   count[key] += 1
//
// Geometry: a CTA of FQ_NT threads owns one tile of FQ_TB bytes at a time (claimed in stream order).
//   1. 128-bit loads -> shared memory; the newline census is taken from the registers on the way
//   2. block scan of the newline counts; the tile's line index base comes from a single-word decoupled look-back
//   3. every line of the tile gets a slot (start, end, line index mod 4); sequence lines are cut into units of
//      FQ_W window-end positions; a scan over the rows gives every unit a number
//   4. one thread per unit: 12 + 16 bytes -> 2-bit codes + validity bits with SIMD-in-register arithmetic,
//      forward and reverse-complement words, 16 canonical keys in registers
//   5. software write-combining: every key is appended (one shared-memory atomic) to the bucket of its slice in
//      shared memory; full 32-byte chunks of a bucket are reserved in the slice's queue in HBM (one global atomic per
//      slice and round, issued early so that its latency hides behind the next round's k-mer arithmetic) and copied
//      out with 128-bit stores.  The slice is taken from the MIDDLE bits of the key: min(f, rc) skews the top and
//      the bottom bases of a canonical k-mer but leaves the central ones uniform, so the buckets fill evenly.
#include <cuda_runtime.h>

#include <string>

#include "../../include/kpopcount.h"
#include "kpc_fastq.h"

#define FQ_CUDA_CHECK(x)                                                                                      \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      throw KpcError(KPC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +        \
                                     std::to_string(__LINE__) + ")");                                         \
  } while (0)

namespace {

constexpr int FQ_NT = 512;                        // threads per CTA
constexpr int FQ_NW = FQ_NT / 32;
constexpr int FQ_W = 16;                          // window-end positions per unit (one thread)
constexpr int FQ_CTX = 12;                        // context bytes loaded before a unit (>= k - 1)
constexpr int FQ_TB = 32768;                      // tile bytes
constexpr int FQ_PIECES = FQ_TB / (16 * FQ_NT);   // 16-byte vectors per thread
constexpr int FQ_HALO = 16;
constexpr int FQ_MAXROWS = FQ_NT;                 // sequence lines per batch
constexpr int FQ_MAXSLOTS = 4 * FQ_NT;            // lines per batch
constexpr int FQ_MAXSLICES = 512;
constexpr int FQ_BIAS = 17;                       // positions are stored + FQ_BIAS (they start at -17)
constexpr int FQ_BUCKET_ENTRIES = 24576;          // shared-memory bucket space (u16 entries) shared by all slices
constexpr int FQ_CHUNK = 16;                      // entries per copy-out chunk (32 bytes: one L2 sector)
constexpr int FQ_RETRIES = 2;                     // rounds of bucket overflow before keys are counted in place
constexpr uint16_t FQ_PAD = 0xFFFFu;              // queue entry that pads the last chunk of a CTA (skipped by fq_count)
static_assert(FQ_PIECES * FQ_NW % 32 == 0, "census scan layout");
static_assert(FQ_MAXSLICES <= FQ_NT, "one thread per slice in the scan");

struct FqSmem {
  alignas(16) uint8_t raw[FQ_HALO + FQ_TB + 48];  // raw[16 + i] = tile byte i
  alignas(16) uint16_t bucket[FQ_BUCKET_ENTRIES + 2 * FQ_CHUNK];  // slice s owns [s * cap, (s + 1) * cap)
  unsigned long long G;                           // number of '\n' in the stream before the tile
  uint32_t fill[FQ_MAXSLICES];                    // entries in the bucket (may run past cap while appending)
  uint32_t fl_g[FQ_MAXSLICES];                    // queue position reserved for the chunks being copied out
  uint32_t fl_n[FQ_MAXSLICES];                    // entries being copied out (multiple of FQ_CHUNK)
  uint32_t ubase[FQ_MAXROWS + 1];
  uint32_t wtot_a[FQ_PIECES * FQ_NW];
  uint32_t wtot_b[32];
  uint32_t tileq[2];
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
// 0x80 in every byte of w that equals '\n'
__device__ __forceinline__ uint32_t nl_mask(uint32_t w) {
  uint32_t y = w ^ 0x0A0A0A0Au;
  return ~((((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) | 0x7F7F7F7Fu);
}
// Four input bytes (first byte in the low bits) -> 2-bit codes packed first-base-most-significant in 8 bits, and
// 4 validity bits in the same order.  A=0 C=1 G=2 T=3, case-insensitive (KMers.ml:272-277 after Sequences.ml:52-58).
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &codes8, uint32_t &valid4) {
  uint32_t x = (w >> 1) & 0x03030303u;                 // A->0 C->1 T->2 G->3
  uint32_t c = x ^ ((x >> 1) & 0x01010101u);           // A->0 C->1 G->2 T->3
  codes8 = (c * 0x40100401u) >> 24;
  // the letter each byte would have to be for its code, through a 4-entry byte table in a register
  uint32_t t = x | (x >> 4);
  uint32_t sel = __byte_perm(t, 0u, 0x4420u);
  uint32_t expect = __byte_perm(0x47544341u, 0u, sel);  // table: 0->'A' 1->'C' 2->'T' 3->'G'
  uint32_t y = (w & 0xDFDFDFDFu) ^ expect;              // zero byte <=> valid base
  uint32_t nz = (((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
  valid4 = (((nz >> 7) * 0x08040201u) >> 24) ^ 0xFu;
}


// slice / bin of a key and back: slice = key bits [lo, lo + sb), bin = the other bits packed together
__device__ __forceinline__ uint32_t fq_slice_of(uint32_t key, int lo, uint32_t smask) { return (key >> lo) & smask; }
__device__ __forceinline__ uint32_t fq_bin_of(uint32_t key, int lo, int sb, uint32_t lomask) {
  return ((key >> (lo + sb)) << lo) | (key & lomask);
}
__device__ __forceinline__ uint32_t fq_key_of(uint32_t slice, uint32_t bin, int lo, int sb, uint32_t lomask) {
  return ((bin >> lo) << (lo + sb)) | (slice << lo) | (bin & lomask);
}

// number of '\n' before the tile: decoupled look-back over one word per tile (2 status bits + 62 value bits)
constexpr unsigned long long FQ_ST_AGG = 1ull << 62, FQ_ST_INC = 2ull << 62, FQ_VAL = (1ull << 62) - 1ull;
__device__ __forceinline__ unsigned long long lookback(unsigned long long *state, uint32_t tile, uint32_t total,
                                                       unsigned long long g_in, int lane) {
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(state, FQ_ST_INC | (g_in + total));
    return g_in;
  }
  if (lane == 0) st_relaxed_u64(state + tile, FQ_ST_AGG | (unsigned long long)total);
  unsigned long long acc = 0;
  long long j0 = (long long)tile - 1;
  for (;;) {
    const long long j = j0 - lane;
    unsigned long long v;
    if (j >= 0) v = ld_relaxed_u64(state + j);
    else if (j == -1) v = FQ_ST_INC | g_in;
    else v = FQ_ST_AGG;  // never used: lies behind the inclusive entry at j == -1
    const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
    const unsigned inv_mask = __ballot_sync(0xffffffffu, (v >> 62) == 0ull);
    const int first = inc_mask ? (__ffs(inc_mask) - 1) : 32;
    const unsigned need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
    if (inv_mask & need) continue;  // a predecessor has not published yet
    unsigned long long val = ((need >> lane) & 1u) ? (v & FQ_VAL) : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    acc += val;
    if (first < 32) break;
    j0 -= 32;
  }
  if (lane == 0) st_relaxed_u64(state + tile, FQ_ST_INC | (acc + total));
  return acc;
}


// ---- bucket copy-out (step 5) ------------------------------------------------------------------------------------
// reserve: thread s owns slice s.  Called after a barrier that follows the appends.  Whole chunks of the bucket are
// reserved in the slice's queue; the result of the atomic is not needed before fq_flush_copy.
__device__ __forceinline__ void fq_flush_reserve(FqSmem &S, const KpcFqLaunch &p, int tid, uint32_t NS, uint32_t cap,
                                                 bool final, uint32_t &my_n, uint32_t &my_g) {
  my_n = 0; my_g = 0;
  if ((uint32_t)tid < NS) {
    uint32_t f = S.fill[tid];
    if (f > cap) f = cap;
    uint32_t n = f & ~(uint32_t)(FQ_CHUNK - 1);
    if (final && n < f) {  // the CTA is leaving: pad the last chunk
      for (uint32_t i = f; i < n + FQ_CHUNK; ++i) S.bucket[tid * cap + i] = FQ_PAD;
      n += FQ_CHUNK;
      f = n;
    }
    S.fill[tid] = f - n;
    my_n = n;
    if (n) my_g = atomicAdd(p.qcursor + tid, n);
  }
}
// copy: two lanes per slice move the reserved chunks with 128-bit accesses, then the bucket's remainder (< one
// chunk) moves to the front.  Ends with a barrier: the buckets may be appended to again.
__device__ __forceinline__ void fq_flush_copy(FqSmem &S, const KpcFqLaunch &p, int tid, uint32_t NS, uint32_t cap,
                                              uint32_t my_n, uint32_t my_g, int lo, int sb, uint32_t lomask) {
  if ((uint32_t)tid < NS) { S.fl_n[tid] = my_n; S.fl_g[tid] = my_g; }
  __syncthreads();
  const int half = 8 * (tid & 1);
  for (uint32_t s = tid >> 1; s < NS; s += FQ_NT / 2) {
    const uint32_t n = S.fl_n[s];
    if (!n) continue;
    const uint32_t g = S.fl_g[s];
    const uint32_t qc = __ldg(p.qcap + s);
    uint16_t *dst = p.queue + __ldg(p.qbase + s) + g + half;
    uint16_t *src = S.bucket + s * cap + half;
    for (uint32_t c = 0; c < n; c += FQ_CHUNK) {
      const uint4 v = *reinterpret_cast<const uint4 *>(src + c);
      if (g + c + FQ_CHUNK <= qc) {
        *reinterpret_cast<uint4 *>(dst + c) = v;
      } else {  // queue full: count in place
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t e = (ww[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
          if (e != FQ_PAD) atomicAdd(p.table + fq_key_of(s, e, lo, sb, lomask), 1u);
        }
      }
    }
    if (n < cap) {
      const uint4 v = *reinterpret_cast<const uint4 *>(src + n);
      *reinterpret_cast<uint4 *>(src) = v;
    }
  }
  __syncthreads();
}

template <bool DS, int KT>
__global__ void __launch_bounds__(FQ_NT, 2) fq_partition_kernel(const KpcFqLaunch p) {
  extern __shared__ __align__(16) uint8_t fq_smem_raw[];
  FqSmem &S = *reinterpret_cast<FqSmem *>(fq_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k = KT ? KT : p.k;
  const uint32_t kmask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
  const int slo = KT == 12 ? 7 : p.lo_bits;           // slice = key bits [slo, slo + sb)
  const int sb = KT == 12 ? 9 : p.slice_bits;
  const uint32_t NS = KT == 12 ? 512u : p.n_slices;
  const uint32_t smask = NS - 1u, lomask = (1u << slo) - 1u;
  const uint32_t cap = (uint32_t)FQ_BUCKET_ENTRIES / NS;  // bucket capacity per slice: a multiple of FQ_CHUNK
  uint32_t my_n = 0, my_g = 0;                           // copy-out in flight for slice tid
  bool flush_pending = false;
  uint32_t next_tile = 0;                                // thread 0: the tile claimed for the next iteration

  if (tid < FQ_MAXSLICES) S.fill[tid] = 0;
  if (tid < 48) S.raw[FQ_HALO + FQ_TB + tid] = 0;
  if (tid == 0) S.tileq[0] = atomicAdd(p.counters, 1u);
  __syncthreads();

  for (uint32_t it = 0;; ++it) {
    const uint32_t tile = S.tileq[it & 1];
    if (tile >= p.n_tiles) break;
    const uint64_t t0 = (uint64_t)tile * FQ_TB;
    const int len = (int)((p.n - t0) < (uint64_t)FQ_TB ? (p.n - t0) : (uint64_t)FQ_TB);

    // ---- 1. load + newline census from the registers ------------------------------------------------------
    uint32_t nl[FQ_PIECES][4];
    uint32_t c[FQ_PIECES], inc[FQ_PIECES];
#pragma unroll
    for (int q = 0; q < FQ_PIECES; ++q) {
      const int off = 16 * (q * FQ_NT + tid);
      uint4 x = make_uint4(0u, 0u, 0u, 0u);
      if (off < len) x = ldg_stream(p.data + t0 + off);
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
      nl[q][0] = nl_mask(x.x); nl[q][1] = nl_mask(x.y); nl[q][2] = nl_mask(x.z); nl[q][3] = nl_mask(x.w);
      c[q] = __popc(nl[q][0]) + __popc(nl[q][1]) + __popc(nl[q][2]) + __popc(nl[q][3]);
    }
    if (tid == FQ_NT - 1) {
      uint4 h = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);  // a launch without halo starts a line
      if (t0 > 0 || p.halo_ok) h = ldg_stream(p.data + t0 - 16);
      *reinterpret_cast<uint4 *>(S.raw) = h;
    }
#pragma unroll
    for (int q = 0; q < FQ_PIECES; ++q) inc[q] = warp_incl_scan(c[q], lane);
    if (lane == 31) {
#pragma unroll
      for (int q = 0; q < FQ_PIECES; ++q) S.wtot_a[q * FQ_NW + w] = inc[q];
    }
    __syncthreads();  // (1) raw[] and wtot_a[] are complete
    bool claimed = false;

    // newline index of the first newline of every piece (pieces are ordered (q, thread))
    uint32_t base[FQ_PIECES];
    uint32_t N;
    {
      constexpr int VPL = FQ_PIECES * FQ_NW / 32;
      uint32_t a[VPL], s = 0;
#pragma unroll
      for (int i = 0; i < VPL; ++i) { a[i] = S.wtot_a[VPL * lane + i]; s += a[i]; }
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
        base[q] = e + inc[q] - c[q];
      }
    }
    if (w == 0) {
      // the line the tile starts in: where did it begin?
      int head = -17;
      if (lane < 16 && S.raw[15 - lane] == '\n') head = -(lane + 1);
      const unsigned m = __ballot_sync(0xffffffffu, head != -17);
      if (m) head = -(__ffs(m));
      const unsigned long long g = lookback(p.tile_state, tile, N, p.carry_in->s1.count, lane);
      if (lane == 0) { S.G = g; S.head = head; }
    }

    // nlpos[j] = position of newline (lo - 1 + j), j = 0 .. FQ_MAXSLOTS; first batch straight from the registers
#pragma unroll
    for (int q = 0; q < FQ_PIECES; ++q) {
      if (c[q]) {
        uint32_t idx = base[q];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          uint32_t v = nl[q][m];
          while (v) {
            const int bit = __ffs(v) - 1;
            v &= v - 1u;
            if (idx < (uint32_t)FQ_MAXSLOTS) S.nlpos[idx + 1u] = (uint16_t)(16 * (q * FQ_NT + tid) + 4 * m + (bit >> 3) + FQ_BIAS);
            ++idx;
          }
        }
      }
    }

    // ---- 2./3. lines -> slots -> rows -> units, in batches of FQ_MAXSLOTS lines ------------------------------
    for (uint32_t lo = 0; lo < N + 1u; lo += FQ_MAXSLOTS) {
      if (lo) {
        __syncthreads();  // the previous batch is done with nlpos[] / ubase[]
        // tiles with more than FQ_MAXSLOTS lines (rare): the newline masks are recomputed from shared memory
#pragma unroll
        for (int q = 0; q < FQ_PIECES; ++q) {
          if (c[q]) {
            const uint4 x = *reinterpret_cast<const uint4 *>(S.raw + FQ_HALO + 16 * (q * FQ_NT + tid));
            const uint32_t xm[4] = {nl_mask(x.x), nl_mask(x.y), nl_mask(x.z), nl_mask(x.w)};
            uint32_t idx = base[q];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              uint32_t v = xm[m];
              while (v) {
                const int bit = __ffs(v) - 1;
                v &= v - 1u;
                const uint32_t j = idx + 1u - lo;  // wraps for idx + 1 < lo: rejected by the range test
                if (j <= (uint32_t)FQ_MAXSLOTS) S.nlpos[j] = (uint16_t)(16 * (q * FQ_NT + tid) + 4 * m + (bit >> 3) + FQ_BIAS);
                ++idx;
              }
            }
          }
        }
      }
      __syncthreads();  // (2) nlpos[] of the batch, G and head are visible
      const unsigned long long G = S.G;
      if (tid == 0) {
        if (lo == 0) S.nlpos[0] = (uint16_t)(S.head + FQ_BIAS);
        if (N + 1u - lo <= (uint32_t)FQ_MAXSLOTS) S.nlpos[N + 1u - lo] = (uint16_t)(len + FQ_BIAS);  // the last line ends with the tile
      }
      const uint32_t hi = (N + 1u < lo + FQ_MAXSLOTS) ? N + 1u : lo + FQ_MAXSLOTS;  // slots [lo, hi)
      const uint32_t jrow0 = (uint32_t)((1ull - (G + lo)) & 3ull);                  // first slot (batch relative) on phase 1
      const uint32_t NR = (hi - lo > jrow0) ? (hi - lo - jrow0 + 3u) / 4u : 0u;
      __syncthreads();  // (2b) the two entries written by thread 0
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
      uint32_t nunits = 0;
      if ((uint32_t)tid < NR) {
        const uint32_t j = jrow0 + 4u * tid;
        const int pm = (int)S.nlpos[j] - FQ_BIAS, e = (int)S.nlpos[j + 1] - FQ_BIAS;
        const int a = pm + k > 0 ? pm + k : 0;  // first window end: line start + k - 1, inside the tile
        if (G + lo + j < p.max_lines && e > a) nunits = (uint32_t)(e - a + FQ_W - 1) / FQ_W;
      }
      uint32_t U;
      const uint32_t ub = block_excl_scan(nunits, S.wtot_b, U, lane, w);  // (3)
      if (tid < FQ_MAXROWS) S.ubase[tid] = ub;
      if (tid == 0) S.ubase[FQ_MAXROWS] = U;
      __syncthreads();  // (4)

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
      for (uint32_t q0 = 0; q0 < U; q0 += FQ_NT) {
        uint32_t key[FQ_W];
        uint32_t ok = 0;
        const uint32_t q = q0 + tid;
        if (q < U) {
          uint32_t r = 0;
#pragma unroll
          for (uint32_t step = FQ_MAXROWS / 2; step; step >>= 1) {
            const uint32_t cand = r + step;
            if (cand < NR && S.ubase[cand] <= q) r = cand;
          }
          const uint32_t j = jrow0 + 4u * r;
          const int pm = (int)S.nlpos[j] - FQ_BIAS, e = (int)S.nlpos[j + 1] - FQ_BIAS;
          const int a = pm + k > 0 ? pm + k : 0;
          const int p0 = a + (int)(q - S.ubase[r]) * FQ_W;
          const int nvalid = e - p0 < FQ_W ? e - p0 : FQ_W;
          // bytes [p0 - 12, p0 + 16): 8 aligned words, funnel-shifted to 7
          const int A = FQ_HALO + p0 - FQ_CTX;
          const uint32_t *rw = reinterpret_cast<const uint32_t *>(S.raw) + (A >> 2);
          const uint32_t sh = (uint32_t)(A & 3) * 8u;
          uint32_t xw[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) xw[m] = rw[m];
          uint32_t hi24 = 0, lo32 = 0, valid28 = 0;
#pragma unroll
          for (int m = 0; m < 7; ++m) {
            uint32_t c8, v4;
            classify4(__funnelshift_r(xw[m], xw[m + 1], sh), c8, v4);
            if (m < 3) hi24 = (hi24 << 8) | c8; else lo32 = (lo32 << 8) | c8;
            valid28 = (valid28 << 4) | v4;
          }
          // bit b of okm <=> the k bases whose validity bits are b .. b + k - 1 are all valid
          uint32_t okm = valid28;
          {
            int have = 1;  // bit b of okm = AND of validity bits b .. b + have - 1
            while (2 * have <= k) { okm &= okm >> have; have *= 2; }
            if (have < k) okm &= okm >> (k - have);
          }
          ok = okm & 0xFFFFu;
          if (nvalid < FQ_W) ok &= ~((1u << (FQ_W - nvalid)) - 1u);
          // reverse complement of the 28 bases: base i (complemented) at bits 2i+1 : 2i
          uint32_t rlo = 0, rhi = 0;
          if (DS) {
            const uint32_t nh = __brev(lo32), nlw = __brev(hi24);  // 64-bit reversal of hi24:lo32
            const uint32_t xl = __funnelshift_r(nlw, nh, 8), xh = nh >> 8;
            rlo = ~(((xl >> 1) & 0x55555555u) | ((xl & 0x55555555u) << 1));
            rhi = ~(((xh >> 1) & 0x55555555u) | ((xh & 0x55555555u) << 1)) & 0x00FFFFFFu;
          }
#pragma unroll
          for (int jw = 0; jw < FQ_W; ++jw) {
            uint32_t f = __funnelshift_r(lo32, hi24, 2 * (FQ_W - 1 - jw)) & kmask;   // KMers.ml:366-367
            if (DS) {
              const int s = 2 * (FQ_CTX + 1 + jw - k);                                // KMers.ml:368
              uint32_t rc = (s < 32 ? __funnelshift_r(rlo, rhi, s) : (rhi >> (s - 32))) & kmask;
              f = f < rc ? f : rc;                                                    // KMers.ml:388
            }
            key[jw] = f;
          }
        }
        // ---- 5. append to the buckets; whole chunks go out to the queues ----------------------------------------
        // the next tile is claimed as late as possible (tiles are published in claim order: an early claim makes every
        // later tile wait for this CTA), but early enough for the atomic to return before the tile ends
        if (!claimed && hi == N + 1u && q0 + FQ_NT >= U) {
          if (tid == 0) next_tile = atomicAdd(p.counters, 1u);
          claimed = true;
        }
        uint32_t pending = ok;
        for (int tries = 0;; ++tries) {
          if (flush_pending) {  // the copy-out reserved after the previous round (its atomic has had time to return)
            fq_flush_copy(S, p, tid, NS, cap, my_n, my_g, slo, sb, lomask);
            flush_pending = false;
          }
#pragma unroll
          for (int jw = 0; jw < FQ_W; ++jw) {
            const uint32_t bit = 1u << (FQ_W - 1 - jw);
            if (pending & bit) {
              const uint32_t sl = fq_slice_of(key[jw], slo, smask);
              const uint32_t pos = atomicAdd(&S.fill[sl], 1u);
              if (pos < cap) {
                S.bucket[sl * cap + pos] = (uint16_t)fq_bin_of(key[jw], slo, sb, lomask);
                pending &= ~bit;
              } else if (tries >= FQ_RETRIES) {  // a slice that keeps overflowing (skewed input): count in place
                atomicAdd(p.table + key[jw], 1u);
                pending &= ~bit;
              }
            }
          }
          const int any = __syncthreads_or(pending != 0u);
          fq_flush_reserve(S, p, tid, NS, cap, false, my_n, my_g);
          flush_pending = true;
          if (!any) break;
        }
      }
    }
    if (!claimed && tid == 0) next_tile = atomicAdd(p.counters, 1u);
    if (tid == 0) S.tileq[(it + 1) & 1] = next_tile;
    __syncthreads();  // the next tile overwrites raw[], nlpos[] and the scan scratch
  }
  // the CTA leaves: everything still in the buckets goes out, the last chunk of every slice padded
  if (flush_pending) fq_flush_copy(S, p, tid, NS, cap, my_n, my_g, slo, sb, lomask);
  fq_flush_reserve(S, p, tid, NS, cap, true, my_n, my_g);
  fq_flush_copy(S, p, tid, NS, cap, my_n, my_g, slo, sb, lomask);
}

// one CTA per slice at a time: shared-memory histogram of the slice's queue, then RED of the non-zero bins
constexpr int FQ_CNT_NT = 1024;
__device__ __forceinline__ void fq_count2(uint32_t *tbl, uint32_t x) {
  const uint32_t a = x & 0xFFFFu, b = x >> 16;
  if (a != FQ_PAD) atomicAdd(&tbl[a], 1u);
  if (b != FQ_PAD) atomicAdd(&tbl[b], 1u);
}
__global__ void __launch_bounds__(FQ_CNT_NT, 1) fq_count_kernel(const KpcFqLaunch p) {
  extern __shared__ __align__(16) uint8_t fq_smem_raw[];
  uint32_t *tbl = reinterpret_cast<uint32_t *>(fq_smem_raw);
  __shared__ uint32_t s_slice;
  const int tid = threadIdx.x;
  const int lb = p.log_bins, slo = p.lo_bits, sb = p.slice_bits;
  const uint32_t lomask = (1u << slo) - 1u;
  const uint32_t nbins = 1u << lb;
  for (uint32_t i = tid; i < nbins; i += FQ_CNT_NT) tbl[i] = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_slice = atomicAdd(p.counters + 1, 1u);
    __syncthreads();
    const uint32_t b = s_slice;
    if (b >= p.n_slices) break;
    uint32_t cn = p.qcursor[b];
    const uint32_t cap = p.qcap[b];
    if (cn > cap) cn = cap;  // both are multiples of FQ_CHUNK
    if (!cn) continue;
    const uint4 *sv = reinterpret_cast<const uint4 *>(p.queue + p.qbase[b]);
    const uint32_t nvec = cn >> 3;
    uint32_t v = tid;
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
int kpc_fq_lo_bits(int k) { return kpc_fq_log_bins(k) / 2; }
uint32_t kpc_fq_queue_slack() { return FQ_CHUNK * 512u; }  // padding entries per slice: < one chunk per CTA (<= 2 per SM)
bool kpc_fq_supported(int k, int content) {
  return (content == KPC_CONTENT_DNA_SS || content == KPC_CONTENT_DNA_DS) && k >= 4 && k <= 12;
}

static cudaStream_t fq_cs(rt_stream s) { return (cudaStream_t)rt_stream_native(s); }

template <bool DS, int KT>
static void launch_partition(const KpcFqLaunch &L, rt_stream s) {
  auto kern = fq_partition_kernel<DS, KT>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    FQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FqSmem)));
    FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, FQ_NT, sizeof(FqSmem)));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
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
  if (L.k == 12) { if (ds) launch_partition<true, 12>(L, s); else launch_partition<false, 12>(L, s); }
  else { if (ds) launch_partition<true, 0>(L, s); else launch_partition<false, 0>(L, s); }
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
}
