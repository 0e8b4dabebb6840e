// kpc_kernels.cu -- hand-written sm_100a kernels of the KPopCount hot path and their launchers.
//
//   tiles_kernel      framing + linting + rolling canonical k-mers + sink   (kpc_tile.cuh; one pass over HBM)
//   scan family       block sums -> partial scan -> scatter; used by the compaction / formatter / RLE kernels
//   dense_*           fold / promote / max / extract for the 4^k table
//   format            "%0Wx\t%d\n" text lines on the device
//   hash_*            clear / rehash / extract for the open-addressing table
//   order / tuple     emission order of OCaml's Hashtbl.iter (radix sorts are cub::DeviceRadixSort: library code,
//                     off the counting hot path)
//   synth_fastq       deterministic synthetic reads for the benchmark
//
// Tensor-core free by nature (integer scatter / count work); the roofline is HBM bytes, see DESIGN.md.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include "kpc_kernels.h"
#include "kpc_bucketsort.cuh"
#include "kpc_synth.h"
#include "../../include/kpopcount.h"

#define CUDA_CHECK(x)                                                                                         \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      throw KpcError(KPC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +        \
                                     std::to_string(__LINE__) + ")");                                         \
  } while (0)

static inline cudaStream_t cs(rt_stream s) { return (cudaStream_t)rt_stream_native(s); }

// =================================================================================================
// framing + k-mer kernel
// =================================================================================================
#ifndef KPC_TILE_NT
#define KPC_TILE_NT 256
#endif
#ifndef KPC_TILE_SEG
#define KPC_TILE_SEG 64
#endif

template <int NT, int SEG, int FMT, int CONTENT, class Sink>
__global__ void __launch_bounds__(NT) tiles_kernel(const KpcTileParams p, const Sink sink) {
  typedef KpcTileMachine<NT, SEG, FMT, CONTENT> M;
  __shared__ typename M::Shared sh;
  const int tid = threadIdx.x;
  for (;;) {
    M::claim(sh, p, tid);
    __syncthreads();
    if (sh.tile >= p.n_tiles) break;
    M::load(sh, p, tid);
    __syncthreads();
    M::census(sh, p, tid);
    __syncthreads();
    M::scan_groups(sh, tid);
    __syncthreads();
    M::lookback1(sh, p, tid);
    __syncthreads();
    M::classify(sh, p, tid);
    __syncthreads();
    M::ksummary(sh, p, tid);
    __syncthreads();
    M::kmers(sh, p, sink, tid);
    M::fixup(sh, p, sink, tid);
    __syncthreads();
  }
}

uint32_t kpc_k_tile_bytes() { return KPC_TILE_NT * KPC_TILE_SEG; }

template <int FMT, int CONTENT, class Sink>
static void launch_tiles_t(const KpcTileParams &p, const Sink &sink, rt_stream s) {
  auto kern = tiles_kernel<KPC_TILE_NT, KPC_TILE_SEG, FMT, CONTENT, Sink>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, KPC_TILE_NT, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  long long grid = (long long)rt_sm_count() * blocks_per_sm;
  if (grid > (long long)p.n_tiles) grid = p.n_tiles;
  if (grid < 1) return;
  kern<<<(unsigned)grid, KPC_TILE_NT, 0, cs(s)>>>(p, sink);
  CUDA_CHECK(cudaGetLastError());
}
template <int FMT, int CONTENT>
static void launch_tiles_s(const KpcTileLaunch &L, rt_stream s) {
  switch (L.sink) {
    case KPC_SINK_DENSE: launch_tiles_t<FMT, CONTENT>(L.p, L.dense, s); break;
    case KPC_SINK_HASH: launch_tiles_t<FMT, CONTENT>(L.p, L.hash, s); break;
    case KPC_SINK_TUPLE: launch_tiles_t<FMT, CONTENT>(L.p, L.tuple, s); break;
    case KPC_SINK_BCOUNT: launch_tiles_t<FMT, CONTENT>(L.p, L.bcount, s); break;
    case KPC_SINK_BSCATTER: launch_tiles_t<FMT, CONTENT>(L.p, L.bscatter, s); break;
    default: launch_tiles_t<FMT, CONTENT>(L.p, KpcNullSink(), s); break;
  }
}
template <int FMT>
static void launch_tiles_c(const KpcTileLaunch &L, rt_stream s) {
  switch (L.content) {
    case KPC_CONTENT_DNA_SS: launch_tiles_s<FMT, KPC_CONTENT_DNA_SS>(L, s); break;
    case KPC_CONTENT_DNA_DS: launch_tiles_s<FMT, KPC_CONTENT_DNA_DS>(L, s); break;
    default: launch_tiles_s<FMT, KPC_CONTENT_PROTEIN>(L, s); break;
  }
}
void kpc_k_tiles(const KpcTileLaunch &L, rt_stream s) {
  if (L.fmt == KPC_FMT_FASTQ) launch_tiles_c<KPC_FMT_FASTQ>(L, s);
  else launch_tiles_c<KPC_FMT_FASTA>(L, s);
}

// =================================================================================================
// newline census
// =================================================================================================
__global__ void count_newlines_kernel(const uint8_t *d, uint64_t n, unsigned long long *out) {
  unsigned long long c = 0;
  const uint64_t nvec = n >> 4;
  const uint4 *v = reinterpret_cast<const uint4 *>(d);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 x = __ldg(v + i);
    c += __popc(__vcmpeq4(x.x, 0x0A0A0A0Au) & 0x01010101u) + __popc(__vcmpeq4(x.y, 0x0A0A0A0Au) & 0x01010101u) +
         __popc(__vcmpeq4(x.z, 0x0A0A0A0Au) & 0x01010101u) + __popc(__vcmpeq4(x.w, 0x0A0A0A0Au) & 0x01010101u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 15)) c += d[(nvec << 4) + threadIdx.x] == '\n';
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
void kpc_k_count_newlines(const uint8_t *d, uint64_t n, unsigned long long *out, rt_stream s) {
  if (!n) return;
  uint64_t blocks = (n / 16 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  count_newlines_kernel<<<(unsigned)blocks, 256, 0, cs(s)>>>(d, n, out);
  CUDA_CHECK(cudaGetLastError());
}

// =================================================================================================
// dense table helpers
// =================================================================================================
__global__ void dense_fold_kernel(uint32_t *lo, unsigned long long *hi, uint64_t nbins) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nbins; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = lo[i];
    if (v >= 0x80000000u) { hi[i] += v; lo[i] = 0; }
  }
}
__global__ void dense_promote_kernel(uint32_t *lo, unsigned long long *hi, uint64_t nbins) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nbins; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = lo[i];
    if (v) { hi[i] += v; lo[i] = 0; }
  }
}
__global__ void dense_max_kernel(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins,
                                 unsigned long long *out) {
  unsigned long long m = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nbins; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long v = (unsigned long long)lo[i] + (hi ? hi[i] : 0ull);
    m = v > m ? v : m;
  }
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
static unsigned ew_grid(uint64_t n) {
  uint64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (unsigned)(b ? b : 1);
}
void kpc_k_dense_fold(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream s) {
  dense_fold_kernel<<<ew_grid(nbins), 256, 0, cs(s)>>>(lo, hi, nbins);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void add_u64_kernel(unsigned long long *dst, const unsigned long long *src, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
void kpc_k_add_u64(unsigned long long *dst, const unsigned long long *src, uint64_t n, rt_stream s) {
  add_u64_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(dst, src, n);
  CUDA_CHECK(cudaGetLastError());
}
void kpc_k_dense_promote(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream s) {
  dense_promote_kernel<<<ew_grid(nbins), 256, 0, cs(s)>>>(lo, hi, nbins);
  CUDA_CHECK(cudaGetLastError());
}
void kpc_k_dense_max(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *out,
                     rt_stream s) {
  dense_max_kernel<<<ew_grid(nbins), 256, 0, cs(s)>>>(lo, hi, nbins, out);
  CUDA_CHECK(cudaGetLastError());
}

// =================================================================================================
// scan family: every item i has a length len(i) >= 0; write(i, offset, len) receives the exclusive
// prefix sum of the lengths.  Three launches: block sums, scan of the block sums, scatter.
// =================================================================================================
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total) {
  __shared__ unsigned long long warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protects warp_tot against the previous use
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    unsigned long long t = warp_tot[w];
    if (w < wid) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

template <class LenF>
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(LenF len, uint64_t n, unsigned long long *partial) {
  const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j)
    if (i0 + j < n) s += len(i0 + j);
  unsigned long long tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// in-place exclusive scan of partial[0..np); partial[np] = total
__global__ void __launch_bounds__(1024) scan_partials_kernel(unsigned long long *partial, uint64_t np) {
  __shared__ unsigned long long wt[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint64_t b = 0; b < np; b += 1024) {
    uint64_t i = b + threadIdx.x;
    unsigned long long v = i < np ? partial[i] : 0ull, inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wt[wid] = inc;
    __syncthreads();
    unsigned long long base = carry_s, tot = 0;
    for (int w = 0; w < 32; ++w) {
      unsigned long long t = wt[w];
      if (w < wid) base += t;
      tot += t;
    }
    if (i < np) partial[i] = base + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[np] = carry_s;
}
template <class LenF, class WriteF>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_scatter_kernel(LenF len, WriteF write, uint64_t n, const unsigned long long *partial) {
  const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long l[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    l[j] = (i0 + j < n) ? len(i0 + j) : 0ull;
    s += l[j];
  }
  unsigned long long tot;
  unsigned long long off = partial[blockIdx.x] + block_exclusive_scan(s, &tot);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (i0 + j < n) write(i0 + j, off, l[j]);
    off += l[j];
  }
}
size_t kpc_k_scan_scratch_bytes(uint64_t n) { return ((n + SCAN_BLK - 1) / SCAN_BLK + 2) * sizeof(unsigned long long) + 512; }

// total (number of output units) lands in *total_out
template <class LenF, class WriteF>
static void run_scan(LenF len, WriteF write, uint64_t n, unsigned long long *total_out, void *scratch, rt_stream s) {
  unsigned long long *partial = (unsigned long long *)scratch;
  const uint64_t nb = (n + SCAN_BLK - 1) / SCAN_BLK;
  if (nb == 0) {
    CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(unsigned long long), cs(s)));
    return;
  }
  scan_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, cs(s)>>>(len, n, partial);
  CUDA_CHECK(cudaGetLastError());
  scan_partials_kernel<<<1, 1024, 0, cs(s)>>>(partial, nb);
  CUDA_CHECK(cudaGetLastError());
  scan_scatter_kernel<<<(unsigned)nb, SCAN_THREADS, 0, cs(s)>>>(len, write, n, partial);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(total_out, partial + nb, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, cs(s)));
}

// ---- dense extract ----
struct DenseLen {
  const uint32_t *lo;
  const unsigned long long *hi;
  __device__ unsigned long long operator()(uint64_t i) const {
    return ((unsigned long long)lo[i] + (hi ? hi[i] : 0ull)) != 0ull;
  }
};
struct DenseWrite {
  const uint32_t *lo;
  const unsigned long long *hi;
  unsigned long long *keys, *counts;
  __device__ void operator()(uint64_t i, unsigned long long off, unsigned long long l) const {
    if (l) { keys[off] = i; counts[off] = (unsigned long long)lo[i] + (hi ? hi[i] : 0ull); }
  }
};
void kpc_k_dense_extract(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *keys,
                         unsigned long long *counts, unsigned long long *n_out, void *scratch, rt_stream s) {
  run_scan(DenseLen{lo, hi}, DenseWrite{lo, hi, keys, counts}, nbins, n_out, scratch, s);
}

// ---- formatter ----
// to_hex + Printf "%s\t%d\n" (KMers.ml:270-271, bin/KPopCount.ml:46,60): every entry becomes hex(key) TAB decimal(count) LF.
// A block formats its 2048 entries into shared memory (placed so that shared and global addresses agree mod 16) and
// copies the text out with 128-bit stores.
__device__ __forceinline__ int dec_digits(unsigned long long v) {
  if (v <= 0xFFFFFFFFull) {
    const uint32_t x = (uint32_t)v;
    return 1 + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u) + (x >= 100000u) + (x >= 1000000u) +
           (x >= 10000000u) + (x >= 100000000u) + (x >= 1000000000u);
  }
  int d = 1;
  while (v >= 10ull) { v /= 10ull; ++d; }
  return d;
}
struct FormatLen {
  const unsigned long long *counts;
  int w;
  __device__ unsigned long long operator()(uint64_t i) const {  // an entry with count 0 (sort path filler) prints nothing
    const unsigned long long c = counts[i];
    return c ? (unsigned long long)(w + 2 + dec_digits(c)) : 0ull;
  }
};
__global__ void __launch_bounds__(SCAN_THREADS)
    format_scatter_kernel(const unsigned long long *keys, const unsigned long long *counts, int w, char *out, uint64_t n,
                          const unsigned long long *partial) {
  extern __shared__ __align__(16) char fmt_stage[];
  const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long c[SCAN_ITEMS];
  uint32_t l[SCAN_ITEMS];
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    c[j] = (i0 + j < n) ? counts[i0 + j] : 0ull;
    l[j] = (i0 + j < n && c[j]) ? (uint32_t)(w + 2 + dec_digits(c[j])) : 0u;
    s += l[j];
  }
  unsigned long long tot;
  uint32_t off = (uint32_t)block_exclusive_scan(s, &tot);
  char *g = out + partial[blockIdx.x];
  const uint32_t skew = (uint32_t)((uintptr_t)g & 15u);
  char *st = fmt_stage + skew;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (l[j]) {
      char *o = st + off;
      unsigned long long key = keys[i0 + j];
      for (int q = w - 1; q >= 0; --q) { o[q] = "0123456789abcdef"[key & 15ull]; key >>= 4; }
      o[w] = '\t';
      const int nd = (int)l[j] - w - 2;
      if (c[j] <= 0xFFFFFFFFull) {
        uint32_t v = (uint32_t)c[j];
        for (int q = nd - 1; q >= 0; --q) { const uint32_t d = v / 10u; o[w + 1 + q] = (char)('0' + (v - d * 10u)); v = d; }
      } else {
        unsigned long long v = c[j];
        for (int q = nd - 1; q >= 0; --q) { o[w + 1 + q] = (char)('0' + (int)(v % 10ull)); v /= 10ull; }
      }
      o[l[j] - 1] = '\n';
      off += l[j];
    }
  }
  __syncthreads();
  const uint32_t total = (uint32_t)tot;
  const uint32_t head = min(total, (16u - skew) & 15u);
  if (threadIdx.x < head) g[threadIdx.x] = st[threadIdx.x];
  const uint32_t nvec = (total - head) >> 4;
  const uint4 *sv = reinterpret_cast<const uint4 *>(st + head);
  uint4 *gv = reinterpret_cast<uint4 *>(g + head);
  for (uint32_t v = threadIdx.x; v < nvec; v += SCAN_THREADS) gv[v] = sv[v];
  const uint32_t done = head + (nvec << 4);
  if (threadIdx.x < total - done) g[done + threadIdx.x] = st[done + threadIdx.x];
}
void kpc_k_format(const unsigned long long *keys, const unsigned long long *counts, uint64_t n, int hex_width,
                  char *out, unsigned long long *out_len, void *scratch, rt_stream s) {
  unsigned long long *partial = (unsigned long long *)scratch;
  const uint64_t nb = (n + SCAN_BLK - 1) / SCAN_BLK;
  if (nb == 0) {
    CUDA_CHECK(cudaMemsetAsync(out_len, 0, sizeof(unsigned long long), cs(s)));
    return;
  }
  const size_t smem = (size_t)SCAN_BLK * (size_t)(hex_width + 22) + 32;  // 20 decimal digits at most
  static size_t smem_set_dev[64] = {0};
  size_t &smem_set = smem_set_dev[rt_current_device() & 63];
  if (smem > smem_set) {
    CUDA_CHECK(cudaFuncSetAttribute(format_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  scan_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, cs(s)>>>(FormatLen{counts, hex_width}, n, partial);
  CUDA_CHECK(cudaGetLastError());
  scan_partials_kernel<<<1, 1024, 0, cs(s)>>>(partial, nb);
  CUDA_CHECK(cudaGetLastError());
  format_scatter_kernel<<<(unsigned)nb, SCAN_THREADS, smem, cs(s)>>>(keys, counts, hex_width, out, n, partial);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(out_len, partial + nb, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, cs(s)));
}

// =================================================================================================
// hash table helpers
// =================================================================================================
__global__ void hash_clear_kernel(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks,
                                  uint64_t cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
    keys[i] = ~0ull; counts[i] = 0ull; ranks[i] = ~0ull;
  }
}
void kpc_k_hash_clear(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks, uint64_t cap,
                      rt_stream s) {
  hash_clear_kernel<<<ew_grid(cap), 256, 0, cs(s)>>>(keys, counts, ranks, cap);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void hash_rehash_kernel(const unsigned long long *okeys, const unsigned long long *ocounts,
                                   const unsigned long long *oranks, uint64_t ocap, KpcHashSink nw) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < ocap; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long key = okeys[i];
    if (key == ~0ull) continue;
    nw.insert(key, oranks[i], ocounts[i], true);
  }
}
void kpc_k_hash_rehash(const unsigned long long *okeys, const unsigned long long *ocounts,
                       const unsigned long long *oranks, uint64_t ocap, KpcHashSink nw, rt_stream s) {
  hash_rehash_kernel<<<ew_grid(ocap), 256, 0, cs(s)>>>(okeys, ocounts, oranks, ocap, nw);
  CUDA_CHECK(cudaGetLastError());
}
struct HashLen {
  const unsigned long long *keys, *counts;
  __device__ unsigned long long operator()(uint64_t i) const { return keys[i] != ~0ull && (long long)counts[i] > 0; }
};
struct HashWrite {
  const unsigned long long *keys, *counts, *ranks;
  unsigned long long *okeys, *ocounts, *oranks;
  __device__ void operator()(uint64_t i, unsigned long long off, unsigned long long l) const {
    if (l) { okeys[off] = keys[i]; ocounts[off] = counts[i]; oranks[off] = ranks[i]; }
  }
};
void kpc_k_hash_extract(const unsigned long long *keys, const unsigned long long *counts,
                        const unsigned long long *ranks, uint64_t cap, unsigned long long *okeys,
                        unsigned long long *ocounts, unsigned long long *oranks, unsigned long long *n_out,
                        void *scratch, rt_stream s) {
  run_scan(HashLen{keys, counts}, HashWrite{keys, counts, ranks, okeys, ocounts, oranks}, cap, n_out, scratch, s);
}

// =================================================================================================
// sort path: bucket offsets + per-bucket finalize (kpc_bucketsort.cuh)
// =================================================================================================
struct BucketLen {
  const uint32_t *hist;
  __device__ unsigned long long operator()(uint64_t i) const { return hist[i]; }
};
struct BucketWrite {
  uint32_t *offsets;
  __device__ void operator()(uint64_t i, unsigned long long off, unsigned long long) const { offsets[i] = (uint32_t)off; }
};
__global__ void bucket_total_kernel(const unsigned long long *total, uint32_t *offsets, uint32_t nb) {
  offsets[nb] = (uint32_t)*total;
}
void kpc_k_bucket_offsets(const uint32_t *hist, uint32_t nb, uint32_t *offsets, void *scratch, rt_stream s) {
  unsigned long long *total = (unsigned long long *)scratch;
  run_scan(BucketLen{hist}, BucketWrite{offsets}, nb, total, (char *)scratch + 256, s);
  bucket_total_kernel<<<1, 1, 0, cs(s)>>>(total, offsets, nb);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void __launch_bounds__(256) bucket_scatter_staged_kernel(const KpcPair *stage, const unsigned long long *n,
                                                                    const KpcBucketScatterSink sink) {
  const unsigned long long N = *n, stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const KpcPair pr = stage[i];
    sink.emit(pr.key, pr.rank, 0);
  }
}
void kpc_k_bucket_scatter_staged(const KpcPair *stage, const unsigned long long *n, const KpcBucketScatterSink &sink,
                                 rt_stream s) {
  bucket_scatter_staged_kernel<<<148 * 8, 256, 0, cs(s)>>>(stage, n, sink);
  CUDA_CHECK(cudaGetLastError());
}
extern __shared__ __align__(16) uint8_t kpc_dyn_smem[];
__global__ void __launch_bounds__(256) bucket_finalize_small_kernel(const KpcBucketFinalize F) {
  kpc_bucket_finalize_small_body<256>(F);
}
__global__ void __launch_bounds__(256) bucket_finalize_heavy_kernel(const KpcBucketFinalize F) {
  kpc_bucket_finalize_heavy_body<256>(F, kpc_dyn_smem);
}
void kpc_k_bucket_finalize(const KpcBucketFinalize &F, rt_stream s) {
  static bool attr_dev[64] = {false};
  bool &attr = attr_dev[rt_current_device() & 63];
  if (!attr) {
    CUDA_CHECK(cudaFuncSetAttribute(bucket_finalize_heavy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(KpcBsHeavySmem)));
    attr = true;
  }
  bucket_finalize_small_kernel<<<ew_grid(F.nb), 256, 0, cs(s)>>>(F);
  CUDA_CHECK(cudaGetLastError());
  // the heavy groups (usually none): a fixed grid walks the list the first kernel left
  bucket_finalize_heavy_kernel<<<148, 256, sizeof(KpcBsHeavySmem), cs(s)>>>(F);
  CUDA_CHECK(cudaGetLastError());
}

// =================================================================================================
// ordering (OCaml Hashtbl.iter order) and -L tuple reduction
// =================================================================================================
static size_t cub_sort_temp_bytes(uint64_t n) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  size_t t2 = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, t2, (const unsigned long long *)nullptr,
                                            (unsigned long long *)nullptr, (const uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int)n);
  return (t > t2 ? t : t2) + 256;
}
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
// scratch layout: ka[n] kb[n] (u64) | ia[n] ib[n] (u32) | tmp[n] (u64) | cub temp | scan scratch
size_t kpc_k_order_scratch_bytes(uint64_t n) {
  if (n == 0) n = 1;
  return al(n * 8) * 3 + al(n * 4) * 2 + al(cub_sort_temp_bytes(n)) + al(kpc_k_scan_scratch_bytes(n)) + 1024;
}
struct OrderScratch {
  unsigned long long *ka, *kb, *tmp;
  uint32_t *ia, *ib;
  void *cub_temp;
  size_t cub_bytes;
  void *scan;
  OrderScratch(void *scratch, uint64_t n) {
    if (n == 0) n = 1;
    char *p = (char *)scratch;
    ka = (unsigned long long *)p; p += al(n * 8);
    kb = (unsigned long long *)p; p += al(n * 8);
    tmp = (unsigned long long *)p; p += al(n * 8);
    ia = (uint32_t *)p; p += al(n * 4);
    ib = (uint32_t *)p; p += al(n * 4);
    cub_bytes = cub_sort_temp_bytes(n);
    cub_temp = p; p += al(cub_bytes);
    scan = p;
  }
};
__global__ void iota_kernel(uint32_t *idx, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    idx[i] = (uint32_t)i;
}
__global__ void gather_u64_kernel(const unsigned long long *src, const uint32_t *idx, unsigned long long *dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
__global__ void gather_u32_kernel(const uint32_t *src, const uint32_t *idx, uint32_t *dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
__global__ void bucket_key_kernel(const unsigned long long *keys, const uint32_t *recs, const uint32_t *idx,
                                  unsigned long long bmask, const unsigned long long *bmask_per_rec,
                                  uint32_t rec_off, unsigned long long *dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t j = idx[i];
    unsigned long long m = (recs && bmask_per_rec) ? bmask_per_rec[recs[j] - rec_off] : bmask;
    dst[i] = keys[j] & m;
  }
}
__global__ void widen_u32_kernel(const uint32_t *src, const uint32_t *idx, unsigned long long *dst, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
static void permute_u64(unsigned long long *a, const uint32_t *idx, unsigned long long *tmp, uint64_t n, rt_stream s) {
  gather_u64_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(a, idx, tmp, n);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(a, tmp, n * 8, cudaMemcpyDeviceToDevice, cs(s)));
}
static void permute_u32(uint32_t *a, const uint32_t *idx, uint32_t *tmp, uint64_t n, rt_stream s) {
  gather_u32_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(a, idx, tmp, n);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(a, tmp, n * 4, cudaMemcpyDeviceToDevice, cs(s)));
}
// stable sort of (sort key ka, permutation ia) -> (kb, ib); returns with the result in (ka, ia) again
static void sort_step(OrderScratch &S, uint64_t n, bool descending, int end_bit, rt_stream s) {
  size_t tb = S.cub_bytes;
  if (descending)
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(S.cub_temp, tb, S.ka, S.kb, S.ia, S.ib, (int)n, 0, end_bit, cs(s)));
  else
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(S.cub_temp, tb, S.ka, S.kb, S.ia, S.ib, (int)n, 0, end_bit, cs(s)));
  std::swap(S.ka, S.kb);
  std::swap(S.ia, S.ib);
}
void kpc_k_order_entries(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks,
                         uint32_t *recs, uint64_t n, uint64_t bmask, const unsigned long long *bmask_per_rec,
                         uint32_t rec_off, void *scratch, rt_stream s) {
  if (n == 0) return;
  if (n >= 0x7fffffffull) throw KpcError(KPC_E_UNSUPPORTED, "more than 2^31 distinct k-mers in one dump");
  OrderScratch S(scratch, n);
  iota_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(S.ia, n);
  CUDA_CHECK(cudaGetLastError());
  // 1. newest first: descending rank
  CUDA_CHECK(cudaMemcpyAsync(S.ka, ranks, n * 8, cudaMemcpyDeviceToDevice, cs(s)));
  sort_step(S, n, true, 64, s);
  // 2. stable by bucket index (key mod B)
  bucket_key_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(keys, recs, S.ia, bmask, bmask_per_rec, rec_off, S.ka, n);
  CUDA_CHECK(cudaGetLastError());
  sort_step(S, n, false, 64, s);
  // 3. stable by record
  if (recs) {
    widen_u32_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(recs, S.ia, S.ka, n);
    CUDA_CHECK(cudaGetLastError());
    sort_step(S, n, false, 32, s);
  }
  permute_u64(keys, S.ia, S.tmp, n, s);
  permute_u64(counts, S.ia, S.tmp, n, s);
  permute_u64(ranks, S.ia, S.tmp, n, s);
  if (recs) permute_u32(recs, S.ia, (uint32_t *)S.tmp, n, s);
}

struct HeadLen {
  const unsigned long long *keys;
  const uint32_t *recs;
  __device__ unsigned long long operator()(uint64_t i) const {
    return i == 0 || keys[i] != keys[i - 1] || recs[i] != recs[i - 1];
  }
};
struct HeadWrite {
  const unsigned long long *keys, *ranks;
  const uint32_t *recs;
  unsigned long long *okeys, *ocounts, *oranks;
  uint32_t *orecs;
  __device__ void operator()(uint64_t i, unsigned long long off, unsigned long long l) const {
    unsigned long long e = l ? off : off - 1;  // entry this tuple belongs to
    if (l) { okeys[e] = keys[i]; orecs[e] = recs[i]; }
    atomicAdd(ocounts + e, 1ull);
    atomicMin(oranks + e, ranks[i]);
  }
};
__global__ void fill_u64_kernel(unsigned long long *a, unsigned long long v, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) a[i] = v;
}
__device__ static unsigned long long lower_bound_u32(const uint32_t *a, unsigned long long n, uint32_t v) {
  unsigned long long lo = 0, hi = n;
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void __launch_bounds__(256) tuple_bounds_kernel(const uint32_t *trecs, unsigned long long n, const uint32_t *erec,
                                                           unsigned long long n_entries, uint32_t g_lo, uint32_t g_hi,
                                                           unsigned long long *bounds, unsigned long long *per_rec) {
  if (blockIdx.x == 0 && threadIdx.x < 2)
    bounds[threadIdx.x] = threadIdx.x == 0 ? lower_bound_u32(trecs, n, g_hi) : lower_bound_u32(erec, n_entries, g_hi);
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += stride) {
    const uint32_t r = erec[i];
    if (r >= g_lo && r < g_hi) atomicAdd(per_rec + (r - g_lo), 1ull);
  }
}
void kpc_k_tuple_bounds(const uint32_t *trecs, uint64_t n, const uint32_t *erec, uint64_t n_entries, uint32_t g_lo,
                        uint32_t g_hi, unsigned long long *bounds, unsigned long long *per_rec, rt_stream s) {
  CUDA_CHECK(cudaMemsetAsync(per_rec, 0, (size_t)(g_hi - g_lo) * sizeof(unsigned long long), cs(s)));
  tuple_bounds_kernel<<<ew_grid(n_entries ? n_entries : 1), 256, 0, cs(s)>>>(trecs, n, erec, n_entries, g_lo, g_hi, bounds, per_rec);
  CUDA_CHECK(cudaGetLastError());
}
void kpc_k_tuple_reduce(unsigned long long *keys, unsigned long long *ranks, uint32_t *recs, uint64_t n,
                        unsigned long long *okeys, unsigned long long *ocounts, unsigned long long *oranks,
                        uint32_t *orecs, unsigned long long *n_out, void *scratch, rt_stream s) {
  if (n == 0) {
    CUDA_CHECK(cudaMemsetAsync(n_out, 0, 8, cs(s)));
    return;
  }
  if (n >= 0x7fffffffull) throw KpcError(KPC_E_UNSUPPORTED, "more than 2^31 k-mer windows in one -L batch");
  OrderScratch S(scratch, n);
  iota_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(S.ia, n);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(S.ka, keys, n * 8, cudaMemcpyDeviceToDevice, cs(s)));
  sort_step(S, n, false, 64, s);
  widen_u32_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(recs, S.ia, S.ka, n);
  CUDA_CHECK(cudaGetLastError());
  sort_step(S, n, false, 32, s);
  permute_u64(keys, S.ia, S.tmp, n, s);
  permute_u64(ranks, S.ia, S.tmp, n, s);
  permute_u32(recs, S.ia, (uint32_t *)S.tmp, n, s);
  CUDA_CHECK(cudaMemsetAsync(ocounts, 0, n * 8, cs(s)));
  fill_u64_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(oranks, ~0ull, n);
  CUDA_CHECK(cudaGetLastError());
  run_scan(HeadLen{keys, recs}, HeadWrite{keys, ranks, recs, okeys, ocounts, oranks, orecs}, n, n_out, S.scan, s);
}
__global__ void rec_counts_kernel(const uint32_t *recs, uint64_t n, unsigned long long *cnt, uint64_t n_recs) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t r = recs[i];
    if (r < n_recs) atomicAdd(cnt + r, 1ull);
  }
}
void kpc_k_rec_counts(const uint32_t *recs, uint64_t n, unsigned long long *cnt, uint64_t n_recs, rt_stream s) {
  CUDA_CHECK(cudaMemsetAsync(cnt, 0, n_recs * 8, cs(s)));
  if (!n) return;
  rec_counts_kernel<<<ew_grid(n), 256, 0, cs(s)>>>(recs, n, cnt, n_recs);
  CUDA_CHECK(cudaGetLastError());
}

// =================================================================================================
// synthetic reads: one warp per record, lanes stride over its bytes (coalesced byte stores)
// =================================================================================================
__global__ void synth_fastq_kernel(uint8_t *out, uint64_t first, uint64_t nrec, uint64_t seed, uint64_t base_off) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (uint64_t r = warp; r < nrec; r += nwarps) {
    const uint64_t rec = first + r;
    const uint32_t nd = kpc_synth_digits(rec);
    const uint32_t len = 307u + nd;
    uint8_t *o = out + (kpc_synth_record_offset(rec) - base_off);
    for (uint32_t q = lane; q < len; q += 32) o[q] = kpc_synth_byte(seed, rec, nd, q);
  }
}
void kpc_k_synth_fastq(uint8_t *out, uint64_t first_record, uint64_t n_records, uint64_t seed, rt_stream s) {
  if (!n_records) return;
  uint64_t blocks = (n_records * 32 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  synth_fastq_kernel<<<(unsigned)blocks, 256, 0, cs(s)>>>(out, first_record, n_records, seed,
                                                          kpc_synth_record_offset(first_record));
  CUDA_CHECK(cudaGetLastError());
}
