// kpc_kernels.h -- launch interface of the device kernels (kpc_kernels.cu).  All pointers are device
// pointers unless stated otherwise; every call is asynchronous on the given stream.
#pragma once
#include "kpc_rt.h"
#include "kpc_tile.cuh"

enum { KPC_SINK_NULL = 0, KPC_SINK_DENSE = 1, KPC_SINK_HASH = 2, KPC_SINK_TUPLE = 3, KPC_SINK_BCOUNT = 4, KPC_SINK_BSCATTER = 5 };

struct KpcTileLaunch {
  int fmt, content, sink;
  KpcTileParams p;
  KpcDenseSink dense;
  KpcHashSink hash;
  KpcTupleSink tuple;
  KpcBucketCountSink bcount;
  KpcBucketScatterSink bscatter;
};

// bytes per tile of the framing kernel (the engine sizes the descriptor array with it)
uint32_t kpc_k_tile_bytes();
// framing + rolling k-mers + sink over one launch (p.n bytes); p.tile_counter must be zero
void kpc_k_tiles(const KpcTileLaunch &L, rt_stream s);
// *out += number of '\n' in d[0..n)
void kpc_k_count_newlines(const uint8_t *d, uint64_t n, unsigned long long *out, rt_stream s);

// ---- dense table (k small) ----
// bins whose u32 counter reached 2^31 are moved into the u64 side table (exactness on very deep inputs)
void kpc_k_dense_fold(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream s);
// hi[i] += lo[i]; lo[i] = 0   (used before a 64-bit reduction across GPUs)
void kpc_k_dense_promote(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream s);
// dst[i] += src[i]   (sum of the 64-bit tables of two devices)
void kpc_k_add_u64(unsigned long long *dst, const unsigned long long *src, uint64_t n, rt_stream s);
// *out = max over bins of lo[i] + hi[i]
void kpc_k_dense_max(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *out,
                     rt_stream s);
// scratch needed by the scan-based kernels below for n items
size_t kpc_k_scan_scratch_bytes(uint64_t n);
// non-zero bins in ascending index order -> keys[], counts[]; *n_out = how many
void kpc_k_dense_extract(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *keys,
                         unsigned long long *counts, unsigned long long *n_out, void *scratch, rt_stream s);

// ---- text ----
// "%0Wx\t%d\n" per entry (bin/KPopCount.ml:46,60; KMers.ml:270-271); *out_len = bytes written.
// out must hold n * (hex_width + 22) bytes.
void kpc_k_format(const unsigned long long *keys, const unsigned long long *counts, uint64_t n, int hex_width,
                  char *out, unsigned long long *out_len, void *scratch, rt_stream s);

// ---- hash table (k large, or small -M) ----
void kpc_k_hash_clear(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks, uint64_t cap,
                      rt_stream s);
// re-insert every live entry of the old table into the (cleared) new one
void kpc_k_hash_rehash(const unsigned long long *okeys, const unsigned long long *ocounts,
                       const unsigned long long *oranks, uint64_t ocap, KpcHashSink nw, rt_stream s);
// entries with count > 0 -> (keys, counts, ranks); *n_out = how many (order unspecified)
void kpc_k_hash_extract(const unsigned long long *keys, const unsigned long long *counts,
                        const unsigned long long *ranks, uint64_t cap, unsigned long long *okeys,
                        unsigned long long *ocounts, unsigned long long *oranks, unsigned long long *n_out,
                        void *scratch, rt_stream s);

// ---- sort path (kpc_bucketsort.cuh): a whole large-k sample without a hash table ----
// exclusive prefix sums of hist[0..nb) into offsets[0..nb]; offsets[nb] = total
void kpc_k_bucket_offsets(const uint32_t *hist, uint32_t nb, uint32_t *offsets, void *scratch, rt_stream s);
// scatter pass over staged (key, rank) pairs (KpcBucketCountSink::stage_*): *n pairs, n read on the device
void kpc_k_bucket_scatter_staged(const KpcPair *stage, const unsigned long long *n, const KpcBucketScatterSink &sink,
                                 rt_stream s);
struct KpcBucketFinalize {
  const uint32_t *offsets;            // nb + 1 entries
  uint32_t nb;
  const KpcPair *pairs;               // in: (key, rank) pairs grouped by coarse bucket
  unsigned long long *keys, *ranks;   // out: entries
  unsigned long long *counts;         // out
  unsigned long long bmask;           // B - 1
  uint32_t *heavy_list;               // groups left to the CTA-wide kernel
  uint32_t heavy_cap;
  unsigned long long *stats;          // [0] distinct keys, [1] too-heavy flag, [2] number of heavy groups
};

// every bucket range of (keys, ranks): duplicates merged (counts = multiplicity, ranks = first occurrence), entries
// left in Hashtbl.iter order, the slots duplicates leave behind get key ~0 / count 0.  stats[0] += distinct keys,
// stats[1] != 0 when a range is too large for this path (the caller then falls back to the hash table).
void kpc_k_bucket_finalize(const KpcBucketFinalize &F, rt_stream s);

// ---- ordering ----
// OCaml Hashtbl.iter order: ascending (key mod B), newest (largest rank) first inside a
// bucket; recs (optional, may be null) is the major key for -L batches.  Sorts the arrays in place.
// B - 1 is bmask, or bmask_per_rec[recs[i] - rec_off] when that device array is given (B may grow between records).
size_t kpc_k_order_scratch_bytes(uint64_t n);
void kpc_k_order_entries(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks,
                         uint32_t *recs, uint64_t n, uint64_t bmask, const unsigned long long *bmask_per_rec,
                         uint32_t rec_off, void *scratch, rt_stream s);
// -L: sort tuples by (rec, key), then run-length: one entry per distinct (rec, key) with its count and min rank.
// *n_out = number of entries.  Output arrays must hold n items.
// -L bookkeeping on the device: bounds[0] = tuples (sorted by record) of records < g_hi, bounds[1] = entries of records
// < g_hi; per_rec[r - g_lo] = entries of record r for g_lo <= r < g_hi (per_rec zeroed here)
void kpc_k_tuple_bounds(const uint32_t *trecs, uint64_t n, const uint32_t *erec, uint64_t n_entries, uint32_t g_lo,
                        uint32_t g_hi, unsigned long long *bounds, unsigned long long *per_rec, rt_stream s);
void kpc_k_tuple_reduce(unsigned long long *keys, unsigned long long *ranks, uint32_t *recs, uint64_t n,
                        unsigned long long *okeys, unsigned long long *ocounts, unsigned long long *oranks,
                        uint32_t *orecs, unsigned long long *n_out, void *scratch, rt_stream s);
// per-record entry counts: cnt[r] = #entries with recs == r (r < n_recs); recs sorted ascending
void kpc_k_rec_counts(const uint32_t *recs, uint64_t n, unsigned long long *cnt, uint64_t n_recs, rt_stream s);

// ---- synthetic input (bench / tests): record i of the C3/C5 shape of SURVEY.md 8d ----
void kpc_k_synth_fastq(uint8_t *out, uint64_t first_record, uint64_t n_records, uint64_t seed, rt_stream s);
