// kpc_fastq.h -- launch interface of the fast FASTQ / dense-table pipeline (kpc_fastq.cu).
//
// The generic tile machine (kpc_tile.cuh) sends one RED per k-mer to the L2-resident 4^k table; on a B200 that
// caps the k = 12 path at ~190 G k-mers/s (profiles/r1_microbench_primitives.jsonl).  This pipeline replaces it
// for FASTQ + DNA + dense table, 8 <= k <= 12, with two kernels per launch of the byte stream:
//
//   fq_partition   record framing (Files.FASTQ.iter_se, Files.ml:201-221) + linting (Sequences.ml:41-67) +
//                  forward / reverse-complement k-mers and min (KMers.ml:357-389), 16 windows per thread with
//                  SIMD-in-register classification; every canonical key is appended to a shared-memory bucket of
//                  its "slice" (middle bits of the key) and whole 32-byte chunks of the buckets are appended,
//                  16 bits per k-mer, to per-slice queues in HBM (software write-combining)
//   fq_count       one CTA per slice: the slice of the table (2^15 u32 bins) lives in shared memory, queue entries
//                  are counted with shared-memory atomics, non-zero bins are added to the global table
//                  (IntHashFrequencies.add, KMers.ml:107-111)
//
// Queues have a fixed capacity per slice; k-mers that do not fit (heavily skewed inputs) are counted directly in
// the global table with RED, so the result is exact whatever the input.
#pragma once
#include "kpc_rt.h"
#include "kpc_tile.cuh"

struct KpcFqLaunch {
  const uint8_t *data;      // launch bytes; 16-byte aligned; readable up to the next 16-byte boundary past n
  uint64_t n;
  uint64_t abs_base;        // stream offset of data[0]
  int halo_ok;              // data[-16 .. 0) is readable and holds the 16 stream bytes before the launch
                            // (otherwise the launch must start at a line start)
  uint64_t max_lines;       // bytes on lines >= max_lines are ignored (incomplete last record, -p cap)
  int k;
  int content;              // KPC_CONTENT_DNA_SS / KPC_CONTENT_DNA_DS
  const KpcStreamCarry *carry_in;
  KpcStreamCarry *carry_out;
  unsigned long long *err_line;
  unsigned long long *tile_state;  // n_tiles words, zero on entry (newline-count look-back)
  uint32_t *counters;              // [0] tile claim counter, [1] slice claim counter: zero on entry
  uint32_t n_tiles;
  // slices and queues
  int log_bins;                    // bins per slice = 1 << log_bins (<= 15)
  int lo_bits, slice_bits;         // slice = key bits [lo_bits, lo_bits + slice_bits); bin = the other bits, packed
  uint32_t n_slices;               // 1 << slice_bits = 4^k >> log_bins  (128 or 512)
  uint16_t *queue;
  const unsigned long long *qbase; // first entry of every slice's queue (multiple of 16)
  const uint32_t *qcap;            // capacity of every slice's queue, in entries (multiple of 16)
  uint32_t *qcursor;               // entries appended so far (may exceed the capacity): zero on entry
  uint32_t *table;                 // the dense 4^k table
};

uint32_t kpc_fq_tile_bytes();
// bins per slice: at most 2^15 (one u16 queue entry, a 128 KiB shared-memory table) and at least 128 slices so
// that the counting kernel has enough CTAs; the slice index is cut out of the middle of the key
inline int kpc_fq_log_bins(int k) { return 2 * k - 7 < 15 ? 2 * k - 7 : 15; }
inline int kpc_fq_lo_bits(int k) { return k == 12 ? 8 : kpc_fq_log_bins(k) / 2; }  // k = 12: byte-aligned fields (fq_append_k12)
inline uint32_t kpc_fq_queue_slack() { return 16u * 512u; }  // padding entries per slice: < one chunk per CTA (<= 2 per SM)
inline bool kpc_fq_supported(int k, int content) {
  return (content == KPC_CONTENT_DNA_SS || content == KPC_CONTENT_DNA_DS) && k >= 4 && k <= 12;
}
void kpc_fq_partition(const KpcFqLaunch &L, rt_stream s);
void kpc_fq_count(const KpcFqLaunch &L, rt_stream s);
// optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline figure)
void kpc_fq_timing_enable(bool on);
void kpc_fq_timing_read(double *partition_ms, double *count_ms, unsigned long long *launches, unsigned long long *bytes);
