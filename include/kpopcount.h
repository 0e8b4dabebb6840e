/* kpopcount.h -- C ABI of libkpopcount_gpu.so: the B200 back end of KPop's k-mer spectrum counting stage.
 *
 * The reference (PaoloRibeca/KPop @ a1fda68) is pure OCaml and has no FFI on this path; the seam this
 * library replaces is the functor application in bin/KPopCount.ml:
 *
 *     KMerCounter (KIH : KMers.IntHash_t) . compute ~linter store max_results_size label fname
 *                                                                         (bin/KPopCount.ml:20-64)
 * i.e. all of   Files.ReadsIterate.iter   (BiOCamLib/lib/Files.ml:350-368; FASTA.iter :96-122,
 *                                           FASTQ.iter_se :201-221, iter_pe :222-250)
 *               Sequences.Lint.dnaize / proteinize   (BiOCamLib/lib/Sequences.ml:41-67, 87-151)
 *               KIH.iterc                 (BiOCamLib/lib/KMers.ml:228-257, 319-349, 357-389)
 *               IntHashFrequencies.create/add/length/iter/clear   (KMers.ml:99-113, Better.ml:700-705,741)
 *               KIH.to_hex + Printf "%s\t%d\n"       (KMers.ml:151,270-271; bin/KPopCount.ml:34,45,46,60)
 * taken together, because record splitting happens on the device.  The OCaml host keeps argv handling
 * (bin/KPopCount.ml:105-238) and file I/O, streams raw file bytes into kpc_feed() and receives the text of
 * the spectra, in the reference's order, through the sink callback.  INTEGRATION.md shows the `external`
 * declarations and C stubs a maintainer would add; kpop_b200/csrc/kpopcount_main.cpp is the same host logic
 * in C++ (there is no OCaml toolchain in this image).
 *
 * Conventions: every function returns KPC_OK (0) or a negative KPC_E_* code; kpc_error() gives the message
 * of the last failure on that context.  No C++ exception crosses this boundary.  One caller thread per
 * context.  All sizes are bytes.  Output text handed to the sink is only valid during the callback.
 */
#ifndef KPOPCOUNT_H
#define KPOPCOUNT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kpc_ctx kpc_ctx; /* opaque; one per KPopCount invocation (one KMerCounter.compute call) */

/* -C / --content (bin/KPopCount.ml:66-82) */
enum { KPC_DNA_SS = 0, KPC_DNA_DS = 1, KPC_PROTEIN = 2 };
/* -f / -s / -p (Files.Type, BiOCamLib/lib/Files.ml:315-323) */
enum { KPC_FASTA = 0, KPC_FASTQ_SE = 1, KPC_FASTQ_PE = 2 };

enum {
  KPC_OK = 0,
  KPC_E_ARG = -1,             /* bad argument to this API (CLI: exit 1) */
  KPC_E_K_RANGE = -2,         /* k > 30 (DNA) or > 12 (protein): Failure in KMers.ml:264-267 / 145-148 (exit 2) */
  KPC_E_MALFORMED_FASTQ = -3, /* Files.ml:213-214 / 241-243, or an empty tag / '+' line (exit 2) */
  KPC_E_QUOTES_IN_NAME = -4,  /* Matrix.Base.Quotes_in_name, BiOCamLib/lib/Matrix.ml:83-99 (exit 2 in -L mode) */
  KPC_E_IO = -5,              /* the sink reported a failure */
  KPC_E_CUDA = -6,            /* CUDA error, or no usable GPU: there is NO CPU fallback */
  KPC_E_NOMEM = -7,
  KPC_E_STATE = -8,           /* calls out of order */
  KPC_E_UNSUPPORTED = -9,     /* an input shape this implementation refuses rather than get wrong */
  KPC_E_PE_MISMATCH = -10     /* -p files hold different numbers of records: call kpc_set_pair_limit and re-run */
};

/* receives output text in final order; return 0 to continue, non-zero to abort with KPC_E_IO */
typedef int (*kpc_sink_fn)(void *user, const char *bytes, size_t n);

/* KIHF.create max_results_size + the functor's k check (bin/KPopCount.ml:35,239-249).
 * label "" selects -L behaviour exactly as in the reference (bin/KPopCount.ml:39,44).
 * n_devices > 1 (device_ids lists them): FASTQ inputs of dense-table runs (k <= 12 for DNA, one label) are cut into chunks
 * at line starts and spread over the devices; kpc_finish adds the tables up on device_ids[0] over peer copies.  Anything
 * else (FASTA, whose k-mers span lines; large k; -L) runs on device_ids[0] alone: results never depend on n_devices. */
int kpc_create(kpc_ctx **out, int k, int content, long long max_results_size, const char *label, int n_devices,
               const int *device_ids);
void kpc_destroy(kpc_ctx *ctx);
const char *kpc_error(const kpc_ctx *ctx);

/* where the text goes (stdout / open_out fname in bin/KPopCount.ml:27-31).  Must be set before kpc_begin:
 * the label header is written when the first input starts (:33-34) and -L / spill dumps happen while feeding. */
int kpc_set_sink(kpc_ctx *ctx, kpc_sink_fn sink, void *user);
/* alternative to a callback: the text is written (device -> host copies included) straight into a caller-owned host
 * buffer, e.g. a pinned Bigarray; NULL switches back to the callback.  kpc_reset rewinds it. */
int kpc_set_sink_buffer(kpc_ctx *ctx, void *host_buffer, size_t capacity);
unsigned long long kpc_sink_buffer_used(const kpc_ctx *ctx);

/* pinned host buffers owned by the context, for zero-copy reads (slot in [0, kpc_staging_slots)).
 * Blocks until the previous kpc_feed from that slot has left the buffer. */
int kpc_staging_slots(const kpc_ctx *ctx);
void *kpc_staging(kpc_ctx *ctx, int slot, size_t *capacity);

/* one element of Files.ReadsIterate.t (a file, or a pair of files), in argv order */
int kpc_begin(kpc_ctx *ctx, int format);
/* raw file bytes of mate 0 (or 1 for the second file of -p); eof != 0 with the last call for that mate.
 * bytes may point into a kpc_staging buffer (asynchronous) or anywhere else (consumed before returning). */
int kpc_feed(kpc_ctx *ctx, int mate, const void *bytes, size_t n, int eof);
/* same, for bytes that already live in device memory (16-byte aligned); used by benchmarks.  The whole input in one call
 * (eof != 0).  Dense-table runs take any size; hash-table runs (large k) take inputs up to the staging size (64 MiB). */
int kpc_feed_device(kpc_ctx *ctx, int mate, const void *device_bytes, size_t n, int eof);
/* FASTQ.iter_pe stops at the shorter file (Files.ml:228-247): after KPC_E_PE_MISMATCH, create a fresh context,
 * set the limit to the reported number of complete pairs and feed again */
int kpc_set_pair_limit(kpc_ctx *ctx, long long n_pairs);
long long kpc_complete_pairs(const kpc_ctx *ctx);
/* inputs that cannot be read twice (pipes): with on != 0 paired files are woven pair by pair on the host in EVERY mode
 * (otherwise only where the reference's order matters: large k, -L, -M spills), so a shorter mate file ends the input
 * exactly as FASTQ.iter_pe does and KPC_E_PE_MISMATCH never arises.  Call before kpc_begin. */
int kpc_set_single_pass(kpc_ctx *ctx, int on);
int kpc_end(kpc_ctx *ctx);
/* the final KIHF.iter (bin/KPopCount.ml:60) */
int kpc_finish(kpc_ctx *ctx);

/* number of k-mer windows counted so far (sum of all emitted counts); needs a device sync */
int kpc_kmers_counted(kpc_ctx *ctx, unsigned long long *out);

/* ---- multi-GPU (dense table, one context per GPU / per rank) ---------------------------------------------
 * Read-chunk sharding: every rank feeds its own records, then the 4^k tables are summed (NCCL all-reduce /
 * reduce over NVLink, done by the caller on the pointers below) and rank 0 calls kpc_finish.
 * kpc_dense_table returns KPC_E_STATE when the context is not on the dense path. */
int kpc_dense_table(kpc_ctx *ctx, void **lo_u32, void **hi_u64, unsigned long long *n_bins);
/* max over bins: if the sum of all ranks' maxima is < 2^32 the u32 table can be reduced as is */
int kpc_dense_max(kpc_ctx *ctx, unsigned long long *max_count);
/* otherwise: hi += lo, lo = 0 on every rank, reduce hi (u64) instead */
int kpc_dense_promote(kpc_ctx *ctx);

/* number of line feeds in n bytes of DEVICE memory (16-byte aligned).  Read-chunk sharding of one FASTQ file needs the
 * line index every shard starts at (records are four lines and cannot be recognised locally: Files.ml:201-221); the
 * shards' counts are the only data exchanged.  Synchronous. */
int kpc_count_newlines(kpc_ctx *ctx, const void *device_bytes, size_t n, unsigned long long *count);
/* 1 when some bin has been folded into the 64-bit side table (a reduction across ranks must then use kpc_dense_promote) */
int kpc_dense_has_hi(kpc_ctx *ctx, int *has_hi);

/* ---- multi-GPU, hash-table runs (large k): one sample cut into read chunks, one context per GPU / per rank -------------
 * Every rank counts its records with insertion ranks that are valid for the WHOLE stream (kpc_set_record_base: index of
 * the rank's first record), exports its distinct k-mers, the ranks exchange them by owner (a contiguous range of OCaml
 * bucket indices per rank), every owner merges what it receives -- counts add, the smallest insertion rank wins -- and
 * dumps its range: the concatenation in rank order is the spectrum (sparse merge).  Exact as long as the
 * merged table stays below -M (no dump before the end); kpop_b200/distributed.py checks that. */
int kpc_set_record_base(kpc_ctx *ctx, unsigned long long first_record);
/* device arrays (u64) of the n distinct k-mers counted so far: key, count, insertion rank; slots whose key is ~0 are empty
 * and must be skipped.  Valid until the next call on the context. */
int kpc_hash_export(kpc_ctx *ctx, void **keys, void **counts, void **ranks, unsigned long long *n_slots);
/* inserts n entries (device arrays, u64) into the table, after emptying it when clear_first != 0: equal keys merge */
int kpc_hash_import(kpc_ctx *ctx, const void *keys, const void *counts, const void *ranks, unsigned long long n, int clear_first);
/* B, the bucket count of OCaml's Hashtbl for this run (the dump is ordered by key mod B) */
unsigned long long kpc_bucket_count(const kpc_ctx *ctx);

/* empty the tables and forget every input, keeping all allocations (a context can then be used for another run) */
int kpc_reset(kpc_ctx *ctx);
/* the same, and the next run gets this spectrum label (bin/KPopCount.ml:158-172).  Batch use: one context serves many
 * samples, one KMerCounter.compute each -- what a `Parallel ... KPopCount -l <sample>` loop (README.md:579,1020) does with
 * one process per sample.  The label must stay non-empty (-l) or stay empty (-L): the table kind is fixed at creation. */
int kpc_reset_label(kpc_ctx *ctx, const char *label);
/* benchmarking with device-resident input: format the dump on the device but leave the text in HBM */
int kpc_discard_text(kpc_ctx *ctx, int discard);
unsigned long long kpc_text_bytes(const kpc_ctx *ctx); /* bytes of spectra text produced since create / reset */

/* ---- instrumentation ---------------------------------------------------------------------------------------- */
void *kpc_stream(kpc_ctx *ctx);          /* cudaStream_t the counting kernels are launched on */
int kpc_sync(kpc_ctx *ctx);
unsigned long long kpc_kernel_launches(const kpc_ctx *ctx); /* kernels of this library launched so far */
const char *kpc_backend(void);           /* "cuda" */
/* measurement aid (bench.py): CUDA-event times of the fast FASTQ kernels since the last read, summed over launches */
int kpc_profile_enable(kpc_ctx *ctx, int on);
int kpc_profile_read(kpc_ctx *ctx, double *partition_ms, double *count_ms, unsigned long long *launches,
                     unsigned long long *bytes);
/* synthetic single-end FASTQ of the benchmark shape written to device memory (see kpc_synth.h) */
int kpc_synth_fastq(kpc_ctx *ctx, void *device_out, unsigned long long first_record,
                    unsigned long long n_records, unsigned long long seed);
unsigned long long kpc_synth_offset(unsigned long long record);

#ifdef __cplusplus
}
#endif
#endif /* KPOPCOUNT_H */
