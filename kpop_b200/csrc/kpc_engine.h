// kpc_engine.h -- host side of libkpopcount_gpu: the state machine behind the C ABI (include/kpopcount.h).
//
// It restates KMerCounter.compute (bin/KPopCount.ml:26-63) around the device kernels:
//   * cuts the incoming byte stream of every input into launches (line aligned when possible), keeping back
//     the last lines of a FASTQ stream so that the "EOF inside a record drops it" rule of FASTQ.iter_se
//     (Files.ml:217) can be applied exactly,
//   * picks the table for the run:  DENSE  4^k / 32^k u32 bins when every key is its own OCaml bucket and the
//     table can never reach -M (default parameters, k <= 12);  HASH  open addressing + insertion ranks, with
//     the dump/clear rule of bin/KPopCount.ml:39 reproduced epoch by epoch;  TUPLE  for -L (one spectrum per
//     record): (record, key) tuples reduced by sort + run-length,
//   * orders and formats the dump on the device and hands the text to the sink.
#pragma once
#include <string>
#include <vector>

#include "../../include/kpopcount.h"
#include "kpc_kernels.h"

struct KpcEngineConfig {
  int k = 12;
  int content = KPC_DNA_DS;
  long long max_results_size = 16777216;
  std::string label;
  int device = 0;
};

class KpcEngine {
 public:
  explicit KpcEngine(const KpcEngineConfig &cfg);
  ~KpcEngine();

  void set_sink(kpc_sink_fn fn, void *user) { sink_ = fn; sink_user_ = user; discard_text_ = false; }
  void set_discard_text(bool d) { discard_text_ = d; }
  // the text goes straight into a caller-owned host buffer (device -> host copies land there, nothing in between)
  void set_sink_buffer(void *buf, size_t cap) { out_buf_ = (char *)buf; out_cap_ = cap; out_used_ = 0; if (buf) discard_text_ = false; }
  unsigned long long sink_buffer_used() const { return out_used_; }
  unsigned long long text_bytes() const { return text_bytes_; }
  void reset();
  void reset_label(const std::string &label);
  int staging_slots() const { return kStagingSlots; }
  void *staging(int slot, size_t *capacity);
  void begin(int format);
  void feed(int mate, const uint8_t *bytes, size_t n, bool eof);
  void feed_device(int mate, const uint8_t *dev, size_t n, bool eof);
  void set_pair_limit(long long n) { pair_limit_ = n; }
  void set_single_pass(bool on) { pe_single_pass_ = on; }
  long long complete_pairs() const { return complete_pairs_; }
  void end();
  void finish();
  unsigned long long kmers_counted();
  void dense_table(void **lo, void **hi, unsigned long long *nbins);
  unsigned long long dense_max();
  bool dense_has_hi() const { return dense_hi_ != nullptr; }
  // hash-table runs sharded over several contexts (kpop_b200/distributed.py: sparse merge)
  void set_record_base(unsigned long long first_record);
  void hash_export(void **keys, void **counts, void **ranks, unsigned long long *n_slots);
  void hash_import(const unsigned long long *keys, const unsigned long long *counts, const unsigned long long *ranks,
                   unsigned long long n, bool clear_first);
  unsigned long long bucket_count() const { return buckets_; }
  unsigned long long count_newlines_device(const uint8_t *dev, size_t n);
  void dense_promote();
  void *native_stream() { return rt_stream_native(compute_); }
  void sync();
  unsigned long long launches() const { return launches_; }
  void synth_fastq(void *dev_out, unsigned long long first, unsigned long long n, unsigned long long seed);

  // ---- shards of a FASTQ stream counted on several devices (kpc_multi.cpp); dense-table runs only ----
  struct Piece {
    const uint8_t *p;
    size_t n;
  };
  // copies the pieces (host memory) into the next ring buffer and counts its line feeds, all on the copy stream; returns
  // the ring slot.  ev (optional) is recorded on the copy stream behind the copies.
  int shard_upload(const Piece *pc, int npc, size_t len, rt_event ev1, rt_event ev2);
  unsigned long long shard_census(int slot);  // waits for the upload; number of line feeds in the shard
  // counts the shard: it starts a line, `lines_before` line feeds precede it in its stream; lines >= max_lines are ignored
  void shard_count(int slot, int mate, uint64_t lines_before, uint64_t max_lines);
  unsigned long long shard_err_line(int mate);  // first malformed line seen by this engine (~0 = none); synchronises
  void shard_begin(int format, bool with_header);
  void shard_end();
  size_t chunk_cap() const { return chunk_cap_; }
  void sync_copy() { rt_stream_sync(copy_); }
  // this engine's table += a 64-bit table that lives on another device (peer copy into scratch, then an add kernel)
  void dense_add_remote(const unsigned long long *remote_hi, unsigned long long nbins, int remote_device);
  rt_event new_event() { rt_set_device(cfg_.device); return rt_event_create(); }

  enum Mode { DENSE, HASH, TUPLE };
  Mode mode() const { return mode_; }
  int device() const { return cfg_.device; }

 private:
  static const int kStagingSlots = 3;
  static const int kRing = 3;

  struct StreamState {  // one per mate of the current input
    uint64_t fed = 0;   // bytes handed to the device so far == stream offset of the next launch
    uint8_t *hold[2] = {nullptr, nullptr};  // pinned; kept-back tail of the stream (ping-pong: one may be in flight)
    size_t hold_cap[2] = {0, 0};
    size_t hold_len = 0;
    int hold_cur = 0;
    rt_event hold_ev[2] = {nullptr, nullptr};
    bool hold_busy[2] = {false, false};
    KpcStreamCarry *carry[2] = {nullptr, nullptr};  // device; [cur] is the state at offset `fed`
    int cur = 0;
    unsigned long long *err_line = nullptr;  // device
    bool eof = false;
    bool any = false;
    uint8_t last_byte = '\n';
    bool tail_unsafe = false;
    bool at_line_start = true;  // the next launch begins at the first byte of a line
    uint64_t total_lines = 0;
    uint64_t records = 0;     // FASTQ: complete records (known at end of file); FASTA: header lines
    uint64_t final_recs = 0;  // -L: records of this mate that can be written out
  };
  struct RingSlot {
    uint8_t *buf = nullptr;
    rt_event computed = nullptr;
    bool used = false;
    rt_event censused = nullptr;  // shards: upload + line-feed census done
    size_t len = 0;
  };

  // --- paired-end input of hash-table / -L runs: the two mate streams are woven into one, pair by pair, on the host ---
  void pe_enqueue(int mate, const uint8_t *bytes, size_t n, bool eof);
  void pe_pump(bool finishing);
  bool pe_single_pass_ = false;      // weave paired files on the dense path too (inputs that cannot be read twice)
  bool pe_weave_ = false;            // the current input is a woven pair of files (the engine sees single-end records)
  std::vector<uint8_t> pe_q_[2];     // bytes of each mate not yet woven
  size_t pe_scan_[2] = {0, 0};       // how far each queue has been searched for line feeds
  size_t pe_rec_end_[2] = {0, 0};    // end of the first complete record in each queue (0 = none yet)
  int pe_lines_[2] = {0, 0};         // line feeds of the record being completed
  bool pe_eof_[2] = {false, false};
  bool pe_feeding_ = false;          // the weaver itself is calling feed()
  bool pe_done_ = false;             // one mate has ended: FASTQ.iter_pe stops there
  std::vector<uint8_t> pe_out_;
  uint64_t pe_pairs_ = 0;

  // --- stream cutting ---
  void reset_stream(StreamState &st);
  void hold_append(StreamState &st, const uint8_t *p, size_t n);
  size_t choose_cut(StreamState &st, const Piece *pc, int npc, size_t limit);
  void submit_host(StreamState &st, int mate, const Piece *pc, int npc, size_t len, bool final_launch, bool foreign);
  void run_launch(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch, bool halo_ok);
  bool fq_usable() const;
  void fq_ensure(size_t len);
  void fq_launch(StreamState &st, const uint8_t *dev, size_t len, uint64_t max_lines, bool halo_ok);
  uint64_t final_line_cap(StreamState &st, const uint8_t *dev, size_t len);
  RingSlot &next_slot();

  // --- kernels ---
  void launch_tiles(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch, uint64_t max_lines,
                    int sink_kind, const KpcHashSink *hs, const KpcTupleSink *ts, bool with_recs,
                    unsigned long long *probe_pos, uint64_t probe_from, const KpcBucketCountSink *bc = nullptr,
                    const KpcBucketScatterSink *bs = nullptr);
  void advance(StreamState &st, size_t len);
  void ensure_desc(uint64_t n_tiles);

  // --- mode specific ---
  void dense_after_launch(size_t len);
  void dense_fold_now();
  void dense_finish();
  void hash_init();
  void hash_process(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch, uint64_t max_lines);
  void hash_ensure_capacity(uint64_t need);
  void hash_dump(bool clear);
  void hash_finish();
  bool sort_usable(const StreamState &st, size_t len, bool final_launch) const;
  bool sort_process(StreamState &st, int mate, const uint8_t *dev, size_t len, uint64_t max_lines);
  void sort_migrate();
  void tuple_process(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch, uint64_t max_lines);
  void tuple_flush(bool input_done);
  void tuple_finish();
  KpcHashSink hash_sink(uint64_t lo, uint64_t hi, long long sign) const;

  // --- output ---
  void emit(const char *p, size_t n);
  void emit_entries(unsigned long long *keys, unsigned long long *counts, uint64_t n);
  void *scratch(size_t bytes);
  void *scratch2(size_t bytes);
  uint64_t grow_buckets(uint64_t size_reached);

  KpcEngineConfig cfg_;
  Mode mode_;
  int sbits_, hex_width_;
  uint64_t nbins_ = 0;
  uint64_t buckets_;  // current OCaml bucket count B (power of two, never shrinks)
  kpc_sink_fn sink_ = nullptr;
  void *sink_user_ = nullptr;
  bool discard_text_ = false;
  char *out_buf_ = nullptr;
  size_t out_cap_ = 0, out_used_ = 0;
  unsigned long long text_bytes_ = 0;
  bool header_done_ = false;
  bool failed_ = false;
  int format_ = -1;
  bool in_input_ = false;
  long long pair_limit_ = -1;
  long long complete_pairs_ = -1;
  uint64_t rank_base_ = 0;    // FASTA: bytes of all previous inputs; FASTQ: records of all previous inputs
  unsigned long long launches_ = 0;

  rt_stream compute_ = nullptr, copy_ = nullptr;
  size_t chunk_cap_;
  uint32_t tile_bytes_;
  RingSlot ring_[kRing];
  int ring_next_ = 0;
  uint8_t *staging_[kStagingSlots] = {nullptr, nullptr, nullptr};
  rt_event staging_ev_[kStagingSlots] = {nullptr, nullptr, nullptr};
  bool staging_busy_[kStagingSlots] = {false, false, false};
  StreamState streams_[2];

  KpcTileDesc *desc_ = nullptr;
  uint64_t desc_cap_ = 0;
  uint32_t *tile_counter_ = nullptr;
  uint32_t epoch_ = 0;
  unsigned long long *d_tmp_ = nullptr;  // 8 x u64 device scalars
  unsigned long long *d_per_rec_ = nullptr, *h_per_rec_ = nullptr;  // -L: [0] tuples kept, [1] entries final, [2..] entries per record
  size_t per_rec_cap_ = 0;
  unsigned long long *h_tmp_ = nullptr;  // pinned mirror
  void *scratch_ = nullptr;
  size_t scratch_cap_ = 0;
  void *scratch2_ = nullptr;
  size_t scratch2_cap_ = 0;
  char *h_out_ = nullptr;  // pinned bounce buffer for text
  size_t h_out_cap_ = 0;

  // dense
  uint32_t *dense_lo_ = nullptr;
  unsigned long long *dense_hi_ = nullptr;
  uint64_t dense_since_fold_ = 0;

  // fast FASTQ pipeline (kpc_fastq.h)
  bool fq_enabled_ = true;
  size_t fq_launch_bytes_ = 0, fq_alloc_len_ = 0, fq_state_tiles_ = 0;
  size_t fq_zero_bytes_ = 0, fq_cur_off_ = 0, fq_cap_off_ = 0, fq_base_off_ = 0;
  int fq_log_bins_ = 15;
  uint32_t fq_slices_ = 0;
  uint16_t *fq_queue_ = nullptr;
  uint8_t *fq_meta_ = nullptr;
  unsigned long long *fq_state_ = nullptr;

  // hash
  unsigned long long *hkeys_ = nullptr, *hcounts_ = nullptr, *hranks_ = nullptr;
  uint64_t hcap_ = 0;
  unsigned long long *d_hstat_ = nullptr;  // [0] distinct, [1] overflow
  uint64_t hdistinct_ = 0;
  uint64_t epoch_rank_lo_ = 0;  // windows with rank < this belong to already dumped epochs
  uint64_t max_lines_cap_ = ~0ull;            // shards: upper line limit of the launch being issued
  unsigned long long *d_shard_nl_ = nullptr;  // per ring slot: line feeds of the shard (device / pinned)
  unsigned long long *h_shard_nl_ = nullptr;
  KpcStreamCarry *h_shard_carry_ = nullptr;   // pinned, one per ring slot
  // sort path: a whole sample counted without the hash table (kpc_bucketsort.cuh); its entries wait here for finish()
  bool sort_enabled_ = true;
  bool sorted_pending_ = false;
  uint64_t sorted_slots_ = 0;     // entry slots (windows of the sample); slots with count 0 print nothing
  uint64_t sorted_distinct_ = 0;
  void *sort_buf_ = nullptr;
  size_t sort_buf_cap_ = 0;
  unsigned long long *skeys_ = nullptr, *scounts_ = nullptr, *sranks_ = nullptr;

  // tuple (-L)
  unsigned long long *tkeys_ = nullptr, *tranks_ = nullptr;
  uint32_t *trecs_ = nullptr;
  uint64_t tcap_ = 0;
  unsigned long long *d_tn_ = nullptr;
  uint64_t tn_ = 0;           // tuples currently buffered (all belong to records >= rec_done_)
  uint64_t rec_done_[2] = {0, 0};  // per mate: records already written out
  struct RecInfo {
    unsigned long long tag_start = ~0ull, tag_end = ~0ull;
    std::string tag;
    bool have_tag = false;
  };
  std::vector<RecInfo> open_recs_[2];  // records >= rec_done_ whose header has been seen
  uint64_t tuple_rec_base_ = 0;  // first record index of the name table of the current launch
  KpcRecEntry *d_recs_ = nullptr;
  uint64_t d_recs_cap_ = 0;
  std::vector<uint8_t> tag_bytes_;  // host copy of the bytes of the current launch (record names are cut out of it)
};
