// kpc_engine.cpp -- see kpc_engine.h.  Reference semantics cited inline (paths relative to the KPop tree).
#include "kpc_engine.h"
#include "kpc_fastq.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

uint64_t pow2_at_least(uint64_t lo, uint64_t n) {  // Hashtbl.create: power_2_above 16 n
  uint64_t s = lo;
  while (s < n) s <<= 1;
  return s;
}
size_t env_size(const char *name, size_t dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  return (size_t)strtoull(e, nullptr, 10);
}
// positions (offsets in the rope, ascending piece order) of the last `want` '\n' inside the first `limit` bytes
int last_newlines(const uint8_t *const *ptr, const size_t *len, int npc, size_t limit, int want, size_t *pos) {
  int found = 0;
  size_t starts[4];
  size_t off = 0;
  for (int i = 0; i < npc; ++i) { starts[i] = off; off += len[i]; }
  for (int i = npc - 1; i >= 0 && found < want; --i) {
    if (starts[i] >= limit) continue;
    size_t n = std::min(len[i], limit - starts[i]);
    while (n > 0 && found < want) {
      const void *q = memrchr(ptr[i], '\n', n);
      if (!q) break;
      size_t at = (size_t)((const uint8_t *)q - ptr[i]);
      pos[found++] = starts[i] + at;
      n = at;
    }
  }
  return found;
}
// Matrix.Base.strip_external_quotes_and_check (BiOCamLib/lib/Matrix.ml:83-99); false = Quotes_in_name
bool strip_quotes(const std::string &s0, std::string &out) {
  size_t l = s0.size();
  if (l == 0) { out.clear(); return true; }
  if (l == 1 && s0[0] == '"') return false;
  out = s0;
  if (out[0] == '"' && out[l - 1] == '"') out = out.substr(1, l - 2);
  return out.find('"') == std::string::npos;
}

}  // namespace

// =================================================================================================
// construction
// =================================================================================================
KpcEngine::KpcEngine(const KpcEngineConfig &cfg) : cfg_(cfg) {
  if (cfg.k < 1 || cfg.max_results_size < 1) throw KpcError(KPC_E_ARG, "k and max_results_size must be positive");
  if (cfg.content < KPC_DNA_SS || cfg.content > KPC_PROTEIN) throw KpcError(KPC_E_ARG, "unknown content");
  if (cfg.k > kpc_max_k(cfg.content))  // KMers.ml:264-267 / 145-148
    throw KpcError(KPC_E_K_RANGE, std::string("Invalid argument (k must be <= ") +
                                      std::to_string(kpc_max_k(cfg.content)) + ", found " + std::to_string(cfg.k) + ")");
  sbits_ = kpc_symbol_bits(cfg.content);
  hex_width_ = kpc_hex_width(cfg.content, cfg.k);
  buckets_ = pow2_at_least(16, (uint64_t)cfg.max_results_size);

  // table choice
  const int bits = sbits_ * cfg.k;
  if (cfg.label.empty()) {
    mode_ = TUPLE;  // -L, or -l "" (bin/KPopCount.ml:39,44)
  } else {
    bool dense = false;
    if (bits <= 24 && (1ull << bits) <= buckets_) {
      // largest number of distinct keys that can ever be in the table; the spill rule (size >= M) must be unreachable
      unsigned __int128 distinct;
      if (cfg.content == KPC_DNA_DS) {
        distinct = ((unsigned __int128)1 << (2 * cfg.k)) / 2;
        if (cfg.k % 2 == 0) distinct += ((unsigned __int128)1 << cfg.k) / 2;  // palindromes are their own reverse complement
      } else if (cfg.content == KPC_DNA_SS) {
        distinct = (unsigned __int128)1 << (2 * cfg.k);
      } else {
        distinct = 1;
        for (int i = 0; i < cfg.k; ++i) distinct *= 22;
      }
      dense = distinct < (unsigned __int128)cfg.max_results_size;
    }
    mode_ = dense ? DENSE : HASH;
  }

  rt_init(cfg.device);
  compute_ = rt_stream_create();
  copy_ = rt_stream_create();
  chunk_cap_ = env_size("KPC_CHUNK_BYTES", (size_t)64 << 20);
  if (chunk_cap_ < 64) chunk_cap_ = 64;
  tile_bytes_ = kpc_k_tile_bytes();
  {
    const char *e = getenv("KPC_FAST");
    fq_enabled_ = !(e && e[0] == '0');
    const char *e2 = getenv("KPC_SORT_PATH");
    sort_enabled_ = !(e2 && e2[0] == '0');
    fq_launch_bytes_ = env_size("KPC_FQ_LAUNCH_BYTES", (size_t)1 << 30);
    const size_t tb = kpc_fq_tile_bytes();
    fq_launch_bytes_ = std::max<size_t>(tb, fq_launch_bytes_ / tb * tb);
    fq_launch_bytes_ = std::min<size_t>(fq_launch_bytes_, (size_t)2 << 30);  // u32 queue cursors, u32 tile indices
  }
  tile_counter_ = (uint32_t *)rt_dmalloc(64);
  d_tmp_ = (unsigned long long *)rt_dmalloc(64 * sizeof(unsigned long long));
  h_tmp_ = (unsigned long long *)rt_hmalloc(64 * sizeof(unsigned long long));
  for (int m = 0; m < 2; ++m) {
    StreamState &st = streams_[m];
    for (int j = 0; j < 2; ++j) {
      st.carry[j] = (KpcStreamCarry *)rt_dmalloc(sizeof(KpcStreamCarry));
      st.hold_ev[j] = rt_event_create();
    }
    st.err_line = (unsigned long long *)rt_dmalloc(sizeof(unsigned long long));
  }
  for (int i = 0; i < kRing; ++i) { ring_[i].computed = rt_event_create(); ring_[i].censused = rt_event_create(); }
  d_shard_nl_ = (unsigned long long *)rt_dmalloc(kRing * sizeof(unsigned long long));
  h_shard_nl_ = (unsigned long long *)rt_hmalloc(kRing * sizeof(unsigned long long));
  h_shard_carry_ = (KpcStreamCarry *)rt_hmalloc(kRing * sizeof(KpcStreamCarry));
  for (int i = 0; i < kStagingSlots; ++i) staging_ev_[i] = rt_event_create();

  if (mode_ == DENSE) {
    nbins_ = 1ull << bits;
    dense_lo_ = (uint32_t *)rt_dmalloc(nbins_ * sizeof(uint32_t));
    rt_memset(dense_lo_, 0, nbins_ * sizeof(uint32_t), compute_);
  } else if (mode_ == HASH) {
    hash_init();
  } else {
    d_tn_ = (unsigned long long *)rt_dmalloc(sizeof(unsigned long long));
    rt_memset(d_tn_, 0, sizeof(unsigned long long), compute_);
  }
}

KpcEngine::~KpcEngine() {
  try {
    rt_stream_sync(compute_);
    rt_stream_sync(copy_);
  } catch (...) {
  }
  for (int m = 0; m < 2; ++m) {
    StreamState &st = streams_[m];
    for (int j = 0; j < 2; ++j) {
      rt_dfree(st.carry[j]);
      if (st.hold[j]) rt_hfree(st.hold[j]);
      rt_event_destroy(st.hold_ev[j]);
    }
    rt_dfree(st.err_line);
  }
  for (int i = 0; i < kRing; ++i) {
    if (ring_[i].buf) rt_dfree(ring_[i].buf);
    rt_event_destroy(ring_[i].computed);
    rt_event_destroy(ring_[i].censused);
  }
  for (int i = 0; i < kStagingSlots; ++i) {
    if (staging_[i]) rt_hfree(staging_[i]);
    rt_event_destroy(staging_ev_[i]);
  }
  rt_dfree(d_shard_nl_); rt_hfree(h_shard_nl_); rt_hfree(h_shard_carry_);
  rt_dfree(desc_); rt_dfree(tile_counter_); rt_dfree(d_tmp_); rt_hfree(h_tmp_);
  if (d_per_rec_) { rt_dfree(d_per_rec_); rt_hfree(h_per_rec_); }
  rt_dfree(scratch_); rt_dfree(scratch2_); rt_dfree(sort_buf_);
  if (h_out_) rt_hfree(h_out_);
  rt_dfree(dense_lo_); rt_dfree(dense_hi_);
  rt_dfree(fq_queue_); rt_dfree(fq_meta_); rt_dfree(fq_state_);
  rt_dfree(hkeys_); rt_dfree(hcounts_); rt_dfree(hranks_); rt_dfree(d_hstat_);
  rt_dfree(tkeys_); rt_dfree(tranks_); rt_dfree(trecs_); rt_dfree(d_tn_); rt_dfree(d_recs_);
  rt_stream_destroy(compute_);
  rt_stream_destroy(copy_);
}

void *KpcEngine::scratch(size_t bytes) {
  if (bytes > scratch_cap_) {
    rt_stream_sync(compute_);
    rt_dfree(scratch_);
    scratch_cap_ = bytes + bytes / 4 + 4096;
    scratch_ = rt_dmalloc(scratch_cap_);
  }
  return scratch_;
}
void *KpcEngine::scratch2(size_t bytes) {
  if (bytes > scratch2_cap_) {
    rt_stream_sync(compute_);
    rt_dfree(scratch2_);
    scratch2_cap_ = bytes + bytes / 4 + 4096;
    scratch2_ = rt_dmalloc(scratch2_cap_);
  }
  return scratch2_;
}

void KpcEngine::sync() {
  rt_stream_sync(copy_);
  rt_stream_sync(compute_);
}

void *KpcEngine::staging(int slot, size_t *capacity) {
  if (slot < 0 || slot >= kStagingSlots) throw KpcError(KPC_E_ARG, "staging slot out of range");
  if (!staging_[slot]) staging_[slot] = (uint8_t *)rt_hmalloc(chunk_cap_);
  if (staging_busy_[slot]) {
    rt_event_sync(staging_ev_[slot]);
    staging_busy_[slot] = false;
  }
  if (capacity) *capacity = chunk_cap_;
  return staging_[slot];
}

// =================================================================================================
// output
// =================================================================================================
void KpcEngine::emit(const char *p, size_t n) {
  if (!n || discard_text_) return;
  if (out_buf_) {
    if (out_used_ + n > out_cap_) throw KpcError(KPC_E_IO, "the sink buffer is too small for the spectra text");
    memcpy(out_buf_ + out_used_, p, n);
    out_used_ += n;
    return;
  }
  if (!sink_) throw KpcError(KPC_E_STATE, "no sink set (kpc_set_sink)");
  if (sink_(sink_user_, p, n) != 0) throw KpcError(KPC_E_IO, "the output sink reported a failure");
}

// device arrays (keys, counts) in final order -> text -> sink
void KpcEngine::emit_entries(unsigned long long *keys, unsigned long long *counts, uint64_t n) {
  if (!n) return;
  const size_t per = (size_t)hex_width_ + 22;
  const size_t batch = std::max<size_t>(1, ((size_t)256 << 20) / per);  // bound the device text buffer
  if (!h_out_) {
    h_out_cap_ = (size_t)16 << 20;
    h_out_ = (char *)rt_hmalloc(h_out_cap_);
  }
  for (uint64_t i0 = 0; i0 < n; i0 += batch) {
    const uint64_t m = std::min<uint64_t>(batch, n - i0);
    char *d_text = (char *)scratch2(m * per + kpc_k_scan_scratch_bytes(m) + 256);
    void *scan = d_text + ((m * per + 255) & ~(size_t)255);
    kpc_k_format(keys + i0, counts + i0, m, hex_width_, d_text, d_tmp_ + 8, scan, compute_);
    launches_ += 3;
    rt_d2h(h_tmp_ + 8, d_tmp_ + 8, sizeof(unsigned long long), compute_);
    rt_stream_sync(compute_);
    const size_t total = (size_t)h_tmp_[8];
    text_bytes_ += total;
    if (discard_text_) continue;  // device-resident benchmarking: the text stays in HBM
    if (out_buf_) {               // caller-owned host buffer: one copy, straight to its place
      if (out_used_ + total > out_cap_) throw KpcError(KPC_E_IO, "the sink buffer is too small for the spectra text");
      rt_d2h(out_buf_ + out_used_, d_text, total, compute_);
      rt_stream_sync(compute_);
      out_used_ += total;
      continue;
    }
    for (size_t o = 0; o < total; o += h_out_cap_) {
      const size_t c = std::min(h_out_cap_, total - o);
      rt_d2h(h_out_, d_text + o, c, compute_);
      rt_stream_sync(compute_);
      emit(h_out_, c);
    }
  }
}

// Hashtbl.add doubles the bucket array whenever size > 2 * buckets; the array never shrinks (clear keeps it)
uint64_t KpcEngine::grow_buckets(uint64_t size_reached) {
  while (size_reached > 2 * buckets_) buckets_ <<= 1;
  return buckets_;
}

// =================================================================================================
// input: begin / feed / end
// =================================================================================================
void KpcEngine::reset_stream(StreamState &st) {
  st.fed = 0; st.hold_len = 0; st.cur = 0; st.eof = false; st.any = false; st.last_byte = '\n';
  st.tail_unsafe = false; st.total_lines = 0; st.records = 0; st.final_recs = 0; st.at_line_start = true;
  KpcStreamCarry c;
  memset(&c, 0, sizeof c);
  c.s1 = kpc_s1_identity();
  c.kc = kpc_kc_identity();
  c.kc.closed = 1;
  c.last_byte = '\n';  // a virtual line feed before the stream makes its first byte a line start
  memcpy(h_tmp_ + 16, &c, sizeof c);
  rt_h2d(st.carry[0], h_tmp_ + 16, sizeof c, compute_);
  rt_h2d(st.carry[1], h_tmp_ + 16, sizeof c, compute_);
  rt_memset(st.err_line, 0xff, sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
}

// back to the state right after construction (tables emptied), keeping every allocation
void KpcEngine::reset() {
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_reset inside an input");
  rt_stream_sync(copy_);
  header_done_ = false; failed_ = false; rank_base_ = 0; pair_limit_ = -1; complete_pairs_ = -1;
  text_bytes_ = 0;
  out_used_ = 0;
  buckets_ = pow2_at_least(16, (uint64_t)cfg_.max_results_size);
  if (mode_ == DENSE) {
    rt_memset(dense_lo_, 0, nbins_ * sizeof(uint32_t), compute_);
    if (dense_hi_) rt_memset(dense_hi_, 0, nbins_ * sizeof(unsigned long long), compute_);
    dense_since_fold_ = 0;
  } else if (mode_ == HASH) {
    if (hcap_) { kpc_k_hash_clear(hkeys_, hcounts_, hranks_, hcap_, compute_); ++launches_; }
    rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
    hdistinct_ = 0; epoch_rank_lo_ = 0;
    sorted_pending_ = false; sorted_slots_ = 0; sorted_distinct_ = 0;
  } else {
    tn_ = 0;
    rt_memset(d_tn_, 0, sizeof(unsigned long long), compute_);
  }
}

void KpcEngine::reset_label(const std::string &label) {
  std::string clean;
  if (!strip_quotes(label, clean)) throw KpcError(KPC_E_QUOTES_IN_NAME, "Spectrum labels must not contain quotes");
  if (clean.empty() != cfg_.label.empty())
    throw KpcError(KPC_E_ARG, "kpc_reset_label cannot switch between -l and -L (the table kind is fixed at creation)");
  reset();
  cfg_.label = clean;
}

void KpcEngine::begin(int format) {
  if (failed_) throw KpcError(KPC_E_STATE, "context is in a failed state");
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_begin while another input is open");
  if (format != KPC_FASTA && format != KPC_FASTQ_SE && format != KPC_FASTQ_PE) throw KpcError(KPC_E_ARG, "unknown format");
  if (!header_done_) {  // bin/KPopCount.ml:33-34: the label goes out before any input is opened
    header_done_ = true;
    if (!cfg_.label.empty()) {
      std::string h = "\t" + cfg_.label + "\n";
      emit(h.data(), h.size());
    }
  }
  if (sorted_pending_) sort_migrate();  // a further input for the same spectrum: continue on the hash table
  // Paired-end files on the hash-table / -L paths: FASTQ.iter_pe hands out mate 1 and mate 2 of every pair in turn
  // (Files.ml:222-250, 363-368), and both the dump rule (bin/KPopCount.ml:39) and Hashtbl's order depend on that order.
  // The two byte streams are therefore woven into ONE single-end stream of complete pairs on the host (line splitting
  // with memchr: these are not the throughput paths) and everything downstream sees single-end records.
  // The dense table does not care about order: there the mates are counted independently and a shorter mate is found at
  // kpc_end (KPC_E_PE_MISMATCH, the caller runs again with a pair limit) -- unless the caller cannot read its inputs
  // twice (pipes) and has asked for one pass, which weaves as well.
  pe_weave_ = format == KPC_FASTQ_PE && (mode_ != DENSE || pe_single_pass_);
  if (pe_weave_) {
    format = KPC_FASTQ_SE;
    for (int m = 0; m < 2; ++m) { pe_q_[m].clear(); pe_scan_[m] = 0; pe_rec_end_[m] = 0; pe_lines_[m] = 0; pe_eof_[m] = false; }
    pe_done_ = false;
    pe_pairs_ = 0;
    pe_out_.clear();
  }
  format_ = format;
  in_input_ = true;
  reset_stream(streams_[0]);
  if (format == KPC_FASTQ_PE) reset_stream(streams_[1]);
  for (int m = 0; m < 2; ++m) { open_recs_[m].clear(); rec_done_[m] = 0; }
  if (mode_ == TUPLE) {  // tuples of records the previous input never completed (dropped pairs) must not leak
    tn_ = 0;
    rt_memset(d_tn_, 0, sizeof(unsigned long long), compute_);
    rt_stream_sync(compute_);
  }
}

void KpcEngine::hold_append(StreamState &st, const uint8_t *p, size_t n) {
  const int c = st.hold_cur;
  if (st.hold_len + n > st.hold_cap[c]) {
    size_t ncap = std::max<size_t>(st.hold_len + n, st.hold_cap[c] * 2);
    ncap = std::max<size_t>(ncap, std::min<size_t>(chunk_cap_ + 64, (size_t)1 << 20));
    uint8_t *nb = (uint8_t *)rt_hmalloc(ncap);
    if (st.hold_len) memcpy(nb, st.hold[c], st.hold_len);
    if (st.hold[c]) rt_hfree(st.hold[c]);
    st.hold[c] = nb;
    st.hold_cap[c] = ncap;
  }
  if (n) memcpy(st.hold[c] + st.hold_len, p, n);
  st.hold_len += n;
}

// Where to end a non-final launch inside the first `limit` bytes of the pending data.
//  FASTA: after the last line feed (a launch never ends on a '>' that opens a line, so the one-byte look-ahead
//         of the framing kernel stays inside the launch); no line feed at all: anywhere.
//  FASTQ: after the 5th line feed from the end, so that every line handed over is followed by at least four
//         more lines of the stream: its record is complete whatever comes next (Files.ml:204-217).
size_t KpcEngine::choose_cut(StreamState &st, const Piece *pc, int npc, size_t limit) {
  const uint8_t *ptr[4];
  size_t len[4];
  for (int i = 0; i < npc; ++i) { ptr[i] = pc[i].p; len[i] = pc[i].n; }
  size_t pos[5];
  if (format_ == KPC_FASTA) {
    int f = last_newlines(ptr, len, npc, limit, 1, pos);
    return f ? pos[0] + 1 : limit;
  }
  int f = last_newlines(ptr, len, npc, limit, 5, pos);
  if (f == 5) return pos[4] + 1;
  st.tail_unsafe = true;  // lines longer than a fifth of the staging size
  return f ? pos[0] + 1 : limit;
}

KpcEngine::RingSlot &KpcEngine::next_slot() {
  RingSlot &r = ring_[ring_next_];
  ring_next_ = (ring_next_ + 1) % kRing;
  if (!r.buf) r.buf = (uint8_t *)rt_dmalloc(chunk_cap_ + 256);
  return r;
}

void KpcEngine::feed(int mate, const uint8_t *bytes, size_t n, bool eof) {
  if (failed_) throw KpcError(KPC_E_STATE, "context is in a failed state");
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_feed outside kpc_begin / kpc_end");
  if (pe_weave_ && !pe_feeding_) { pe_enqueue(mate, bytes, n, eof); return; }
  if (mate < 0 || mate > 1 || (mate == 1 && format_ != KPC_FASTQ_PE)) throw KpcError(KPC_E_ARG, "bad mate index");
  StreamState &st = streams_[mate];
  if (st.eof) throw KpcError(KPC_E_STATE, "kpc_feed after eof");
  if (n) { st.any = true; st.last_byte = bytes[n - 1]; }

  int staging_slot = -1;
  for (int i = 0; i < kStagingSlots; ++i)
    if (staging_[i] && bytes >= staging_[i] && bytes < staging_[i] + chunk_cap_) staging_slot = i;
  bool issued_from_caller = false;

  const uint8_t *src = bytes;
  size_t rem = n;
  for (;;) {
    const size_t pending = st.hold_len + rem;
    if (pending < chunk_cap_ || (pending == chunk_cap_ && eof)) break;
    // carve one non-final launch out of the first chunk_cap_ bytes of [hold | src]
    Piece pc[2];
    int npc = 0;
    if (st.hold_len) pc[npc++] = Piece{st.hold[st.hold_cur], st.hold_len};
    if (rem) pc[npc++] = Piece{src, rem};
    size_t cut = choose_cut(st, pc, npc, chunk_cap_);
    if (cut <= st.hold_len) {
      // the whole launch comes out of the hold buffer (only with tiny staging sizes)
      Piece one{st.hold[st.hold_cur], cut};
      submit_host(st, mate, &one, 1, cut, false, false);
      rt_stream_sync(copy_);
      memmove(st.hold[st.hold_cur], st.hold[st.hold_cur] + cut, st.hold_len - cut);
      st.hold_len -= cut;
      st.hold_busy[st.hold_cur] = false;
      continue;
    }
    const size_t from_src = cut - st.hold_len;
    Piece two[2];
    int n2 = 0;
    if (st.hold_len) two[n2++] = Piece{st.hold[st.hold_cur], st.hold_len};
    two[n2++] = Piece{src, from_src};
    submit_host(st, mate, two, n2, cut, false, true);
    issued_from_caller = true;
    if (st.hold_len) {  // that hold buffer is in flight now: continue in the other one
      st.hold_cur ^= 1;
      st.hold_len = 0;
      if (st.hold_busy[st.hold_cur]) { rt_event_sync(st.hold_ev[st.hold_cur]); st.hold_busy[st.hold_cur] = false; }
    }
    src += from_src;
    rem -= from_src;
  }
  if (!eof) {
    if (rem) hold_append(st, src, rem);
  } else {
    st.eof = true;
    // End of file: everything that is left, plus a virtual line feed when the last line is unterminated
    // (input_line returns such a line like any other one).
    Piece pc[3];
    int npc = 0;
    size_t len = 0;
    if (st.hold_len) { pc[npc++] = Piece{st.hold[st.hold_cur], st.hold_len}; len += st.hold_len; }
    if (rem) { pc[npc++] = Piece{src, rem}; len += rem; issued_from_caller = true; }
    static const uint8_t nl = '\n';
    if (st.any && st.last_byte != '\n') { pc[npc++] = Piece{&nl, 1}; len += 1; }
    submit_host(st, mate, pc, npc, len, true, rem != 0);
    st.hold_len = 0;
  }
  if (issued_from_caller) {
    if (staging_slot >= 0) {
      rt_event_record(staging_ev_[staging_slot], copy_);
      staging_busy_[staging_slot] = true;
    } else {
      rt_stream_sync(copy_);  // the caller may reuse its buffer as soon as we return
    }
  }
}

// ---- paired-end weaving ---------------------------------------------------------------------------------------------
void KpcEngine::pe_enqueue(int mate, const uint8_t *bytes, size_t n, bool eof) {
  if (mate < 0 || mate > 1) throw KpcError(KPC_E_ARG, "bad mate index");
  if (pe_eof_[mate]) throw KpcError(KPC_E_STATE, "kpc_feed after eof");
  if (!pe_done_) pe_q_[mate].insert(pe_q_[mate].end(), bytes, bytes + n);
  if (eof) {
    pe_eof_[mate] = true;
    // input_line returns a last line without line feed like any other one
    if (!pe_done_ && !pe_q_[mate].empty() && pe_q_[mate].back() != '\n') pe_q_[mate].push_back('\n');
  }
  pe_pump(false);
}
// moves every complete PAIR of records (four lines from each mate) into the woven stream and feeds it on
void KpcEngine::pe_pump(bool finishing) {
  auto find_record = [&](int m) {  // end of the first complete record of queue m, 0 if there is none yet
    if (pe_rec_end_[m]) return pe_rec_end_[m];
    std::vector<uint8_t> &q = pe_q_[m];
    while (pe_scan_[m] < q.size()) {
      const void *p = memchr(q.data() + pe_scan_[m], '\n', q.size() - pe_scan_[m]);
      if (!p) { pe_scan_[m] = q.size(); break; }
      pe_scan_[m] = (size_t)((const uint8_t *)p - q.data()) + 1;
      if (++pe_lines_[m] == 4) { pe_lines_[m] = 0; pe_rec_end_[m] = pe_scan_[m]; break; }
    }
    return pe_rec_end_[m];
  };
  size_t used[2] = {0, 0};
  while (!pe_done_) {
    // relative to what has been consumed in this call
    size_t e0 = find_record(0), e1 = find_record(1);
    if (!e0 || !e1) break;
    pe_out_.insert(pe_out_.end(), pe_q_[0].begin() + used[0], pe_q_[0].begin() + e0);
    pe_out_.insert(pe_out_.end(), pe_q_[1].begin() + used[1], pe_q_[1].begin() + e1);
    used[0] = e0; used[1] = e1;
    pe_rec_end_[0] = pe_rec_end_[1] = 0;
    ++pe_pairs_;
    if (pe_out_.size() >= chunk_cap_ / 2) {
      pe_feeding_ = true;
      try { feed(0, pe_out_.data(), pe_out_.size(), false); } catch (...) { pe_feeding_ = false; throw; }
      pe_feeding_ = false;
      pe_out_.clear();
    }
  }
  for (int m = 0; m < 2; ++m) {
    if (used[m]) {
      pe_q_[m].erase(pe_q_[m].begin(), pe_q_[m].begin() + used[m]);
      pe_scan_[m] -= used[m];
      if (pe_rec_end_[m]) pe_rec_end_[m] -= used[m];  // a record of this mate that is still waiting for its partner
    }
  }
  // a mate that has ended with no complete record left: iteration stops (a record of the other mate that was already
  // read is dropped with it, Files.ml:228-247)
  for (int m = 0; m < 2; ++m)
    if (pe_eof_[m] && !find_record(m)) pe_done_ = true;
  if (pe_done_) { pe_q_[0].clear(); pe_q_[1].clear(); pe_scan_[0] = pe_scan_[1] = 0; }
  if ((pe_done_ || finishing) && !streams_[0].eof && (pe_eof_[0] && pe_eof_[1])) {
    pe_feeding_ = true;
    try { feed(0, pe_out_.data(), pe_out_.size(), true); } catch (...) { pe_feeding_ = false; throw; }
    pe_feeding_ = false;
    pe_out_.clear();
  }
}

// copy the pieces into the next ring buffer and run the launch
void KpcEngine::submit_host(StreamState &st, int mate, const Piece *pc, int npc, size_t len, bool final_launch,
                            bool /*foreign*/) {
  if (len > chunk_cap_ + 1) throw KpcError(KPC_E_STATE, "internal: launch larger than the staging size");
  RingSlot &slot = next_slot();
  if (slot.used) rt_stream_wait(copy_, slot.computed);  // the kernels that read this buffer must be done
  size_t off = 0;
  for (int i = 0; i < npc; ++i) {
    if (!pc[i].n) continue;
    rt_h2d(slot.buf + off, pc[i].p, pc[i].n, copy_);
    off += pc[i].n;
    for (int j = 0; j < 2; ++j)
      if (pc[i].p == st.hold[j]) { rt_event_record(st.hold_ev[j], copy_); st.hold_busy[j] = true; }
  }
  rt_memset(slot.buf + off, 0, 16, copy_);  // the kernels read whole 16-byte vectors: the bytes past the end are masked, not undefined
  rt_event ev = rt_event_create();
  rt_event_record(ev, copy_);
  rt_stream_wait(compute_, ev);
  rt_event_destroy(ev);
  // -L needs the record names: they are cut out of the host copy of the launch after the kernel has run
  if (mode_ == TUPLE) {
    tag_bytes_.resize(len);
    size_t o = 0;
    for (int i = 0; i < npc; ++i) { if (pc[i].n) memcpy(tag_bytes_.data() + o, pc[i].p, pc[i].n); o += pc[i].n; }
  }
  run_launch(st, mate, slot.buf, len, final_launch, false);
  for (int i = npc - 1; i >= 0; --i)
    if (pc[i].n) { st.at_line_start = pc[i].p[pc[i].n - 1] == '\n'; break; }
  rt_event_record(slot.computed, compute_);
  slot.used = true;
}

void KpcEngine::feed_device(int mate, const uint8_t *dev, size_t n, bool eof) {
  if (failed_) throw KpcError(KPC_E_STATE, "context is in a failed state");
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_feed_device outside kpc_begin / kpc_end");
  if (!eof) throw KpcError(KPC_E_UNSUPPORTED, "kpc_feed_device takes a whole input in one call (eof must be set)");
  if (mate < 0 || mate > 1 || (mate == 1 && format_ != KPC_FASTQ_PE)) throw KpcError(KPC_E_ARG, "bad mate index");
  if (((uintptr_t)dev & 15) != 0) throw KpcError(KPC_E_ARG, "device input must be 16-byte aligned");
  StreamState &st = streams_[mate];
  if (st.eof || st.fed || st.hold_len) throw KpcError(KPC_E_UNSUPPORTED, "kpc_feed_device cannot be mixed with kpc_feed");
  if (mode_ != DENSE) {
    // hash-table / -L runs: one launch, through a ring buffer (room for the line feed a last unterminated line gets)
    if (mode_ == TUPLE) throw KpcError(KPC_E_UNSUPPORTED, "kpc_feed_device is not available with -L (record names are cut on the host)");
    if (n > chunk_cap_) throw KpcError(KPC_E_UNSUPPORTED, "device inputs on the hash-table path are limited to the staging size");
    st.eof = true;
    RingSlot &slot = next_slot();
    if (slot.used) rt_stream_wait(compute_, slot.computed);
    size_t tl = n;
    if (n) {
      st.any = true;
      rt_d2d(slot.buf, dev, n, compute_);
      rt_d2h(h_tmp_ + 24, dev + (n - 1), 1, compute_);
      rt_stream_sync(compute_);
      st.last_byte = *(const uint8_t *)(h_tmp_ + 24);
      if (st.last_byte != '\n') {
        *(uint8_t *)(h_tmp_ + 24) = '\n';
        rt_h2d(slot.buf + tl, h_tmp_ + 24, 1, compute_);
        tl += 1;
      }
    }
    run_launch(st, mate, slot.buf, tl, true, false);
    rt_event_record(slot.computed, compute_);
    slot.used = true;
    return;
  }
  st.eof = true;
  if (!n) { run_launch(st, mate, dev, 0, true, false); return; }
  st.any = true;
  // find the cut between the in-place body and the tail (last lines) on a host copy of the end of the input
  size_t win = std::min<size_t>(n, std::min<size_t>(chunk_cap_, (size_t)1 << 20));
  std::vector<uint8_t> tail(win);
  rt_d2h(tail.data(), dev + (n - win), win, compute_);
  rt_stream_sync(compute_);
  st.last_byte = tail[win - 1];
  Piece pc{tail.data(), win};
  size_t cut_in_win = (win == n && n <= chunk_cap_) ? 0 : choose_cut(st, &pc, 1, win);
  size_t cut = (n - win) + cut_in_win;
  cut &= ~(size_t)15;  // the tail launch starts at a fresh aligned buffer; the body must end on a 16-byte boundary
  if (n - cut > chunk_cap_) throw KpcError(KPC_E_UNSUPPORTED, "lines longer than the staging size in a device input");
  if (cut) {
    // the body may now end in the middle of a line: legal for FASTQ (line-exact hold-back is what matters and the
    // tail still holds the last five line feeds); for FASTA make sure it does not end on a line-opening '>'
    if (format_ == KPC_FASTA && cut >= (n - win) + 1 && tail[cut - (n - win) - 1] == '>') {
      if (cut >= 16) cut -= 16; else cut = 0;
    }
  }
  if (cut) {
    if (fq_usable()) {
      // the fast pipeline works on bounded launches (its queues are sized for one): consecutive pieces of the same
      // device buffer, so every piece but the first can read the 16 bytes before it
      const size_t sub = fq_launch_bytes_;
      for (size_t off = 0; off < cut; off += sub) run_launch(st, mate, dev + off, std::min(sub, cut - off), false, off > 0);
    } else {
      // the generic kernel: pieces of at most 1 GiB as well, so that the u32 counters can be folded in between; a piece of
      // a FASTA stream must not end on a '>' that opens a line (the framing kernel looks one byte ahead there)
      const size_t sub = (size_t)1 << 30;
      size_t off = 0;
      while (off < cut) {
        size_t end = std::min(cut, off + sub);
        while (format_ == KPC_FASTA && end < cut && end > off + 16) {
          rt_d2h(h_tmp_ + 25, dev + end - 1, 1, compute_);
          rt_stream_sync(compute_);
          if (*(const uint8_t *)(h_tmp_ + 25) != '>') break;
          end -= 16;
        }
        run_launch(st, mate, dev + off, end - off, false, false);
        off = end;
      }
    }
    const size_t w0 = n - win;
    st.at_line_start = cut > w0 ? tail[cut - w0 - 1] == '\n' : false;
  }
  RingSlot &slot = next_slot();
  if (slot.used) rt_stream_wait(compute_, slot.computed);
  size_t tl = n - cut;
  rt_d2d(slot.buf, dev + cut, tl, compute_);
  if (st.last_byte != '\n') {
    h_tmp_[24] = '\n';
    rt_h2d(slot.buf + tl, h_tmp_ + 24, 1, compute_);
    tl += 1;
  }
  run_launch(st, mate, slot.buf, tl, true, false);
  rt_event_record(slot.computed, compute_);
  slot.used = true;
}

// =================================================================================================
// shards (kpc_multi.cpp): the stream of an input is cut at line starts by the caller, every shard goes to one device
// =================================================================================================
void KpcEngine::shard_begin(int format, bool with_header) {
  const bool saved = discard_text_;
  if (!with_header) discard_text_ = true;  // the label header is written once, by the first engine
  begin(format);
  discard_text_ = saved;
}
int KpcEngine::shard_upload(const Piece *pc, int npc, size_t len, rt_event ev1, rt_event ev2) {
  if (mode_ != DENSE) throw KpcError(KPC_E_STATE, "internal: shards are a dense-table facility");
  if (len > chunk_cap_ + 1) throw KpcError(KPC_E_STATE, "internal: shard larger than the staging size");
  const int id = ring_next_;
  RingSlot &slot = next_slot();
  // the launch that used this buffer (and its pinned carry / census words) must be done: the host waits here, which also
  // bounds how far it can run ahead of the device
  if (slot.used) rt_event_sync(slot.computed);
  size_t off = 0;
  for (int i = 0; i < npc; ++i) {
    if (!pc[i].n) continue;
    rt_h2d(slot.buf + off, pc[i].p, pc[i].n, copy_);
    off += pc[i].n;
  }
  if (ev1) rt_event_record(ev1, copy_);
  if (ev2) rt_event_record(ev2, copy_);
  rt_memset(d_shard_nl_ + id, 0, sizeof(unsigned long long), copy_);
  kpc_k_count_newlines(slot.buf, len, d_shard_nl_ + id, copy_);
  launches_ += len ? 1 : 0;
  rt_d2h(h_shard_nl_ + id, d_shard_nl_ + id, sizeof(unsigned long long), copy_);
  rt_event_record(slot.censused, copy_);
  slot.len = len;
  slot.used = true;
  return id;
}
unsigned long long KpcEngine::shard_census(int id) {
  rt_event_sync(ring_[id].censused);
  return h_shard_nl_[id];
}
void KpcEngine::shard_count(int id, int mate, uint64_t lines_before, uint64_t max_lines) {
  StreamState &st = streams_[mate];
  RingSlot &slot = ring_[id];
  KpcStreamCarry c;
  memset(&c, 0, sizeof c);
  c.s1 = kpc_s1_identity();
  c.s1.count = lines_before;
  c.kc = kpc_kc_identity();
  c.kc.closed = 1;
  c.last_byte = '\n';
  h_shard_carry_[id] = c;  // (the slot's previous launch has been waited for by shard_upload)
  rt_stream_wait(compute_, slot.censused);
  rt_h2d(st.carry[st.cur], h_shard_carry_ + id, sizeof c, compute_);
  st.at_line_start = true;
  st.any = true;
  max_lines_cap_ = max_lines;
  run_launch(st, mate, slot.buf, slot.len, false, false);
  max_lines_cap_ = ~0ull;
  rt_event_record(slot.computed, compute_);
}
unsigned long long KpcEngine::shard_err_line(int mate) {
  rt_d2h(h_tmp_ + mate, streams_[mate].err_line, sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
  return h_tmp_[mate];
}
void KpcEngine::dense_add_remote(const unsigned long long *remote_hi, unsigned long long nbins, int remote_device) {
  if (mode_ != DENSE || nbins != nbins_) throw KpcError(KPC_E_STATE, "internal: table shapes differ between devices");
  dense_promote();  // everything of this engine in the 64-bit table as well
  unsigned long long *tmp = (unsigned long long *)scratch(nbins_ * sizeof(unsigned long long));
  rt_peer_copy(tmp, cfg_.device, remote_hi, remote_device, nbins_ * sizeof(unsigned long long), compute_);
  kpc_k_add_u64(dense_hi_, tmp, nbins_, compute_);
  ++launches_;
  rt_stream_sync(compute_);
}
void KpcEngine::shard_end() {
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_end without kpc_begin");
  in_input_ = false;
  rt_stream_sync(copy_);
  rt_stream_sync(compute_);
}

void KpcEngine::ensure_desc(uint64_t n_tiles) {
  if (n_tiles <= desc_cap_) return;
  rt_stream_sync(compute_);
  rt_dfree(desc_);
  desc_cap_ = n_tiles + n_tiles / 8 + 64;
  desc_ = (KpcTileDesc *)rt_dmalloc(desc_cap_ * sizeof(KpcTileDesc));
  rt_memset(desc_, 0, desc_cap_ * sizeof(KpcTileDesc), compute_);
  // flags carry the launch epoch: restart it so that stale values of a previous array cannot match
  epoch_ = 0;
}

// number of lines the final launch may use: complete records only (FASTQ.iter_se drops a record cut by EOF)
uint64_t KpcEngine::final_line_cap(StreamState &st, const uint8_t *dev, size_t len) {
  rt_memset(d_tmp_, 0, sizeof(unsigned long long), compute_);
  kpc_k_count_newlines(dev, len, d_tmp_, compute_);
  launches_ += len ? 1 : 0;
  rt_d2h(h_tmp_, d_tmp_, sizeof(unsigned long long), compute_);
  rt_d2h(h_tmp_ + 32, st.carry[st.cur], sizeof(KpcStreamCarry), compute_);
  rt_stream_sync(compute_);
  KpcStreamCarry c;
  memcpy(&c, h_tmp_ + 32, sizeof c);
  st.total_lines = c.s1.count + h_tmp_[0];
  uint64_t recs = st.total_lines / 4;
  if (st.tail_unsafe && (st.total_lines % 4) != 0)
    throw KpcError(KPC_E_UNSUPPORTED,
                   "truncated FASTQ whose last lines are longer than the staging size (raise KPC_CHUNK_BYTES)");
  st.records = recs;
  return recs * 4;
}

// FASTQ + DNA + dense table + 8 <= k <= 12: the partition / shared-memory-count pipeline of kpc_fastq.cu
bool KpcEngine::fq_usable() const {
  return fq_enabled_ && mode_ == DENSE && format_ != KPC_FASTA && kpc_fq_supported(cfg_.k, cfg_.content);
}

void KpcEngine::fq_ensure(size_t len) {
  if (len <= fq_alloc_len_) return;
  rt_stream_sync(compute_);
  rt_dfree(fq_queue_); rt_dfree(fq_meta_); rt_dfree(fq_state_);
  fq_queue_ = nullptr; fq_meta_ = nullptr; fq_state_ = nullptr;
  fq_alloc_len_ = std::max<size_t>(len, std::min<size_t>(fq_launch_bytes_, (size_t)8 << 20));
  fq_log_bins_ = kpc_fq_log_bins(cfg_.k);
  fq_slices_ = (uint32_t)(nbins_ >> fq_log_bins_);
  // every byte yields at most one k-mer (FASTQ of 150-base reads has ~0.45) and slices are cut out of the middle
  // of the key, where canonical k-mers of uniform reads are uniform: capacities are 1.25 k-mers per byte spread
  // evenly, plus the padding the CTAs leave behind; anything beyond is counted in place by the kernel.
  // (KPC_FQ_QUEUE_PERMILLE / KPC_FQ_QUEUE_SLACK: test knobs that make the queues overflow on small inputs)
  const double fq_queue_factor = (double)env_size("KPC_FQ_QUEUE_PERMILLE", 1250) / 1000.0;
  const unsigned long long fq_queue_slack = env_size("KPC_FQ_QUEUE_SLACK", kpc_fq_queue_slack());
  std::vector<unsigned long long> base(fq_slices_);
  std::vector<uint32_t> cap(fq_slices_);
  unsigned long long total = 0;
  for (uint32_t b = 0; b < fq_slices_; ++b) {
    unsigned long long c = (unsigned long long)((double)fq_alloc_len_ * fq_queue_factor / fq_slices_) + fq_queue_slack;
    c = (c + 15) & ~15ull;
    if (c > 0xfffffff0ull) c = 0xfffffff0ull;
    base[b] = total;
    cap[b] = (uint32_t)c;
    total += c;
  }
  fq_queue_ = (uint16_t *)rt_dmalloc(total * sizeof(uint16_t) + 64);
  // meta: [0, 64) claim counters | cursors | capacities | bases
  const size_t cur_off = 64, cap_off = cur_off + 4 * (size_t)fq_slices_, base_off = (cap_off + 4 * (size_t)fq_slices_ + 7) & ~(size_t)7;
  fq_meta_ = (uint8_t *)rt_dmalloc(base_off + 8 * (size_t)fq_slices_);
  fq_zero_bytes_ = cap_off;
  fq_cur_off_ = cur_off; fq_cap_off_ = cap_off; fq_base_off_ = base_off;
  rt_h2d(fq_meta_ + cap_off, cap.data(), 4 * (size_t)fq_slices_, compute_);
  rt_h2d(fq_meta_ + base_off, base.data(), 8 * (size_t)fq_slices_, compute_);
  rt_stream_sync(compute_);  // the vectors go out of scope
  fq_state_tiles_ = (fq_alloc_len_ + kpc_fq_tile_bytes() - 1) / kpc_fq_tile_bytes() + 1;
  fq_state_ = (unsigned long long *)rt_dmalloc(fq_state_tiles_ * 8);
}

void KpcEngine::fq_launch(StreamState &st, const uint8_t *dev, size_t len, uint64_t max_lines, bool halo_ok) {
  fq_ensure(len);
  KpcFqLaunch L;
  memset(&L, 0, sizeof L);
  L.data = dev;
  L.n = len;
  L.abs_base = st.fed;
  L.halo_ok = halo_ok ? 1 : 0;
  L.max_lines = max_lines;
  L.k = cfg_.k;
  L.content = cfg_.content;
  L.carry_in = st.carry[st.cur];
  L.carry_out = st.carry[st.cur ^ 1];
  L.err_line = st.err_line;
  L.tile_state = fq_state_;
  L.counters = (uint32_t *)fq_meta_;
  L.n_tiles = (uint32_t)((len + kpc_fq_tile_bytes() - 1) / kpc_fq_tile_bytes());
  L.log_bins = fq_log_bins_;
  L.lo_bits = kpc_fq_lo_bits(cfg_.k);
  L.slice_bits = 2 * cfg_.k - fq_log_bins_;
  L.n_slices = fq_slices_;
  L.queue = fq_queue_;
  L.qbase = (const unsigned long long *)(fq_meta_ + fq_base_off_);
  L.qcap = (const uint32_t *)(fq_meta_ + fq_cap_off_);
  L.qcursor = (uint32_t *)(fq_meta_ + fq_cur_off_);
  L.table = dense_lo_;
  rt_memset(fq_state_, 0, (size_t)L.n_tiles * 8, compute_);
  rt_memset(fq_meta_, 0, fq_zero_bytes_, compute_);
  kpc_fq_partition(L, compute_);
  kpc_fq_count(L, compute_);
  launches_ += 2;
}

void KpcEngine::run_launch(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch, bool halo_ok) {
  uint64_t max_lines = ~0ull;
  if (format_ != KPC_FASTA) {
    if (final_launch) max_lines = final_line_cap(st, dev, len);
    if (pair_limit_ >= 0) max_lines = std::min<uint64_t>(max_lines, (uint64_t)pair_limit_ * 4);
    max_lines = std::min<uint64_t>(max_lines, max_lines_cap_);
  }
  switch (mode_) {
    case DENSE: {
      // every byte yields at most one window: fold the u32 counters BEFORE a launch that could carry one past 2^32
      if (dense_since_fold_ + len >= (1ull << 31)) dense_fold_now();
      if (len && fq_usable() && (halo_ok || st.at_line_start) && len <= fq_launch_bytes_)
        fq_launch(st, dev, len, max_lines, halo_ok);
      else
        launch_tiles(st, mate, dev, len, final_launch, max_lines, KPC_SINK_DENSE, nullptr, nullptr, false, nullptr, 0);
      advance(st, len);
      dense_after_launch(len);
      break;
    }
    case HASH:
      if (!(sort_usable(st, len, final_launch) && sort_process(st, mate, dev, len, max_lines)))
        hash_process(st, mate, dev, len, final_launch, max_lines);
      break;
    case TUPLE: tuple_process(st, mate, dev, len, final_launch, max_lines); break;
  }
}

void KpcEngine::launch_tiles(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch,
                             uint64_t max_lines, int sink_kind, const KpcHashSink *hs, const KpcTupleSink *ts,
                             bool with_recs, unsigned long long *probe_pos, uint64_t probe_from,
                             const KpcBucketCountSink *bc, const KpcBucketScatterSink *bs) {
  const uint64_t n_tiles = (len + tile_bytes_ - 1) / tile_bytes_;
  if (n_tiles == 0) {
    // nothing to scan: the state simply carries over
    rt_d2d(st.carry[st.cur ^ 1], st.carry[st.cur], sizeof(KpcStreamCarry), compute_);
    return;
  }
  if (n_tiles >= 0xffffffffull) throw KpcError(KPC_E_UNSUPPORTED, "launch too large");
  ensure_desc(n_tiles);
  ++epoch_;
  if (epoch_ >= (1u << 30)) {  // flags hold epoch << 2
    rt_memset(desc_, 0, desc_cap_ * sizeof(KpcTileDesc), compute_);
    epoch_ = 1;
  }
  rt_memset(tile_counter_, 0, sizeof(uint32_t), compute_);
  KpcTileLaunch L;
  memset(&L, 0, sizeof L);
  L.fmt = format_ == KPC_FASTA ? KPC_FMT_FASTA : KPC_FMT_FASTQ;
  L.content = cfg_.content;
  L.sink = sink_kind;
  if (getenv("KPC_DEBUG_NULL_SINK")) L.sink = KPC_SINK_NULL;  // profiling knob: framing + k-mers, no table traffic
  L.p.data = dev;
  L.p.n = len;
  L.p.abs_base = st.fed;
  L.p.rank_base = rank_base_;
  L.p.rank_mates = format_ == KPC_FASTQ_PE ? 2 : 1;
  L.p.rank_mate = (uint32_t)mate;
  L.p.max_lines = max_lines;
  L.p.k = cfg_.k;
  L.p.epoch = epoch_;
  L.p.n_tiles = (uint32_t)n_tiles;
  L.p.final_launch = final_launch ? 1 : 0;
  L.p.desc = desc_;
  L.p.carry_in = st.carry[st.cur];
  L.p.carry_out = st.carry[st.cur ^ 1];
  L.p.tile_counter = tile_counter_;
  L.p.err_line = st.err_line;
  L.p.rec_tab = with_recs ? d_recs_ : nullptr;
  L.p.probe_pos = probe_pos;
  L.p.probe_from = probe_from;
  if (sink_kind == KPC_SINK_DENSE) L.dense.table = dense_lo_;
  if (hs) L.hash = *hs;
  if (ts) L.tuple = *ts;
  if (bc) L.bcount = *bc;
  if (bs) L.bscatter = *bs;
  L.p.rec_base = tuple_rec_base_;
  L.p.rec_cap = d_recs_cap_;
  kpc_k_tiles(L, compute_);
  ++launches_;
}

void KpcEngine::advance(StreamState &st, size_t len) {
  st.fed += len;
  st.cur ^= 1;
}

void KpcEngine::end() {
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_end without kpc_begin");
  if (pe_weave_) {
    if (!pe_eof_[0] || !pe_eof_[1]) throw KpcError(KPC_E_STATE, "kpc_end before eof was signalled on every mate");
    pe_pump(true);
    complete_pairs_ = (long long)pe_pairs_;
  }
  const int mates = format_ == KPC_FASTQ_PE ? 2 : 1;
  for (int m = 0; m < mates; ++m)
    if (!streams_[m].eof) throw KpcError(KPC_E_STATE, "kpc_end before eof was signalled on every mate");
  in_input_ = false;
  rt_stream_sync(copy_);
  if (format_ != KPC_FASTA) {
    // Malformed records (Files.ml:213-214, 241-243): only complete records / pairs are ever checked
    for (int m = 0; m < mates; ++m) rt_d2h(h_tmp_ + m, streams_[m].err_line, sizeof(unsigned long long), compute_);
    rt_stream_sync(compute_);
    uint64_t recs = streams_[0].records;
    if (mates == 2) {
      const uint64_t r0 = streams_[0].records, r1 = streams_[1].records;
      recs = std::min(r0, r1);
      if (pair_limit_ >= 0) recs = std::min<uint64_t>(recs, (uint64_t)pair_limit_);
      complete_pairs_ = (long long)recs;
      const uint64_t used0 = pair_limit_ >= 0 ? std::min<uint64_t>(r0, (uint64_t)pair_limit_) : r0;
      const uint64_t used1 = pair_limit_ >= 0 ? std::min<uint64_t>(r1, (uint64_t)pair_limit_) : r1;
      if (used0 != used1 && mode_ != TUPLE) {
        failed_ = true;
        throw KpcError(KPC_E_PE_MISMATCH, "paired FASTQ files hold different numbers of records (" +
                                              std::to_string(r0) + " and " + std::to_string(r1) + ")");
      }
    }
    uint64_t bad = ~0ull;
    for (int m = 0; m < mates; ++m)
      if (h_tmp_[m] != ~0ull && h_tmp_[m] / 4 < recs) bad = std::min<uint64_t>(bad, h_tmp_[m] / 4);
    if (mode_ == TUPLE) tuple_flush(true);
    if (bad != ~0ull) {
      failed_ = true;
      // woven pairs: record 2i and 2i + 1 are the mates of pair i; the reference names the last line of the pair
      const uint64_t line = pe_weave_ ? (bad / 2 + 1) * 8 : (bad + 1) * 4 * mates;
      throw KpcError(KPC_E_MALFORMED_FASTQ, "On line " + std::to_string(line) + ": Malformed FASTQ file");
    }
    rank_base_ += recs * mates;
  } else {
    if (mode_ == TUPLE) tuple_flush(true);
    rank_base_ += streams_[0].fed;
  }
}

void KpcEngine::finish() {
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_finish inside an input");
  if (failed_) throw KpcError(KPC_E_STATE, "context is in a failed state");
  if (!header_done_) return;  // no input at all: the reference prints nothing (bin/KPopCount.ml:218)
  switch (mode_) {
    case DENSE: dense_finish(); break;
    case HASH: hash_finish(); break;
    case TUPLE: tuple_finish(); break;
  }
}

// =================================================================================================
// DENSE
// =================================================================================================
// bins at or above 2^31 move into the 64-bit side table.  Between two folds at most 2^31 windows are counted (checked
// before every launch, launches are at most 1 GiB), so a bin that stays in the u32 table is below 2^31 + 2^31: no wrap.
void KpcEngine::dense_fold_now() {
  if (!dense_since_fold_) return;
  if (!dense_hi_) {
    dense_hi_ = (unsigned long long *)rt_dmalloc(nbins_ * sizeof(unsigned long long));
    rt_memset(dense_hi_, 0, nbins_ * sizeof(unsigned long long), compute_);
  }
  kpc_k_dense_fold(dense_lo_, dense_hi_, nbins_, compute_);
  ++launches_;
  dense_since_fold_ = 0;
}
void KpcEngine::dense_after_launch(size_t len) { dense_since_fold_ += len; }

void KpcEngine::dense_finish() {
  // every key is its own bucket (4^k <= B): Hashtbl.iter order is ascending key order
  unsigned long long *keys = (unsigned long long *)scratch(nbins_ * 16 + kpc_k_scan_scratch_bytes(nbins_) + 512);
  unsigned long long *counts = keys + nbins_;
  void *scan = counts + nbins_;
  kpc_k_dense_extract(dense_lo_, dense_hi_, nbins_, keys, counts, d_tmp_ + 4, scan, compute_);
  launches_ += 3;
  rt_d2h(h_tmp_ + 4, d_tmp_ + 4, sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
  emit_entries(keys, counts, h_tmp_[4]);
}

unsigned long long KpcEngine::kmers_counted() {
  if (mode_ != DENSE) throw KpcError(KPC_E_UNSUPPORTED, "kpc_kmers_counted is only available on the dense-table path");
  // sum of the table == sum of all emitted counts
  unsigned long long *keys = (unsigned long long *)scratch(nbins_ * 16 + kpc_k_scan_scratch_bytes(nbins_) + 512);
  unsigned long long *counts = keys + nbins_;
  void *scan = counts + nbins_;
  kpc_k_dense_extract(dense_lo_, dense_hi_, nbins_, keys, counts, d_tmp_ + 4, scan, compute_);
  rt_d2h(h_tmp_ + 4, d_tmp_ + 4, sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
  const uint64_t n = h_tmp_[4];
  std::vector<unsigned long long> hc(n);
  if (n) rt_d2h(hc.data(), counts, n * 8, compute_);
  rt_stream_sync(compute_);
  unsigned long long s = 0;
  for (uint64_t i = 0; i < n; ++i) s += hc[i];
  return s;
}

void KpcEngine::dense_table(void **lo, void **hi, unsigned long long *nbins) {
  if (mode_ != DENSE) throw KpcError(KPC_E_STATE, "not on the dense-table path");
  rt_stream_sync(compute_);
  *lo = dense_lo_;
  *hi = dense_hi_;
  *nbins = nbins_;
}
unsigned long long KpcEngine::dense_max() {
  if (mode_ != DENSE) throw KpcError(KPC_E_STATE, "not on the dense-table path");
  rt_memset(d_tmp_ + 5, 0, 8, compute_);
  kpc_k_dense_max(dense_lo_, dense_hi_, nbins_, d_tmp_ + 5, compute_);
  ++launches_;
  rt_d2h(h_tmp_ + 5, d_tmp_ + 5, 8, compute_);
  rt_stream_sync(compute_);
  return h_tmp_[5];
}
unsigned long long KpcEngine::count_newlines_device(const uint8_t *dev, size_t n) {
  if (((uintptr_t)dev & 15) != 0) throw KpcError(KPC_E_ARG, "device input must be 16-byte aligned");
  rt_memset(d_tmp_ + 9, 0, 8, compute_);
  kpc_k_count_newlines(dev, n, d_tmp_ + 9, compute_);
  launches_ += n ? 1 : 0;
  rt_d2h(h_tmp_ + 9, d_tmp_ + 9, 8, compute_);
  rt_stream_sync(compute_);
  return h_tmp_[9];
}
void KpcEngine::dense_promote() {
  if (mode_ != DENSE) throw KpcError(KPC_E_STATE, "not on the dense-table path");
  if (!dense_hi_) {
    dense_hi_ = (unsigned long long *)rt_dmalloc(nbins_ * sizeof(unsigned long long));
    rt_memset(dense_hi_, 0, nbins_ * sizeof(unsigned long long), compute_);
  }
  kpc_k_dense_promote(dense_lo_, dense_hi_, nbins_, compute_);
  ++launches_;
  dense_since_fold_ = 0;
  rt_stream_sync(compute_);
}

void KpcEngine::synth_fastq(void *dev_out, unsigned long long first, unsigned long long n, unsigned long long seed) {
  kpc_k_synth_fastq((uint8_t *)dev_out, first, n, seed, compute_);
  rt_stream_sync(compute_);
}

// =================================================================================================
// HASH  (k > 12, DNA-ss k = 12, protein k > 4, or a small -M): table + insertion ranks + spill epochs
// =================================================================================================
void KpcEngine::hash_init() {
  d_hstat_ = (unsigned long long *)rt_dmalloc(2 * sizeof(unsigned long long));
  rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
  hcap_ = 0;
  hdistinct_ = 0;
}

KpcHashSink KpcEngine::hash_sink(uint64_t lo, uint64_t hi, long long sign) const {
  KpcHashSink h;
  h.keys = hkeys_; h.counts = hcounts_; h.ranks = hranks_;
  h.n_new = d_hstat_; h.overflow = d_hstat_ + 1;
  h.mask = hcap_ - 1;
  h.rank_lo = lo; h.rank_hi = hi; h.sign = sign;
  return h;
}

// room for `need` distinct keys at a load factor <= 1/2
void KpcEngine::hash_ensure_capacity(uint64_t need) {
  uint64_t want = pow2_at_least(1024, need * 2);
  if (want <= hcap_) return;
  unsigned long long *ok = hkeys_, *oc = hcounts_, *orr = hranks_;
  const uint64_t ocap = hcap_;
  hkeys_ = (unsigned long long *)rt_dmalloc(want * 8);
  hcounts_ = (unsigned long long *)rt_dmalloc(want * 8);
  hranks_ = (unsigned long long *)rt_dmalloc(want * 8);
  hcap_ = want;
  kpc_k_hash_clear(hkeys_, hcounts_, hranks_, hcap_, compute_);
  ++launches_;
  if (ocap) {
    rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
    KpcHashSink nw = hash_sink(0, ~0ull, 1);
    kpc_k_hash_rehash(ok, oc, orr, ocap, nw, compute_);
    ++launches_;
    rt_stream_sync(compute_);
    rt_dfree(ok); rt_dfree(oc); rt_dfree(orr);
  }
}

// dump in Hashtbl.iter order, then (optionally) Hashtbl.clear
void KpcEngine::hash_dump(bool clear) {
  if (hcap_) {
    const size_t need = hcap_ ? (size_t)std::min<uint64_t>(hcap_, hdistinct_ + 1) : 1;
    unsigned long long *ok = (unsigned long long *)scratch(need * 24 + kpc_k_scan_scratch_bytes(hcap_) + 1024);
    unsigned long long *oc = ok + need, *orr = oc + need;
    void *scan = orr + need;
    kpc_k_hash_extract(hkeys_, hcounts_, hranks_, hcap_, ok, oc, orr, d_tmp_ + 6, scan, compute_);
    launches_ += 3;
    rt_d2h(h_tmp_ + 6, d_tmp_ + 6, 8, compute_);
    rt_stream_sync(compute_);
    const uint64_t n = h_tmp_[6];
    const uint64_t B = grow_buckets(n);
    if (n) {
      void *os = scratch2(kpc_k_order_scratch_bytes(n));
      kpc_k_order_entries(ok, oc, orr, nullptr, n, B - 1, nullptr, 0, os, compute_);
      launches_ += 8;
      rt_stream_sync(compute_);
      emit_entries(ok, oc, n);
    }
    if (clear) {
      kpc_k_hash_clear(hkeys_, hcounts_, hranks_, hcap_, compute_);
      ++launches_;
      rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
      hdistinct_ = 0;
    }
  }
}

void KpcEngine::hash_process(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch,
                             uint64_t max_lines) {
  const uint64_t M = (uint64_t)cfg_.max_results_size;
  hash_ensure_capacity(hdistinct_ + len + 16);
  uint64_t lo = epoch_rank_lo_;
  uint64_t hi = ~0ull;
  auto run = [&](uint64_t rlo, uint64_t rhi, long long sign) {
    KpcHashSink h = hash_sink(rlo, rhi, sign);
    launch_tiles(st, mate, dev, len, final_launch, max_lines, KPC_SINK_HASH, &h, nullptr, false, nullptr, 0);
    rt_d2h(h_tmp_, d_hstat_, 2 * sizeof(unsigned long long), compute_);
    rt_d2h(h_tmp_ + 2, st.err_line, sizeof(unsigned long long), compute_);
    rt_stream_sync(compute_);
    if (h_tmp_[1]) throw KpcError(KPC_E_NOMEM, "internal: hash table overflow");
    hdistinct_ = h_tmp_[0];
  };
  run(lo, hi, 1);
  // a malformed FASTQ record ends the run there: nothing from that record on may influence the spill dumps
  if (format_ != KPC_FASTA && h_tmp_[2] != ~0ull && pair_limit_ < 0) {
    const uint64_t bad_rec = pe_weave_ ? (h_tmp_[2] / 4) & ~1ull : h_tmp_[2] / 4;  // woven pairs: the whole pair goes
    const uint64_t mates = format_ == KPC_FASTQ_PE ? 2 : 1;
    const uint64_t stop = (rank_base_ + bad_rec * mates) << 32;
    run(stop, ~0ull, -1);
    hi = stop;
  }
  // the dump/clear rule of bin/KPopCount.ml:39: after a record, if length res >= M
  while (hdistinct_ >= M) {
    // rank of the M-th distinct key of this epoch: the record that holds it is where the table reaches M
    const size_t need = (size_t)hdistinct_ + 1;
    unsigned long long *ok = (unsigned long long *)scratch(need * 24 + kpc_k_scan_scratch_bytes(hcap_) + 1024);
    unsigned long long *oc = ok + need, *orr = oc + need;
    void *scan = orr + need;
    kpc_k_hash_extract(hkeys_, hcounts_, hranks_, hcap_, ok, oc, orr, d_tmp_ + 6, scan, compute_);
    launches_ += 3;
    rt_d2h(h_tmp_ + 6, d_tmp_ + 6, 8, compute_);
    rt_stream_sync(compute_);
    const uint64_t n = h_tmp_[6];
    if (n < M) break;
    std::vector<unsigned long long> ranks(n);
    rt_d2h(ranks.data(), orr, n * 8, compute_);
    rt_stream_sync(compute_);
    std::nth_element(ranks.begin(), ranks.begin() + (M - 1), ranks.end());
    const uint64_t rstar = ranks[M - 1];
    // first rank of the record after the one that holds rstar
    uint64_t next_rec_rank = ~0ull;
    if (format_ != KPC_FASTA) {
      next_rec_rank = ((rstar >> 32) + 1) << 32;
      // is there anything at or after it in what has been fed?  (windows of later launches have larger ranks)
    } else {
      rt_memset(d_tmp_ + 7, 0xff, 8, compute_);
      const uint64_t from = rstar - rank_base_ + 1;  // stream offset just after the window's last symbol
      launch_tiles(st, mate, dev, len, final_launch, max_lines, KPC_SINK_NULL, nullptr, nullptr, false, d_tmp_ + 7, from);
      rt_d2h(h_tmp_ + 7, d_tmp_ + 7, 8, compute_);
      rt_stream_sync(compute_);
      if (h_tmp_[7] != ~0ull) next_rec_rank = rank_base_ + h_tmp_[7];
    }
    bool record_ends_here;
    if (format_ != KPC_FASTA) {
      // (paired-end files arrive here woven into one stream of complete pairs: see begin())
      // the record of rstar is over once the line feed of its sequence line has been seen
      rt_d2h(h_tmp_ + 32, st.carry[st.cur ^ 1], sizeof(KpcStreamCarry), compute_);
      rt_stream_sync(compute_);
      KpcStreamCarry co;
      memcpy(&co, h_tmp_ + 32, sizeof co);
      const uint64_t jstar = (rstar >> 32) - rank_base_;
      record_ends_here = final_launch || co.s1.count >= 4 * jstar + 2;
    } else {
      record_ends_here = next_rec_rank != ~0ull || final_launch;
    }
    if (!record_ends_here) break;  // the record continues in the next launch: decide there
    if (next_rec_rank < hi) {
      // take back everything from the next record on, dump, clear, and count it again into the fresh table
      run(next_rec_rank, hi, -1);
      hdistinct_ = n;  // entries that fell to zero are not printed; size for the bucket growth is recomputed in dump
      hash_dump(true);
      epoch_rank_lo_ = next_rec_rank;
      run(next_rec_rank, hi, 1);
    } else {
      hash_dump(true);  // nothing from a later record has been counted yet
      break;
    }
  }
  advance(st, len);
}

void KpcEngine::hash_finish() {
  if (sorted_pending_) {
    // the entries already stand in Hashtbl.iter order; filler slots (count 0) print nothing
    grow_buckets(sorted_distinct_);
    emit_entries(skeys_, scounts_, sorted_slots_);
    return;
  }
  hash_dump(false);
}

// =================================================================================================
// HASH, sort path: a sample that arrives as ONE launch and cannot reach -M (windows <= bytes < M) is counted by a
// counting sort on OCaml's bucket index instead of a hash table: two passes of the framing kernel (bucket sizes, then
// scatter), a scan, and a per-bucket merge that leaves the entries in Hashtbl.iter order (kpc_bucketsort.cuh).  This is
// the shape of BASELINE.json configs[3] (one assembled genome per KPopCount invocation, the reference's max k).
// =================================================================================================
bool KpcEngine::sort_usable(const StreamState &st, size_t len, bool final_launch) const {
  return sort_enabled_ && final_launch && st.fed == 0 && format_ != KPC_FASTQ_PE && !sorted_pending_ && hdistinct_ == 0 &&
         hcap_ == 0 && len > 0 && (uint64_t)len < (uint64_t)cfg_.max_results_size && len < ((size_t)1 << 31);
}

bool KpcEngine::sort_process(StreamState &st, int mate, const uint8_t *dev, size_t len, uint64_t max_lines) {
  // coarse buckets: the top bits of (key mod B); at most 2^22 of them (16 MiB of counters: L2 resident)
  int lb = 0;
  while ((1ull << lb) < buckets_) ++lb;
  const int lnb = lb < 22 ? lb : 22;
  const uint32_t nb = 1u << lnb;
  const int cshift = lb - lnb;
  const size_t heavy_cap = 4096;
  // layout: hist[nb] | offsets[nb + 1] | heavy[heavy_cap] | stats[4] | scan scratch | keys[len] | ranks[len] | counts[len]
  //         | pairs[len] (scatter target) | staged pairs[len]   (stats[3] counts the staged pairs)
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_hist = 0, o_off = o_hist + al(4 * (size_t)nb), o_heavy = o_off + al(4 * ((size_t)nb + 1)),
               o_stats = o_heavy + al(4 * heavy_cap), o_scan = o_stats + 256,
               o_keys = o_scan + al(kpc_k_scan_scratch_bytes(nb) + 512), o_ranks = o_keys + al(8 * len),
               o_counts = o_ranks + al(8 * len), o_pairs = o_counts + al(8 * len), o_stage = o_pairs + al(16 * len),
               total = o_stage + al(16 * len);
  if (total > sort_buf_cap_) {
    rt_stream_sync(compute_);
    rt_dfree(sort_buf_);
    sort_buf_cap_ = total + total / 8;
    sort_buf_ = rt_dmalloc(sort_buf_cap_);
  }
  char *B = (char *)sort_buf_;
  uint32_t *hist = (uint32_t *)(B + o_hist), *offsets = (uint32_t *)(B + o_off), *heavy = (uint32_t *)(B + o_heavy);
  unsigned long long *stats = (unsigned long long *)(B + o_stats);
  skeys_ = (unsigned long long *)(B + o_keys);
  sranks_ = (unsigned long long *)(B + o_ranks);
  scounts_ = (unsigned long long *)(B + o_counts);
  rt_memset(B, 0, o_scan, compute_);  // hist, offsets, heavy list, stats
  // ONE run of the framing machine: bucket sizes + the (key, rank) pairs in arrival order; the scatter streams them
  KpcBucketCountSink bc;
  bc.hist = hist; bc.bmask = buckets_ - 1; bc.cshift = cshift;
  bc.stage = (KpcPair *)(B + o_stage); bc.stage_n = stats + 3;
  launch_tiles(st, mate, dev, len, true, max_lines, KPC_SINK_BCOUNT, nullptr, nullptr, false, nullptr, 0, &bc, nullptr);
  kpc_k_bucket_offsets(hist, nb, offsets, B + o_scan, compute_);
  launches_ += 4;
  KpcBucketScatterSink bs;
  bs.remaining = hist; bs.offsets = offsets; bs.pairs = (KpcPair *)(B + o_pairs); bs.bmask = buckets_ - 1; bs.cshift = cshift;
  kpc_k_bucket_scatter_staged(bc.stage, bc.stage_n, bs, compute_);
  KpcBucketFinalize F;
  F.offsets = offsets; F.nb = nb; F.pairs = bs.pairs; F.keys = skeys_; F.ranks = sranks_; F.counts = scounts_; F.bmask = buckets_ - 1;
  F.heavy_list = heavy; F.heavy_cap = (uint32_t)heavy_cap; F.stats = stats;
  kpc_k_bucket_finalize(F, compute_);
  launches_ += 2;
  // [0] distinct keys, [1] groups too large for this path; the number of slots is the last offset
  rt_d2h(h_tmp_ + 40, stats, 2 * sizeof(unsigned long long), compute_);
  rt_d2h(h_tmp_ + 42, offsets + nb, sizeof(uint32_t), compute_);
  rt_stream_sync(compute_);
  if (h_tmp_[41]) return false;  // a heavily repeated k-mer: the hash table handles any multiplicity
  sorted_distinct_ = h_tmp_[40];
  sorted_slots_ = *(const uint32_t *)(h_tmp_ + 42);
  sorted_pending_ = true;
  hdistinct_ = sorted_distinct_;
  advance(st, len);
  return true;
}

void KpcEngine::set_record_base(unsigned long long first_record) {
  if (in_input_ || header_done_) throw KpcError(KPC_E_STATE, "kpc_set_record_base after the first kpc_begin");
  rank_base_ = first_record;
}
void KpcEngine::hash_export(void **keys, void **counts, void **ranks, unsigned long long *n_slots) {
  if (mode_ != HASH) throw KpcError(KPC_E_STATE, "not on the hash-table path");
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_hash_export inside an input");
  if (sorted_pending_) {  // the sort path's entries (fillers carry the empty key)
    rt_stream_sync(compute_);
    *keys = skeys_; *counts = scounts_; *ranks = sranks_; *n_slots = sorted_slots_;
    return;
  }
  if (!hcap_) { *keys = *counts = *ranks = nullptr; *n_slots = 0; return; }
  const size_t need = (size_t)std::min<uint64_t>(hcap_, hdistinct_ + 1);
  unsigned long long *ok = (unsigned long long *)scratch(need * 24 + kpc_k_scan_scratch_bytes(hcap_) + 1024);
  unsigned long long *oc = ok + need, *orr = oc + need;
  kpc_k_hash_extract(hkeys_, hcounts_, hranks_, hcap_, ok, oc, orr, d_tmp_ + 6, orr + need, compute_);
  launches_ += 3;
  rt_d2h(h_tmp_ + 6, d_tmp_ + 6, 8, compute_);
  rt_stream_sync(compute_);
  *keys = ok; *counts = oc; *ranks = orr; *n_slots = h_tmp_[6];
}
void KpcEngine::hash_import(const unsigned long long *keys, const unsigned long long *counts, const unsigned long long *ranks,
                            unsigned long long n, bool clear_first) {
  if (mode_ != HASH) throw KpcError(KPC_E_STATE, "not on the hash-table path");
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_hash_import inside an input");
  if (sorted_pending_ && !clear_first) sort_migrate();
  if (clear_first) {
    sorted_pending_ = false;
    if (hcap_) { kpc_k_hash_clear(hkeys_, hcounts_, hranks_, hcap_, compute_); ++launches_; }
    rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
    hdistinct_ = 0;
  }
  header_done_ = true;  // a context that only merges still owes its part of the dump
  hash_ensure_capacity(hdistinct_ + n + 16);
  if (n) {
    KpcHashSink nw = hash_sink(0, ~0ull, 1);
    kpc_k_hash_rehash(keys, counts, ranks, n, nw, compute_);
    ++launches_;
  }
  rt_d2h(h_tmp_, d_hstat_, 2 * sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
  if (h_tmp_[1]) throw KpcError(KPC_E_NOMEM, "internal: hash table overflow");
  hdistinct_ = h_tmp_[0];
}

// the sample counted by the sort path is followed by another input: its entries go into the hash table
void KpcEngine::sort_migrate() {
  sorted_pending_ = false;
  hdistinct_ = 0;
  hash_ensure_capacity(sorted_distinct_ + 16);
  rt_memset(d_hstat_, 0, 2 * sizeof(unsigned long long), compute_);
  KpcHashSink nw = hash_sink(0, ~0ull, 1);
  kpc_k_hash_rehash(skeys_, scounts_, sranks_, sorted_slots_, nw, compute_);  // filler slots carry the empty key
  ++launches_;
  rt_d2h(h_tmp_, d_hstat_, 2 * sizeof(unsigned long long), compute_);
  rt_stream_sync(compute_);
  if (h_tmp_[1]) throw KpcError(KPC_E_NOMEM, "internal: hash table overflow");
  hdistinct_ = h_tmp_[0];
}

// =================================================================================================
// TUPLE  (-L: one spectrum per record).  Synchronous per launch: this is not the throughput path.
// =================================================================================================
void KpcEngine::tuple_process(StreamState &st, int mate, const uint8_t *dev, size_t len, bool final_launch,
                              uint64_t max_lines) {
  const int mates = format_ == KPC_FASTQ_PE ? 2 : 1;
  // room for one tuple per byte on top of what is buffered
  if (tn_ + len + 16 > tcap_) {
    uint64_t ncap = std::max<uint64_t>(tn_ + len + 16, tcap_ * 2);
    unsigned long long *nk = (unsigned long long *)rt_dmalloc(ncap * 8);
    unsigned long long *nr = (unsigned long long *)rt_dmalloc(ncap * 8);
    uint32_t *nc = (uint32_t *)rt_dmalloc(ncap * 4);
    if (tn_) {
      rt_d2d(nk, tkeys_, tn_ * 8, compute_);
      rt_d2d(nr, tranks_, tn_ * 8, compute_);
      rt_d2d(nc, trecs_, tn_ * 4, compute_);
    }
    rt_stream_sync(compute_);
    rt_dfree(tkeys_); rt_dfree(tranks_); rt_dfree(trecs_);
    tkeys_ = nk; tranks_ = nr; trecs_ = nc; tcap_ = ncap;
  }
  // records that can show up in this launch start with the one that is open when it begins
  rt_d2h(h_tmp_ + 32, st.carry[st.cur], sizeof(KpcStreamCarry), compute_);
  rt_stream_sync(compute_);
  KpcStreamCarry cin;
  memcpy(&cin, h_tmp_ + 32, sizeof cin);
  const uint64_t rec_base = format_ == KPC_FASTA ? (cin.s1.count ? cin.s1.count - 1 : 0) : (cin.s1.count >> 2);
  const uint64_t rec_cap = len / 2 + 4;
  if (rec_cap > d_recs_cap_) {
    rt_dfree(d_recs_);
    d_recs_cap_ = rec_cap + rec_cap / 4;
    d_recs_ = (KpcRecEntry *)rt_dmalloc(d_recs_cap_ * sizeof(KpcRecEntry));
  }
  rt_memset(d_recs_, 0xff, d_recs_cap_ * sizeof(KpcRecEntry), compute_);
  KpcTupleSink ts;
  ts.keys = tkeys_; ts.ranks = tranks_; ts.recs = trecs_; ts.n_out = d_tn_; ts.cap = tcap_;
  ts.mates = (uint32_t)mates; ts.mate = (uint32_t)mate;
  tuple_rec_base_ = rec_base;
  launch_tiles(st, mate, dev, len, final_launch, max_lines, KPC_SINK_TUPLE, nullptr, &ts, true, nullptr, 0);
  rt_d2h(h_tmp_ + 32, st.carry[st.cur ^ 1], sizeof(KpcStreamCarry), compute_);
  rt_d2h(h_tmp_, d_tn_, 8, compute_);
  rt_stream_sync(compute_);
  KpcStreamCarry cout;
  memcpy(&cout, h_tmp_ + 32, sizeof cout);
  tn_ = h_tmp_[0];
  if (tn_ > tcap_) throw KpcError(KPC_E_NOMEM, "internal: tuple buffer overflow");
  // ---- record names: cut out of the host copy of this launch ----
  const uint64_t rec_end = format_ == KPC_FASTA ? cout.s1.count : (cout.s1.count >> 2) + 1;
  const uint64_t nrec = std::min<uint64_t>(rec_end > rec_base ? rec_end - rec_base : 0, d_recs_cap_);
  std::vector<KpcRecEntry> ent(nrec);
  if (nrec) {
    rt_d2h(ent.data(), d_recs_, nrec * sizeof(KpcRecEntry), compute_);
    rt_stream_sync(compute_);
  }
  std::vector<RecInfo> &open = open_recs_[mate];
  const uint64_t l0 = st.fed, l1 = st.fed + len;
  for (uint64_t j = 0; j < nrec; ++j) {
    const uint64_t r = rec_base + j;
    if (r < rec_done_[mate]) continue;
    const size_t idx = (size_t)(r - rec_done_[mate]);
    const bool starts = ent[j].tag_start != ~0ull;
    const bool known = idx < open.size() && open[idx].have_tag;
    if (!starts && !known) continue;
    if (idx >= open.size()) open.resize(idx + 1);
    RecInfo &ri = open[idx];
    uint64_t a = ~0ull, b = ~0ull;
    if (starts) { ri.tag_start = ent[j].tag_start; ri.have_tag = true; a = ent[j].tag_start; }
    else if (ri.tag_end == ~0ull) a = l0;  // a name that began in an earlier launch goes on
    if (a != ~0ull) {
      if (ent[j].tag_end != ~0ull) { ri.tag_end = ent[j].tag_end; b = ent[j].tag_end; }
      else b = l1;
      if (b > a) ri.tag.append((const char *)tag_bytes_.data() + (a - l0), (size_t)(b - a));
    }
  }
  advance(st, len);
  // ---- how far this mate's records are final ----
  if (format_ == KPC_FASTA) {
    st.records = cout.s1.count;
    st.final_recs = final_launch ? cout.s1.count : (cout.s1.count ? cout.s1.count - 1 : 0);
  } else {
    // all four lines seen (and, by the hold-back rule of choose_cut, the record is complete)
    st.final_recs = final_launch ? st.records : (cout.s1.count >> 2);
    if (pe_weave_) st.final_recs &= ~1ull;  // a read is only handed out once its whole pair has been checked
  }
  tuple_flush(false);
}

// write out every record that is final on all mates of the current input
void KpcEngine::tuple_flush(bool input_done) {
  const int mates = format_ == KPC_FASTQ_PE ? 2 : 1;
  uint64_t fin[2] = {streams_[0].final_recs, mates == 2 ? streams_[1].final_recs : 0};
  if (format_ != KPC_FASTA) {
    // never go past the first malformed record / pair (Files.ml:213-214, 241-243)
    for (int m = 0; m < mates; ++m) rt_d2h(h_tmp_ + m, streams_[m].err_line, 8, compute_);
    rt_stream_sync(compute_);
    uint64_t bad = ~0ull;
    for (int m = 0; m < mates; ++m)
      if (h_tmp_[m] != ~0ull) bad = std::min<uint64_t>(bad, h_tmp_[m] / 4);
    // woven pairs: a malformed mate takes its whole pair with it (FASTQ.iter_pe checks all eight lines before it hands
    // out either read, Files.ml:241-243)
    if (pe_weave_ && bad != ~0ull) bad &= ~1ull;
    if (input_done && mates == 2) {  // FASTQ.iter_pe stops at the shorter file
      const uint64_t pairs = std::min(streams_[0].records, streams_[1].records);
      fin[0] = std::min(fin[0], pairs);
      fin[1] = std::min(fin[1], pairs);
    }
    for (int m = 0; m < mates; ++m) fin[m] = std::min(fin[m], bad);
  }
  // positions in iteration order interleave the mates: g = record * mates + mate.  A pair is only visited when
  // both of its records are complete (FASTQ.iter_pe reads all eight lines first, Files.ml:228-239).
  const uint64_t g_hi = mates == 2 ? 2 * std::min(fin[0], fin[1]) : fin[0];
  const uint64_t g_lo = mates == 2 ? 2 * rec_done_[0] : rec_done_[0];
  if (g_hi <= g_lo) return;
  if (g_hi >= 0xffffffffull) throw KpcError(KPC_E_UNSUPPORTED, "more than 2^32 records in one -L input");
  const uint64_t nrec = g_hi - g_lo;
  // ---- reduce the buffered tuples: sort by (record, key), run-length ----
  uint64_t n_final = 0, keep_from = tn_;
  unsigned long long *ek = nullptr, *ec = nullptr, *er = nullptr;
  uint32_t *erec = nullptr;
  std::vector<unsigned long long> per_rec(nrec, 0);  // distinct k-mers of every record in [g_lo, g_hi)
  if (tn_) {
    const size_t n = (size_t)tn_;
    char *base = (char *)scratch(n * 28 + 4096);
    ek = (unsigned long long *)base;
    ec = ek + n;
    er = ec + n;
    erec = (uint32_t *)(er + n);
    void *os = scratch2(kpc_k_order_scratch_bytes(n));
    kpc_k_tuple_reduce(tkeys_, tranks_, trecs_, n, ek, ec, er, erec, d_tmp_ + 6, os, compute_);
    launches_ += 10;
    rt_d2h(h_tmp_ + 6, d_tmp_ + 6, 8, compute_);
    rt_stream_sync(compute_);
    const uint64_t n_entries = h_tmp_[6];
    // the tuple arrays are now sorted by record: those of records >= g_hi stay buffered.  Where the cut falls and how
    // many entries every record has is worked out on the device (two binary searches + a histogram over the entries)
    if (nrec) {
      const size_t need = (nrec + 2) * sizeof(unsigned long long);
      if (need > per_rec_cap_) {
        rt_stream_sync(compute_);
        if (d_per_rec_) { rt_dfree(d_per_rec_); rt_hfree(h_per_rec_); }
        per_rec_cap_ = need * 2;
        d_per_rec_ = (unsigned long long *)rt_dmalloc(per_rec_cap_);
        h_per_rec_ = (unsigned long long *)rt_hmalloc(per_rec_cap_);
      }
      kpc_k_tuple_bounds(trecs_, n, erec, n_entries, (uint32_t)g_lo, (uint32_t)g_hi, d_per_rec_, d_per_rec_ + 2, compute_);
      ++launches_;
      rt_d2h(h_per_rec_, d_per_rec_, need, compute_);
      rt_stream_sync(compute_);
      keep_from = h_per_rec_[0];
      n_final = h_per_rec_[1];
      for (uint64_t j = 0; j < nrec; ++j) per_rec[j] = h_per_rec_[2 + j];
    } else {
      keep_from = 0;
      n_final = 0;
    }
  }
  // ---- order inside each record: Hashtbl.iter with the bucket count the table has at that point ----
  const int bits = sbits_ * cfg_.k;
  std::vector<unsigned long long> bmask(nrec);
  bool grew = false;
  bool identity_order = true;
  for (uint64_t j = 0; j < nrec; ++j) {
    const uint64_t before = buckets_;
    grow_buckets(per_rec[j]);
    if (buckets_ != before) grew = true;
    bmask[j] = buckets_ - 1;
    if (bits >= 64 || (1ull << bits) > buckets_) identity_order = false;
  }
  if (bits < 64 && (1ull << bits) <= bmask[0] + 1) identity_order = true;  // B only grows
  if (n_final && !identity_order) {
    const size_t ob = (kpc_k_order_scratch_bytes(n_final) + 255) & ~(size_t)255;
    char *os = (char *)scratch2(ob + nrec * 8 + 256);
    unsigned long long *d_bmask = nullptr;
    if (grew) {
      d_bmask = (unsigned long long *)(os + ob);
      rt_h2d(d_bmask, bmask.data(), nrec * 8, compute_);
    }
    kpc_k_order_entries(ek, ec, er, erec, n_final, buckets_ - 1, d_bmask, (uint32_t)g_lo, os, compute_);
    launches_ += 10;
    rt_stream_sync(compute_);
  }
  // ---- text of all entries, then cut per record and put the header lines in between ----
  std::vector<char> lines;
  {
    struct Collect {
      static int fn(void *u, const char *b, size_t n) {
        std::vector<char> *v = (std::vector<char> *)u;
        v->insert(v->end(), b, b + n);
        return 0;
      }
    };
    kpc_sink_fn saved = sink_;
    void *saved_user = sink_user_;
    char *saved_buf = out_buf_;  // a caller-owned sink buffer receives the final text only
    sink_ = Collect::fn;
    sink_user_ = &lines;
    out_buf_ = nullptr;
    try {
      emit_entries(ek, ec, n_final);
    } catch (...) {
      sink_ = saved; sink_user_ = saved_user; out_buf_ = saved_buf;
      throw;
    }
    sink_ = saved;
    sink_user_ = saved_user;
    out_buf_ = saved_buf;
  }
  size_t lp = 0;
  bool quotes_error = false;
  std::string bad_tag, text;
  for (uint64_t j = 0; j < nrec; ++j) {
    const uint64_t g = g_lo + j;
    const int m = (int)(g % mates);
    const uint64_t r = g / mates;
    const std::vector<RecInfo> &open = open_recs_[m];
    const size_t idx = (size_t)(r - rec_done_[m]);
    const std::string tag = idx < open.size() ? open[idx].tag : std::string();
    size_t lend = lp;
    for (unsigned long long c = 0; c < per_rec[j]; ++c) {
      const char *q = (const char *)memchr(lines.data() + lend, '\n', lines.size() - lend);
      lend = (size_t)(q - lines.data()) + 1;
    }
    const bool skip = format_ == KPC_FASTA && tag.empty();  // Files.ml:101-106: records without a name are dropped
    if (!skip) {
      std::string clean;
      if (!strip_quotes(tag, clean)) { quotes_error = true; bad_tag = tag; break; }  // bin/KPopCount.ml:45
      text.assign("\t");
      text += clean;
      text += "\n";
      emit(text.data(), text.size());
      emit(lines.data() + lp, lend - lp);
    }
    lp = lend;
  }
  // ---- forget what has been written ----
  for (int m = 0; m < mates; ++m) {
    const uint64_t new_done = mates == 2 ? g_hi / 2 : g_hi;
    std::vector<RecInfo> &open = open_recs_[m];
    const size_t drop = (size_t)std::min<uint64_t>(new_done - rec_done_[m], open.size());
    open.erase(open.begin(), open.begin() + drop);
    rec_done_[m] = new_done;
  }
  if (tn_) {
    const uint64_t left = tn_ - keep_from;
    if (left && keep_from) {  // (source and destination may overlap: go through a scratch buffer)
      char *tmp = (char *)scratch2((size_t)left * 8 + 256);
      rt_d2d(tmp, tkeys_ + keep_from, left * 8, compute_);  rt_d2d(tkeys_, tmp, left * 8, compute_);
      rt_d2d(tmp, tranks_ + keep_from, left * 8, compute_); rt_d2d(tranks_, tmp, left * 8, compute_);
      rt_d2d(tmp, trecs_ + keep_from, left * 4, compute_);  rt_d2d(trecs_, tmp, left * 4, compute_);
    }
    tn_ = left;
    h_tmp_[0] = tn_;
    rt_h2d(d_tn_, h_tmp_, 8, compute_);
    rt_stream_sync(compute_);
  }
  if (quotes_error) {
    failed_ = true;
    throw KpcError(KPC_E_QUOTES_IN_NAME, "Quotes_in_name(\"" + bad_tag + "\")");
  }
}

void KpcEngine::tuple_finish() {
  // every record was written when it completed; the final KIHF.iter finds an empty table (bin/KPopCount.ml:49,60)
}
