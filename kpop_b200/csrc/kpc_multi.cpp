// kpc_multi.cpp -- see kpc_multi.h.
#include "kpc_multi.h"

#include <algorithm>
#include <cstring>

namespace {
// offset just behind the `want`-th line feed from the end of the first `limit` bytes of the pieces; 0 when there are fewer
size_t cut_after_nth_last_newline(const KpcEngine::Piece *pc, int npc, size_t limit, int want) {
  size_t starts[4], off = 0;
  for (int i = 0; i < npc; ++i) { starts[i] = off; off += pc[i].n; }
  int found = 0;
  for (int i = npc - 1; i >= 0; --i) {
    if (starts[i] >= limit) continue;
    size_t n = std::min(pc[i].n, limit - starts[i]);
    while (n > 0) {
      const void *q = memrchr(pc[i].p, '\n', n);
      if (!q) break;
      const size_t at = (size_t)((const uint8_t *)q - pc[i].p);
      if (++found == want) return starts[i] + at + 1;
      n = at;
    }
  }
  return 0;
}
}  // namespace

KpcMulti::KpcMulti(const KpcEngineConfig &cfg, const std::vector<int> &devices) : cfg_(cfg) {
  for (int d : devices) {
    KpcEngineConfig c = cfg;
    c.device = d;
    eng_.emplace_back(new KpcEngine(c));
  }
  chunk_cap_ = eng_[0]->chunk_cap();
  rt_set_device(eng_[0]->device());
  for (int m = 0; m < 2; ++m)
    for (int j = 0; j < 2; ++j) st_[m].hold_ev[j] = nullptr;
}

KpcMulti::~KpcMulti() {
  try {
    for (auto &e : eng_) { rt_set_device(e->device()); e->sync(); }
  } catch (...) {
  }
  rt_set_device(eng_[0]->device());
  for (int m = 0; m < 2; ++m)
    for (int j = 0; j < 2; ++j) {
      if (st_[m].hold[j]) rt_hfree(st_[m].hold[j]);
      if (st_[m].hold_ev[j]) { rt_set_device(eng_[st_[m].hold_ev_engine[j]]->device()); rt_event_destroy(st_[m].hold_ev[j]); }
    }
  for (int i = 0; i < kSlots; ++i) {
    if (staging_[i]) rt_hfree(staging_[i]);
    if (staging_ev_[i]) { rt_set_device(eng_[staging_ev_engine_[i]]->device()); rt_event_destroy(staging_ev_[i]); }
  }
  eng_.clear();
}

void *KpcMulti::staging(int slot, size_t *capacity) {
  if (slot < 0 || slot >= kSlots) throw KpcError(KPC_E_ARG, "staging slot out of range");
  if (!staging_[slot]) { rt_set_device(eng_[0]->device()); staging_[slot] = (uint8_t *)rt_hmalloc(chunk_cap_); }
  if (staging_busy_[slot]) {
    rt_event_sync(staging_ev_[slot]);
    staging_busy_[slot] = false;
  }
  if (capacity) *capacity = chunk_cap_;
  return staging_[slot];
}

void KpcMulti::reset() {
  for (auto &e : eng_) { rt_set_device(e->device()); e->reset(); }
  reduced_ = false;
  complete_pairs_ = -1;
}

void KpcMulti::begin(int format) {
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_begin while another input is open");
  if (reduced_) throw KpcError(KPC_E_STATE, "kpc_begin after kpc_finish (call kpc_reset first)");
  format_ = format;
  // what shards: FASTQ into the dense table.  Everything else is the first device's business alone.
  // (paired files in one pass are woven on the host, pair by pair: one stream, one device)
  shard_ = eng_.size() > 1 && format != KPC_FASTA && eng_[0]->mode() == KpcEngine::DENSE &&
           !(format == KPC_FASTQ_PE && single_pass_);
  in_input_ = true;
  if (!shard_) { rt_set_device(eng_[0]->device()); eng_[0]->begin(format); return; }
  for (size_t i = 0; i < eng_.size(); ++i) { rt_set_device(eng_[i]->device()); eng_[i]->shard_begin(format, i == 0); }
  for (int m = 0; m < 2; ++m) {
    MStream &st = st_[m];
    st.hold_len = 0; st.lines = 0; st.records = 0; st.eof = false; st.any = false; st.last_byte = '\n';
    st.pend.valid = false;
  }
}

void KpcMulti::hold_append(MStream &st, const uint8_t *p, size_t n) {
  const int c = st.hold_cur;
  if (st.hold_len + n > st.hold_cap[c]) {
    size_t ncap = std::max<size_t>(st.hold_len + n, st.hold_cap[c] * 2);
    ncap = std::max<size_t>(ncap, std::min<size_t>(chunk_cap_ + 64, (size_t)1 << 20));
    rt_set_device(eng_[0]->device());
    uint8_t *nb = (uint8_t *)rt_hmalloc(ncap);
    if (st.hold_len) memcpy(nb, st.hold[c], st.hold_len);
    if (st.hold[c]) rt_hfree(st.hold[c]);
    st.hold[c] = nb;
    st.hold_cap[c] = ncap;
  }
  if (n) memcpy(st.hold[c] + st.hold_len, p, n);
  st.hold_len += n;
}

// the pending chunk of the stream: its census is in by now (it was uploaded a chunk ago), count it
void KpcMulti::settle(MStream &st, int mate, uint64_t max_lines) {
  if (!st.pend.valid) return;
  KpcEngine &e = *eng_[st.pend.engine];
  rt_set_device(e.device());
  const uint64_t nl = e.shard_census(st.pend.slot);
  e.shard_count(st.pend.slot, mate, st.lines, max_lines);
  st.lines += nl;
  st.pend.valid = false;
}

void KpcMulti::submit(MStream &st, int mate, const KpcEngine::Piece *pc, int npc, size_t len, bool final_chunk,
                      int staging_slot) {
  // upload this chunk first (copies to different devices overlap), then count the chunk before it
  const int ei = next_engine_;
  next_engine_ = (next_engine_ + 1) % (int)eng_.size();
  KpcEngine &e = *eng_[ei];
  rt_set_device(e.device());
  // events that tell when the host buffers may be written again live on the device that copies from them
  auto event_for = [&](rt_event &slot_ev, int &slot_engine) {
    if (slot_ev && slot_engine != ei) {
      rt_set_device(eng_[slot_engine]->device());
      rt_event_destroy(slot_ev);
      slot_ev = nullptr;
      rt_set_device(e.device());
    }
    if (!slot_ev) { slot_ev = e.new_event(); slot_engine = ei; }
    return slot_ev;
  };
  int hold_used = -1;
  for (int i = 0; i < npc; ++i)
    for (int j = 0; j < 2; ++j)
      if (pc[i].n && pc[i].p == st.hold[j]) hold_used = j;
  rt_event ev1 = nullptr, ev2 = nullptr;
  if (staging_slot >= 0) { ev1 = event_for(staging_ev_[staging_slot], staging_ev_engine_[staging_slot]); staging_busy_[staging_slot] = true; }
  if (hold_used >= 0) { ev2 = event_for(st.hold_ev[hold_used], st.hold_ev_engine[hold_used]); st.hold_busy[hold_used] = true; }
  const int slot = e.shard_upload(pc, npc, len, ev1, ev2);
  settle(st, mate, ~0ull);  // every line of a chunk that is not the last one belongs to a complete record
  st.pend.valid = true;
  st.pend.engine = ei;
  st.pend.slot = slot;
  st.pend.final_chunk = final_chunk;
}

void KpcMulti::feed(int mate, const uint8_t *bytes, size_t n, bool eof) {
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_feed outside kpc_begin / kpc_end");
  if (!shard_) { rt_set_device(eng_[0]->device()); eng_[0]->feed(mate, bytes, n, eof); return; }
  if (mate < 0 || mate > 1 || (mate == 1 && format_ != KPC_FASTQ_PE)) throw KpcError(KPC_E_ARG, "bad mate index");
  MStream &st = st_[mate];
  if (st.eof) throw KpcError(KPC_E_STATE, "kpc_feed after eof");
  if (n) { st.any = true; st.last_byte = bytes[n - 1]; }
  int staging_slot = -1;
  for (int i = 0; i < kSlots; ++i)
    if (staging_[i] && bytes >= staging_[i] && bytes < staging_[i] + chunk_cap_) staging_slot = i;
  bool issued_from_caller = false;
  const uint8_t *src = bytes;
  size_t rem = n;
  for (;;) {
    const size_t pending = st.hold_len + rem;
    if (pending < chunk_cap_ || (pending == chunk_cap_ && eof)) break;
    KpcEngine::Piece pc[2];
    int npc = 0;
    if (st.hold_len) pc[npc++] = KpcEngine::Piece{st.hold[st.hold_cur], st.hold_len};
    if (rem) pc[npc++] = KpcEngine::Piece{src, rem};
    // a chunk ends behind the 5th line feed from the end of what is there: every line handed over is followed by four
    // more lines of the stream, so its record is complete whatever comes next (Files.ml:204-217)
    const size_t cut = cut_after_nth_last_newline(pc, npc, chunk_cap_, 5);
    if (cut == 0) throw KpcError(KPC_E_UNSUPPORTED, "lines longer than a fifth of the staging size in a multi-device run (raise KPC_CHUNK_BYTES)");
    if (cut <= st.hold_len) {  // the whole chunk comes out of the hold buffer (tiny staging sizes only)
      KpcEngine::Piece one{st.hold[st.hold_cur], cut};
      submit(st, mate, &one, 1, cut, false, -1);
      eng_[st.pend.engine]->sync_copy();
      st.hold_busy[st.hold_cur] = false;
      memmove(st.hold[st.hold_cur], st.hold[st.hold_cur] + cut, st.hold_len - cut);
      st.hold_len -= cut;
      continue;
    }
    const size_t from_src = cut - st.hold_len;
    KpcEngine::Piece two[2];
    int n2 = 0;
    if (st.hold_len) two[n2++] = KpcEngine::Piece{st.hold[st.hold_cur], st.hold_len};
    two[n2++] = KpcEngine::Piece{src, from_src};
    submit(st, mate, two, n2, cut, false, staging_slot);
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
    // everything that is left, plus a virtual line feed when the last line is unterminated (input_line returns it too)
    KpcEngine::Piece pc[3];
    int npc = 0;
    size_t len = 0;
    if (st.hold_len) { pc[npc++] = KpcEngine::Piece{st.hold[st.hold_cur], st.hold_len}; len += st.hold_len; }
    if (rem) { pc[npc++] = KpcEngine::Piece{src, rem}; len += rem; issued_from_caller = true; }
    static const uint8_t nl = '\n';
    if (st.any && st.last_byte != '\n') { pc[npc++] = KpcEngine::Piece{&nl, 1}; len += 1; }
    if (len) submit(st, mate, pc, npc, len, true, rem ? staging_slot : -1);
    st.hold_len = 0;
    // the last chunk: only complete records count (FASTQ.iter_se drops a record cut by the end of the file)
    if (st.pend.valid) {
      KpcEngine &e = *eng_[st.pend.engine];
      rt_set_device(e.device());
      const uint64_t total = st.lines + e.shard_census(st.pend.slot);
      st.records = total / 4;
      settle(st, mate, st.records * 4);
    } else {
      st.records = st.lines / 4;
    }
  }
  if (issued_from_caller && staging_slot < 0) {
    for (auto &e : eng_) { rt_set_device(e->device()); e->sync_copy(); }  // the caller may reuse its buffer as soon as we return
  }
}

void KpcMulti::end() {
  if (!in_input_) throw KpcError(KPC_E_STATE, "kpc_end without kpc_begin");
  if (!shard_) { in_input_ = false; rt_set_device(eng_[0]->device()); eng_[0]->end(); return; }
  const int mates = format_ == KPC_FASTQ_PE ? 2 : 1;
  for (int m = 0; m < mates; ++m)
    if (!st_[m].eof) throw KpcError(KPC_E_STATE, "kpc_end before eof was signalled on every mate");
  in_input_ = false;
  uint64_t bad[2] = {~0ull, ~0ull};
  for (auto &e : eng_) {
    rt_set_device(e->device());
    for (int m = 0; m < mates; ++m) bad[m] = std::min<uint64_t>(bad[m], e->shard_err_line(m));
    e->shard_end();
  }
  uint64_t recs = st_[0].records;
  if (mates == 2) {
    const uint64_t r0 = st_[0].records, r1 = st_[1].records;
    recs = std::min(r0, r1);
    if (pair_limit_ >= 0) recs = std::min<uint64_t>(recs, (uint64_t)pair_limit_);
    complete_pairs_ = (long long)recs;
    const uint64_t used0 = pair_limit_ >= 0 ? std::min<uint64_t>(r0, (uint64_t)pair_limit_) : r0;
    const uint64_t used1 = pair_limit_ >= 0 ? std::min<uint64_t>(r1, (uint64_t)pair_limit_) : r1;
    if (used0 != used1)
      throw KpcError(KPC_E_PE_MISMATCH, "paired FASTQ files hold different numbers of records (" + std::to_string(r0) +
                                            " and " + std::to_string(r1) + ")");
  }
  uint64_t first_bad = ~0ull;
  for (int m = 0; m < mates; ++m)
    if (bad[m] != ~0ull && bad[m] / 4 < recs) first_bad = std::min<uint64_t>(first_bad, bad[m] / 4);
  if (first_bad != ~0ull)
    throw KpcError(KPC_E_MALFORMED_FASTQ, "On line " + std::to_string((first_bad + 1) * 4 * mates) + ": Malformed FASTQ file");
}

// the tables of the other devices are added to the first one's (64-bit: no width questions), over peer copies
void KpcMulti::reduce_tables() {
  if (eng_.size() < 2 || eng_[0]->mode() != KpcEngine::DENSE || reduced_) return;
  reduced_ = true;
  KpcEngine &e0 = *eng_[0];
  for (size_t i = 1; i < eng_.size(); ++i) {
    KpcEngine &e = *eng_[i];
    rt_set_device(e.device());
    e.dense_promote();  // hi += lo, lo = 0; synchronises
    void *lo, *hi;
    unsigned long long nbins;
    e.dense_table(&lo, &hi, &nbins);
    rt_set_device(e0.device());
    e0.dense_add_remote((const unsigned long long *)hi, nbins, e.device());
  }
}

void KpcMulti::finish() {
  if (in_input_) throw KpcError(KPC_E_STATE, "kpc_finish inside an input");
  reduce_tables();
  rt_set_device(eng_[0]->device());
  eng_[0]->finish();
}
