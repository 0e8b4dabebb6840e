// kpc_multi.h -- several GPUs behind ONE context of the C ABI (kpc_create with n_devices > 1).
//
// The KPopCount path shards by read chunks: a FASTQ stream is cut at line starts, every chunk goes to one
// device -- round robin -- together with the index of the line it starts with, every device fills its own dense 4^k
// table, and kpc_finish sums the tables on the first device (peer copies over NVLink) before the dump.  Line indices
// come from a line-feed census taken on the device that received the chunk; only those 8-byte counts travel between
// the shards.  Everything that does not shard this way (FASTA, whose k-mers span lines and records; the hash-table and
// -L modes, whose output order depends on the order of insertion) runs on the first device alone, unchanged.
#pragma once
#include <memory>
#include <vector>

#include "kpc_engine.h"

class KpcMulti {
 public:
  KpcMulti(const KpcEngineConfig &cfg, const std::vector<int> &devices);
  ~KpcMulti();
  KpcEngine &first() { return *eng_[0]; }
  int n_devices() const { return (int)eng_.size(); }
  bool sharding() const { return shard_; }  // the current input is being spread over the devices

  int staging_slots() const { return kSlots; }
  void *staging(int slot, size_t *capacity);
  void begin(int format);
  void feed(int mate, const uint8_t *bytes, size_t n, bool eof);
  void end();
  void finish();
  void reset();

 private:
  static const int kSlots = 4;
  struct Pending {  // a chunk that has been uploaded but not counted yet (its census fixes the next chunk's line index)
    bool valid = false;
    int engine = 0, slot = 0;
    bool final_chunk = false;
  };
  struct MStream {
    uint8_t *hold[2] = {nullptr, nullptr};  // pinned; the lines kept back until they are known to be complete records
    size_t hold_cap[2] = {0, 0}, hold_len = 0;
    int hold_cur = 0;
    rt_event hold_ev[2] = {nullptr, nullptr};
    int hold_ev_engine[2] = {-1, -1};
    bool hold_busy[2] = {false, false};
    uint64_t lines = 0;        // line feeds of the chunks counted so far = line index of the next chunk
    uint64_t records = 0;
    bool eof = false, any = false;
    uint8_t last_byte = '\n';
    Pending pend;
  };
  void hold_append(MStream &st, const uint8_t *p, size_t n);
  void submit(MStream &st, int mate, const KpcEngine::Piece *pc, int npc, size_t len, bool final_chunk, int staging_slot);
  void settle(MStream &st, int mate, uint64_t max_lines);  // census of the pending chunk -> its launch
  void reduce_tables();

  KpcEngineConfig cfg_;
  std::vector<std::unique_ptr<KpcEngine>> eng_;
  bool shard_ = false, in_input_ = false, reduced_ = false;
  int format_ = -1;
  int next_engine_ = 0;
  size_t chunk_cap_ = 0;
  MStream st_[2];
  uint8_t *staging_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  rt_event staging_ev_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  int staging_ev_engine_[kSlots] = {-1, -1, -1, -1};
  bool staging_busy_[kSlots] = {false, false, false, false};
  long long pair_limit_ = -1;
  long long complete_pairs_ = -1;
  bool single_pass_ = false;

 public:
  void set_pair_limit(long long n) { pair_limit_ = n; for (auto &e : eng_) e->set_pair_limit(n); }
  void set_single_pass(bool on) { single_pass_ = on; for (auto &e : eng_) e->set_single_pass(on); }
  long long complete_pairs() const { return shard_ || complete_pairs_ >= 0 ? complete_pairs_ : eng_[0]->complete_pairs(); }
};
