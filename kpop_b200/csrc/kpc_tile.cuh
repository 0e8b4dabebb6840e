// kpc_tile.cuh -- the per-tile "framing + rolling k-mer" machine of the KPopCount hot path.
//
// One CTA owns one tile (NT threads x SEG bytes) of the raw FASTA/FASTQ byte stream at a time and does,
// in one pass over HBM:
//   1. record / line framing          (replaces Files.FASTA.iter / FASTQ.iter_se, Files.ml:96-122, 201-221)
//   2. linting + symbol coding        (replaces Sequences.Lint.dnaize/proteinize, Sequences.ml:41-67, 87-151)
//   3. rolling forward / reverse-complement k-mer construction and canonicalisation
//                                     (replaces DNAHash*.iteri/iterc, ProteinHash.iteri, KMers.ml:228-257, 319-389)
//   4. hands every fully valid window to a Sink (dense table RED, hash insert, tuple append)
//                                     (replaces IntHashFrequencies.add, KMers.ml:107-111)
//
// Framing state that crosses tile boundaries is carried by a single-pass chained scan ("decoupled
// look-back") over per-tile descriptors, in two stages:
//   stage 1  S1 = { last line-feed position (max), last header start (max), line / header count (sum) }
//   stage 2  KCarry = the last <= k-1 valid symbols before the tile end since the last break
// Windows that use symbols from before the tile start are produced by one "fix-up" thread from the
// stage-2 carry; all other threads only see symbols inside the tile.
//
// Everything here is KPC_HD (host + device) and written as per-thread *phase* functions separated by
// CTA-wide barriers.  The CUDA kernel (kpc_kernels.cu) calls the phases with __syncthreads() between
// them; the test-only emulation harness (tests/emul) calls the same functions thread by thread on the
// CPU, so the framing logic can be fuzzed against the oracle without a GPU.
#pragma once
#include "kpc_common.h"

#if defined(__CUDA_ARCH__)
#define KPC_ON_DEVICE 1
#else
#define KPC_ON_DEVICE 0
#endif

// ------------------------------------------------------------------------------------------------
// small portability layer: atomics and release/acquire flags (host versions are single-threaded)
// ------------------------------------------------------------------------------------------------
KPC_HD void kpc_red_add_u32(uint32_t *p, uint32_t v) {
#if KPC_ON_DEVICE
  atomicAdd(p, v);  // result unused: ptxas emits RED.E.ADD
#else
  *p += v;
#endif
}
KPC_HD uint32_t kpc_atomic_add_u32(uint32_t *p, uint32_t v) {
#if KPC_ON_DEVICE
  return atomicAdd(p, v);
#else
  uint32_t o = *p; *p += v; return o;
#endif
}
KPC_HD unsigned long long kpc_atomic_add_u64(unsigned long long *p, unsigned long long v) {
#if KPC_ON_DEVICE
  return atomicAdd(p, v);
#else
  unsigned long long o = *p; *p += v; return o;
#endif
}
KPC_HD void kpc_atomic_min_u64(unsigned long long *p, unsigned long long v) {
#if KPC_ON_DEVICE
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}
KPC_HD unsigned long long kpc_atomic_cas_u64(unsigned long long *p, unsigned long long cmp, unsigned long long val) {
#if KPC_ON_DEVICE
  return atomicCAS(p, cmp, val);
#else
  unsigned long long o = *p; if (o == cmp) *p = val; return o;
#endif
}
KPC_HD void kpc_flag_release(uint32_t *flag, uint32_t v) {
#if KPC_ON_DEVICE
  __threadfence();
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
#else
  *flag = v;
#endif
}
KPC_HD uint32_t kpc_flag_acquire(const uint32_t *flag) {
#if KPC_ON_DEVICE
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  return v;
#else
  return *flag;
#endif
}
KPC_HD uint64_t kpc_ld_cg_u64(const uint64_t *p) {
#if KPC_ON_DEVICE
  return __ldcg((const unsigned long long *)p);
#else
  return *p;
#endif
}
KPC_HD uint32_t kpc_ld_cg_u32(const uint32_t *p) {
#if KPC_ON_DEVICE
  return __ldcg(p);
#else
  return *p;
#endif
}

// ------------------------------------------------------------------------------------------------
// scan state
// ------------------------------------------------------------------------------------------------
// positions are absolute offsets inside the current input stream, stored +1 (0 = "none yet")
struct KpcS1 {
  uint64_t last_nl;   // max: position+1 of the most recent '\n'
  uint64_t last_hdr;  // max: FASTA only, ((position+1) << 1) | name_is_not_empty of the most recent '>' at a line start
  uint64_t count;     // sum: FASTQ = number of '\n' (line index), FASTA = number of header lines started
};
KPC_HD KpcS1 kpc_s1_identity() { KpcS1 s; s.last_nl = 0; s.last_hdr = 0; s.count = 0; return s; }
KPC_HD KpcS1 kpc_s1_combine(const KpcS1 &x, const KpcS1 &y) {  // x happens before y
  KpcS1 r;
  r.last_nl = x.last_nl > y.last_nl ? x.last_nl : y.last_nl;
  r.last_hdr = x.last_hdr > y.last_hdr ? x.last_hdr : y.last_hdr;
  r.count = x.count + y.count;
  return r;
}
// last <= k-1 valid symbols (most recent in the low bits) since the last break
struct KpcKCarry {
  uint64_t syms;
  uint32_t n;       // number of symbols held
  uint32_t closed;  // 1: a break (or k-1 symbols) was seen, nothing older can matter
};
KPC_HD KpcKCarry kpc_kc_identity() { KpcKCarry c; c.syms = 0; c.n = 0; c.closed = 0; return c; }
KPC_HD KpcKCarry kpc_kc_combine(const KpcKCarry &older, const KpcKCarry &newer, int k, int sbits) {
  if (newer.closed) return newer;
  KpcKCarry r;
  uint32_t need = (uint32_t)(k - 1) - newer.n;
  uint32_t take = older.n < need ? older.n : need;
  uint64_t m = take ? (~0ull >> (64 - take * sbits)) : 0ull;
  r.syms = newer.syms | ((older.syms & m) << (newer.n * sbits));
  r.n = newer.n + take;
  r.closed = (older.closed || r.n == (uint32_t)(k - 1)) ? 1u : 0u;
  return r;
}

// one descriptor per tile; flags hold (epoch << 2) | status so they never need clearing between launches
enum { KPC_ST_AGG = 1, KPC_ST_INC = 2 };
struct KpcTileDesc {
  uint32_t flag1, flag2;
  uint64_t agg1[3], inc1[3];
  uint64_t kagg_syms, kinc_syms;
  uint32_t kagg_n, kagg_closed, kinc_n, kinc_closed;
};
// state at a launch boundary (end of the previous launch of the same input stream)
struct KpcStreamCarry {
  KpcS1 s1;
  KpcKCarry kc;
  uint32_t last_byte;  // byte just before the launch ('\n' at the start of a stream)
  uint32_t pad;
};
// -L mode: where the record names are (absolute stream offsets; ~0 = not seen in this launch)
struct KpcRecEntry {
  unsigned long long tag_start, tag_end;
};

struct KpcTileParams {
  const uint8_t *data;   // launch bytes; 16-byte aligned; readable up to the next 16-byte boundary past n
  uint64_t n;            // number of bytes in this launch
  uint64_t abs_base;     // stream offset of data[0]
  uint64_t rank_base;    // insertion ranks: FASTA rank = rank_base + stream offset of the window's last symbol;
  uint32_t rank_mates;   //   FASTQ rank = ((rank_base + record * rank_mates + rank_mate) << 32) | offset in line
  uint32_t rank_mate;    //   (orders windows as ReadsIterate.iter visits them: pair, mate, position)
  uint64_t max_lines;    // FASTQ: bytes on lines >= max_lines are ignored (incomplete last record, -p cap)
  int k;
  uint32_t epoch;
  uint32_t n_tiles;
  int final_launch;      // last launch of the stream: data[n] is treated as end of file
  KpcTileDesc *desc;
  const KpcStreamCarry *carry_in;
  KpcStreamCarry *carry_out;
  uint32_t *tile_counter;
  unsigned long long *err_line;   // FASTQ: atomicMin of the first malformed line index
  KpcRecEntry *rec_tab;           // optional (-L): indexed by record - rec_base
  uint64_t rec_base;
  uint64_t rec_cap;
  unsigned long long *probe_pos;  // optional: atomicMin of the first record start >= probe_from
  uint64_t probe_from;
};

// ------------------------------------------------------------------------------------------------
// sinks
// ------------------------------------------------------------------------------------------------
struct KpcDenseSink {  // k small: 4^k (or 32^k) u32 bins, L2 resident at k = 12
  uint32_t *table;
  KPC_HD void emit(uint64_t key, uint64_t /*rank*/, uint64_t /*rec*/) const { kpc_red_add_u32(table + key, 1u); }
};
struct KpcNullSink {
  KPC_HD void emit(uint64_t, uint64_t, uint64_t) const {}
};
// open addressing, linear probing; EMPTY key = ~0 (keys use <= 60 bits)
struct KpcHashSink {
  unsigned long long *keys;
  unsigned long long *counts;  // two's complement: sign = -1 un-counts a range again
  unsigned long long *ranks;   // min insertion rank of the key
  unsigned long long *n_new;   // number of slots claimed (distinct keys)
  unsigned long long *overflow;  // set when the table is full
  uint64_t mask;               // capacity - 1
  uint64_t rank_lo, rank_hi;   // only windows with rank_lo <= rank < rank_hi are considered
  long long sign;
  KPC_HD static uint64_t hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
  }
  // delta is added to the key's counter (two's complement); a new slot is only claimed when may_claim
  KPC_HD void insert(uint64_t key, uint64_t rank, unsigned long long delta, bool may_claim) const {
    uint64_t slot = hash(key) & mask;
    for (uint64_t probes = 0; probes <= mask; ++probes) {
      unsigned long long cur = kpc_ld_cg_u64((const uint64_t *)(keys + slot));
      if (cur == ~0ull) {
        if (!may_claim) return;
        cur = kpc_atomic_cas_u64(keys + slot, ~0ull, key);
        if (cur == ~0ull) { kpc_atomic_add_u64(n_new, 1ull); cur = key; }
      }
      if (cur == key) {
        kpc_atomic_add_u64(counts + slot, delta);
        if (may_claim) kpc_atomic_min_u64(ranks + slot, rank);
        return;
      }
      slot = (slot + 1) & mask;
    }
    kpc_atomic_add_u64(overflow, 1ull);
  }
  KPC_HD void emit(uint64_t key, uint64_t rank, uint64_t /*rec*/) const {
    if (rank < rank_lo || rank >= rank_hi) return;
    insert(key, rank, (unsigned long long)sign, sign > 0);
  }
};
// -L mode: (record, key, rank) tuples, reduced by sort + run-length on the device afterwards.
// The record id stored is the position of the record in ReadsIterate.iter order inside the current
// input: record * mates + mate (paired-end files interleave, Files.ml:363-368).
struct KpcTupleSink {
  unsigned long long *keys;
  unsigned long long *ranks;
  uint32_t *recs;
  unsigned long long *n_out;
  uint64_t cap;
  uint32_t mates, mate;
  KPC_HD void emit(uint64_t key, uint64_t rank, uint64_t rec) const {
    unsigned long long i = kpc_atomic_add_u64(n_out, 1ull);
    if (i < cap) { keys[i] = key; ranks[i] = rank; recs[i] = (uint32_t)(rec * mates + mate); }
  }
};

// Sort path (a whole sample in one launch, large k: DESIGN.md section 6): two passes over the input instead of a hash
// table.  Pass 1 counts the windows of every coarse bucket -- the top bits of (key mod B), B = OCaml's bucket count --,
// pass 2 drops every window into its bucket's range of one (key, rank) array (the order inside a range is arbitrary:
// kpc_bucketsort.cuh sorts, merges duplicates and restores Hashtbl.iter order range by range).
// one k-mer window of the sort path: its key and its rank (position in input order); one 16-byte store / load
struct alignas(16) KpcPair { unsigned long long key, rank; };
struct KpcBucketCountSink {
  uint32_t *hist;
  uint64_t bmask;  // B - 1
  int cshift;      // coarse bucket = (key & bmask) >> cshift
  // optional staging: the (key, rank) pairs in arrival order, so that the scatter pass is a plain streaming kernel
  // (kpc_k_bucket_scatter_staged) instead of a second run of the framing machine
  KpcPair *stage;
  unsigned long long *stage_n;
  KPC_HD void emit(uint64_t key, uint64_t rank, uint64_t /*rec*/) const {
    kpc_red_add_u32(hist + ((key & bmask) >> cshift), 1u);
    if (stage) {
#if KPC_ON_DEVICE
      // one atomic per warp: the lanes that are here together take consecutive places
      const unsigned m = __activemask();
      const int leader = __ffs(m) - 1, lane = (int)(threadIdx.x & 31u);
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(stage_n, (unsigned long long)__popc(m));
      base = __shfl_sync(m, base, leader);
      const unsigned long long i = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
#else
      const unsigned long long i = kpc_atomic_add_u64(stage_n, 1ull);
#endif
      KpcPair pr; pr.key = key; pr.rank = rank;
      stage[i] = pr;
    }
  }
};
struct KpcBucketScatterSink {
  uint32_t *remaining;       // pass 1's counts, counted down to zero here
  const uint32_t *offsets;   // exclusive prefix sums of the counts
  KpcPair *pairs;
  uint64_t bmask;
  int cshift;
  KPC_HD void emit(uint64_t key, uint64_t rank, uint64_t /*rec*/) const {
    const uint64_t b = (key & bmask) >> cshift;
    const uint32_t left = kpc_atomic_add_u32(remaining + b, 0xFFFFFFFFu);  // -1
    const uint64_t pos = (uint64_t)offsets[b] + left - 1u;
    KpcPair pr; pr.key = key; pr.rank = rank;
    pairs[pos] = pr;
  }
};

// ------------------------------------------------------------------------------------------------
// shared memory of one CTA
// ------------------------------------------------------------------------------------------------
template <int NT_, int SEG_>
struct KpcTileShared {
  static const int NT = NT_, SEG = SEG_, T = NT_ * SEG_;
  static const int GS = NT_ < 32 ? NT_ : 32;  // scan group = one warp on the device
  static const int NG = NT_ / GS;
  // Tile byte i lives at raw[px(i)]: 4 bytes of padding after every SEG bytes, so that the threads of a warp -- each
  // walking its own SEG consecutive bytes -- hit different banks (with SEG = 64 and no padding thread t reads word
  // 16 t + c: two banks for the whole warp, a 16-way conflict on every byte access).  raw[px(len)] = byte after the
  // tile, `before` = byte before it; cls[] (one class code per byte) uses the same layout.
  static const int TP = T + (T / SEG_) * 4 + 8;
  KPC_HD static int px(int i) { return i + (i / SEG_) * 4; }
  alignas(16) uint8_t raw[TP];
  alignas(16) uint8_t cls[TP];
  uint8_t before;
  KpcS1 pre[NT];     // per-thread census, then exclusive prefix inside its group
  KpcS1 gpre[NG];    // group totals, then exclusive prefix over groups
  KpcS1 tile_in;     // absolute state at the tile start
  KpcKCarry ksum;    // this tile's k-mer carry summary
  uint32_t tile;
  uint32_t len;      // bytes of this tile (T except for the last one)
};

// ------------------------------------------------------------------------------------------------
// the machine
// ------------------------------------------------------------------------------------------------
template <int NT, int SEG, int FMT, int CONTENT>
struct KpcTileMachine {
  typedef KpcTileShared<NT, SEG> Shared;
  static const int T = NT * SEG;
  static const int SB = (CONTENT == KPC_CONTENT_PROTEIN) ? 5 : 2;

  KPC_HD static uint8_t classify_symbol(uint8_t b) {
    return CONTENT == KPC_CONTENT_PROTEIN ? kpc_classify_protein(b) : kpc_classify_dna(b);
  }
  KPC_HD static uint64_t tile_start(const Shared &sh) { return (uint64_t)sh.tile * (uint64_t)T; }

  // ---- phase 0: claim a tile (thread 0), returns false when the launch is exhausted -------------
  KPC_HD static void claim(Shared &sh, const KpcTileParams &p, int tid) {
    if (tid == 0) {
      sh.tile = kpc_atomic_add_u32(p.tile_counter, 1u);
      if (sh.tile < p.n_tiles) {
        uint64_t rem = p.n - (uint64_t)sh.tile * (uint64_t)T;
        sh.len = rem < (uint64_t)T ? (uint32_t)rem : (uint32_t)T;
      } else {
        sh.len = 0;
      }
    }
  }

  // ---- phase 1: tile bytes -> shared memory (coalesced 128-bit loads) -----------------------------
  KPC_HD static void load(Shared &sh, const KpcTileParams &p, int tid) {
    const uint64_t t0 = tile_start(sh);
    const uint8_t *src = p.data + t0;
    const int nvec = (int)((sh.len + 15u) >> 4);
    for (int v = tid; v < nvec; v += NT) {
#if KPC_ON_DEVICE
      uint4 x;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                   : "l"(src + 16 * v));
      if (SEG % 4 == 0) {  // a 4-byte word never straddles a padding gap
        *reinterpret_cast<uint32_t *>(sh.raw + Shared::px(16 * v)) = x.x;
        *reinterpret_cast<uint32_t *>(sh.raw + Shared::px(16 * v + 4)) = x.y;
        *reinterpret_cast<uint32_t *>(sh.raw + Shared::px(16 * v + 8)) = x.z;
        *reinterpret_cast<uint32_t *>(sh.raw + Shared::px(16 * v + 12)) = x.w;
      } else {
        const uint32_t wds[4] = {x.x, x.y, x.z, x.w};
        for (int j = 0; j < 16; ++j) sh.raw[Shared::px(16 * v + j)] = (uint8_t)(wds[j >> 2] >> (8 * (j & 3)));
      }
#else
      for (int j = 0; j < 16; ++j) {
        uint64_t q = t0 + 16 * (uint64_t)v + j;
        sh.raw[Shared::px(16 * v + j)] = q < p.n ? src[16 * v + j] : (uint8_t)'\n';
      }
#endif
    }
    if (tid == 0) {
      // the byte before the tile: previous tile, previous launch, or a virtual '\n' before the stream
      uint8_t before;
      if (t0 > 0) before = p.data[t0 - 1];
      else before = (uint8_t)kpc_ld_cg_u32(&p.carry_in->last_byte);
      sh.before = before;
    }
    if (nvec > 0 && tid == (nvec - 1) % NT) {
      // the byte after the tile (look-ahead for "is the record name empty"), written by the thread that
      // loaded the last vector so that it cannot race with that store.  The end of a launch reads as '\n':
      // true end of file, or a cut the host made just after a line feed (never after a '>' that starts a
      // line, see kpc_engine.cpp).  Bytes past the tile inside the last vector are never interpreted.
      uint64_t q = t0 + sh.len;
      sh.raw[Shared::px((int)sh.len)] = q >= p.n ? (uint8_t)'\n' : p.data[q];
    }
  }

  // tile byte i, -1 <= i <= len
  KPC_HD static uint8_t rawb(const Shared &sh, int i) { return i < 0 ? sh.before : sh.raw[Shared::px(i)]; }
  KPC_HD static int seg_lo(const Shared &sh, int tid) { int a = tid * SEG; return a < (int)sh.len ? a : (int)sh.len; }
  KPC_HD static int seg_hi(const Shared &sh, int tid) { int a = tid * SEG + SEG; return a < (int)sh.len ? a : (int)sh.len; }

  // ---- phase 2: per-thread census of its segment ---------------------------------------------------
  KPC_HD static void census(Shared &sh, const KpcTileParams &p, int tid) {
    const int lo = seg_lo(sh, tid), hi = seg_hi(sh, tid);
    const uint64_t abs0 = p.abs_base + tile_start(sh);
    KpcS1 s = kpc_s1_identity();
    uint8_t prev = rawb(sh, lo - 1);
    for (int i = lo; i < hi; ++i) {
      uint8_t b = sh.raw[Shared::px(i)];
      if (b == '\n') {
        s.last_nl = abs0 + i + 1;
        if (FMT == KPC_FMT_FASTQ) s.count++;
      } else if (FMT == KPC_FMT_FASTA && b == '>' && prev == '\n') {
        uint64_t named = sh.raw[Shared::px(i + 1)] != '\n' ? 1u : 0u;
        s.last_hdr = ((abs0 + i + 1) << 1) | named;
        s.count++;
      }
      prev = b;
    }
    sh.pre[tid] = s;
  }

  // ---- phase 3a: exclusive scan inside each group of GS threads ------------------------------------
  KPC_HD static void scan_groups(Shared &sh, int tid) {
    if (tid % Shared::GS == 0) {
      KpcS1 acc = kpc_s1_identity();
      for (int j = 0; j < Shared::GS; ++j) {
        KpcS1 v = sh.pre[tid + j];
        sh.pre[tid + j] = acc;
        acc = kpc_s1_combine(acc, v);
      }
      sh.gpre[tid / Shared::GS] = acc;
    }
  }

  // ---- phase 3b: thread 0 scans the group totals, publishes the tile aggregate, looks back ---------
  KPC_HD static void lookback1(Shared &sh, const KpcTileParams &p, int tid) {
    if (tid != 0) return;
    KpcS1 acc = kpc_s1_identity();
    for (int g = 0; g < Shared::NG; ++g) {
      KpcS1 v = sh.gpre[g];
      sh.gpre[g] = acc;
      acc = kpc_s1_combine(acc, v);
    }
    KpcTileDesc *d = p.desc + sh.tile;
    d->agg1[0] = acc.last_nl; d->agg1[1] = acc.last_hdr; d->agg1[2] = acc.count;
    kpc_flag_release(&d->flag1, (p.epoch << 2) | KPC_ST_AGG);
    // look back for the exclusive prefix
    KpcS1 pre = kpc_s1_identity();
    long long j = (long long)sh.tile - 1;
    for (;;) {
      if (j < 0) {
        KpcS1 c;
        c.last_nl = kpc_ld_cg_u64(&p.carry_in->s1.last_nl);
        c.last_hdr = kpc_ld_cg_u64(&p.carry_in->s1.last_hdr);
        c.count = kpc_ld_cg_u64(&p.carry_in->s1.count);
        pre = kpc_s1_combine(c, pre);
        break;
      }
      const KpcTileDesc *e = p.desc + j;
      uint32_t f;
      do { f = kpc_flag_acquire(&e->flag1); } while ((f >> 2) != p.epoch);
      KpcS1 v;
      if ((f & 3u) == KPC_ST_INC) {
        v.last_nl = kpc_ld_cg_u64(&e->inc1[0]); v.last_hdr = kpc_ld_cg_u64(&e->inc1[1]); v.count = kpc_ld_cg_u64(&e->inc1[2]);
        pre = kpc_s1_combine(v, pre);
        break;
      }
      v.last_nl = kpc_ld_cg_u64(&e->agg1[0]); v.last_hdr = kpc_ld_cg_u64(&e->agg1[1]); v.count = kpc_ld_cg_u64(&e->agg1[2]);
      pre = kpc_s1_combine(v, pre);
      --j;
    }
    sh.tile_in = pre;
    KpcS1 inc = kpc_s1_combine(pre, acc);
    d->inc1[0] = inc.last_nl; d->inc1[1] = inc.last_hdr; d->inc1[2] = inc.count;
    kpc_flag_release(&d->flag1, (p.epoch << 2) | KPC_ST_INC);
    if (sh.tile == p.n_tiles - 1) p.carry_out->s1 = inc;
  }

  // ---- phase 4: framing -> one class code per byte ---------------------------------------------------
  KPC_HD static void classify(Shared &sh, const KpcTileParams &p, int tid) {
    const int lo = seg_lo(sh, tid), hi = seg_hi(sh, tid);
    if (lo >= hi) return;
    const uint64_t abs0 = p.abs_base + tile_start(sh);
    KpcS1 st = kpc_s1_combine(sh.tile_in, kpc_s1_combine(sh.gpre[tid / Shared::GS], sh.pre[tid]));
    uint8_t prev = rawb(sh, lo - 1);
    if (FMT == KPC_FMT_FASTQ) {
      uint64_t line = st.count;
      for (int i = lo; i < hi; ++i) {
        const uint8_t b = sh.raw[Shared::px(i)];
        const bool at_start = prev == '\n';
        const uint32_t ph = (uint32_t)line & 3u;
        const bool live = line < p.max_lines;
        uint8_t c = KPC_CLS_BREAK;
        if (at_start && live) {
          // tag.[0] <> '@' || tmp.[0] <> '+' (Files.ml:213); an empty tag / '+' line raises as well
          if ((ph == 0 && b != '@') || (ph == 2 && b != '+')) kpc_atomic_min_u64(p.err_line, line);
          if (p.probe_pos && ph == 0 && abs0 + i >= p.probe_from) kpc_atomic_min_u64(p.probe_pos, abs0 + i);
          if (p.rec_tab && ph == 0) {
            uint64_t r = (line >> 2) - p.rec_base;
            if (r < p.rec_cap) p.rec_tab[r].tag_start = abs0 + i + 1;
          }
        }
        if (b == '\n') {
          if (p.rec_tab && ph == 0 && live) {
            uint64_t r = (line >> 2) - p.rec_base;
            if (r < p.rec_cap) p.rec_tab[r].tag_end = abs0 + i;
          }
          ++line;
        } else if (ph == 1 && live) {
          c = classify_symbol(b);
        }
        sh.cls[Shared::px(i)] = c;
        prev = b;
      }
    } else {
      uint64_t last_nl = st.last_nl, last_hdr = st.last_hdr, nrec = st.count;
      for (int i = lo; i < hi; ++i) {
        const uint8_t b = sh.raw[Shared::px(i)];
        uint8_t c;
        if (b == '\n') {
          c = KPC_CLS_SKIP;
          if (p.rec_tab && (last_hdr >> 1) > last_nl) {  // this line feed terminates a header line
            uint64_t r = nrec - 1 - p.rec_base;
            if (r < p.rec_cap) p.rec_tab[r].tag_end = abs0 + i;
          }
          last_nl = abs0 + i + 1;
        } else if (b == '>' && prev == '\n') {
          uint64_t named = sh.raw[Shared::px(i + 1)] != '\n' ? 1u : 0u;
          last_hdr = ((abs0 + i + 1) << 1) | named;
          c = KPC_CLS_BREAK;
          if (p.probe_pos && abs0 + i >= p.probe_from) kpc_atomic_min_u64(p.probe_pos, abs0 + i);
          if (p.rec_tab) {
            uint64_t r = nrec - p.rec_base;
            if (r < p.rec_cap) p.rec_tab[r].tag_start = abs0 + i + 1;
          }
          ++nrec;
        } else if ((last_hdr >> 1) > last_nl) {
          c = KPC_CLS_BREAK;  // inside a header line
        } else if (!(last_hdr & 1u)) {
          c = KPC_CLS_BREAK;  // before the first header, or in a record whose name is empty (Files.ml:101-106)
        } else {
          c = classify_symbol(b);
        }
        sh.cls[Shared::px(i)] = c;
        prev = b;
      }
    }
  }

  // ---- phase 5a: the tile's k-mer carry summary (last thread), published as stage-2 aggregate --------
  KPC_HD static void ksummary(Shared &sh, const KpcTileParams &p, int tid) {
    if (tid != NT - 1) return;
    KpcKCarry c = kpc_kc_identity();
    if (p.k == 1) c.closed = 1;  // windows of one symbol never cross a boundary
    for (int i = (int)sh.len - 1; i >= 0 && !c.closed; --i) {
      uint8_t x = sh.cls[Shared::px(i)];
      if (x == KPC_CLS_SKIP) continue;
      if (x == KPC_CLS_BREAK) { c.closed = 1; break; }
      c.syms |= (uint64_t)x << (c.n * SB);
      if (++c.n == (uint32_t)(p.k - 1)) c.closed = 1;
    }
    sh.ksum = c;
    KpcTileDesc *d = p.desc + sh.tile;
    d->kagg_syms = c.syms; d->kagg_n = c.n; d->kagg_closed = c.closed;
    kpc_flag_release(&d->flag2, (p.epoch << 2) | KPC_ST_AGG);
  }

  // rolling state: f = forward code, r = reverse-complement code (DNA-ds only), len = valid symbols so far
  struct Roll {
    uint64_t f, r;
    uint32_t len;
  };
  KPC_HD static void push(Roll &w, uint32_t c, int k, uint64_t mask_f) {
    w.f = ((w.f << SB) & mask_f) | c;                                    // KMers.ml:366-367 / 331-332 / 237-238
    if (CONTENT == KPC_CONTENT_DNA_DS) w.r = (w.r >> 2) | ((uint64_t)(3u - c) << (2 * (k - 1)));  // :368
    w.len++;
  }
  KPC_HD static uint64_t key_of(const Roll &w) {
    if (CONTENT == KPC_CONTENT_DNA_DS) return w.f < w.r ? w.f : w.r;  // min hash_f hash_r, KMers.ml:388
    return w.f;
  }
  KPC_HD static uint64_t rec_of(uint64_t line_or_nrec) {
    return FMT == KPC_FMT_FASTQ ? (line_or_nrec >> 2) : (line_or_nrec - 1);
  }
  // insertion rank of a window (see KpcTileParams): monotone in the order the reference visits windows
  KPC_HD static uint64_t rank_of(const KpcTileParams &p, uint64_t abs_pos, uint64_t rec, uint64_t line_start) {
    if (FMT == KPC_FMT_FASTQ)
      return ((p.rank_base + rec * p.rank_mates + p.rank_mate) << 32) | ((abs_pos - line_start) & 0xffffffffull);
    return p.rank_base + abs_pos;
  }

  // ---- phase 5b: every thread rolls over its segment, warmed up from the class codes before it -------
  template <class Sink>
  KPC_HD static void kmers(Shared &sh, const KpcTileParams &p, const Sink &sink, int tid) {
    const int lo = seg_lo(sh, tid), hi = seg_hi(sh, tid);
    if (lo >= hi) return;
    const int k = p.k;
    const uint64_t mask_f = (k * SB >= 64) ? ~0ull : ((1ull << (k * SB)) - 1ull);
    const uint64_t abs0 = p.abs_base + tile_start(sh);
    // record index of the segment: FASTQ from the line number, FASTA from the header count
    KpcS1 st = kpc_s1_combine(sh.tile_in, kpc_s1_combine(sh.gpre[tid / Shared::GS], sh.pre[tid]));
    uint64_t cnt = st.count;
    uint64_t line_start = st.last_nl;  // stream offset of the first byte of the current line
    // warm-up: walk back over the class codes (tile start acts as a break: older symbols are the fix-up's)
    Roll w; w.f = 0; w.r = 0; w.len = 0;
    for (int i = lo - 1; i >= 0 && w.len < (uint32_t)(k - 1); --i) {
      uint8_t x = sh.cls[Shared::px(i)];
      if (x == KPC_CLS_SKIP) continue;
      if (x == KPC_CLS_BREAK) break;
      w.f |= (uint64_t)x << (w.len * SB);
      if (CONTENT == KPC_CONTENT_DNA_DS) w.r |= (uint64_t)(3u - x) << (2 * (k - 1 - (int)w.len));
      w.len++;
    }
    for (int i = lo; i < hi; ++i) {
      uint8_t x = sh.cls[Shared::px(i)];
      // keep the record counter in step with classify(): '\n' (FASTQ) / header start (FASTA)
      if (FMT == KPC_FMT_FASTQ) { if (sh.raw[Shared::px(i)] == '\n') { ++cnt; line_start = abs0 + i + 1; } }
      else if (x == KPC_CLS_BREAK && sh.raw[Shared::px(i)] == '>' && rawb(sh, i - 1) == '\n') ++cnt;
      if (x == KPC_CLS_SKIP) continue;
      if (x == KPC_CLS_BREAK) { w.len = 0; continue; }
      push(w, x, k, mask_f);
      if (w.len >= (uint32_t)k) {
        const uint64_t rec = rec_of(cnt);
        sink.emit(key_of(w), rank_of(p, abs0 + i, rec, line_start), rec);
      }
    }
  }

  // ---- phase 6: thread 0 produces the windows that use symbols from before the tile -------------------
  template <class Sink>
  KPC_HD static void fixup(Shared &sh, const KpcTileParams &p, const Sink &sink, int tid) {
    if (tid != 0) return;
    const int k = p.k;
    KpcTileDesc *d = p.desc + sh.tile;
    // stage-2 look-back: exclusive k-mer carry at the tile start
    KpcKCarry pre = kpc_kc_identity();
    if (k == 1) pre.closed = 1;
    long long j = (long long)sh.tile - 1;
    while (!pre.closed) {
      if (j < 0) {
        KpcKCarry c;
        c.syms = kpc_ld_cg_u64(&p.carry_in->kc.syms);
        c.n = kpc_ld_cg_u32(&p.carry_in->kc.n);
        c.closed = 1;
        pre = kpc_kc_combine(c, pre, k, SB);
        pre.closed = 1;
        break;
      }
      const KpcTileDesc *e = p.desc + j;
      uint32_t f;
      do { f = kpc_flag_acquire(&e->flag2); } while ((f >> 2) != p.epoch);
      KpcKCarry v;
      if ((f & 3u) == KPC_ST_INC) {
        v.syms = kpc_ld_cg_u64(&e->kinc_syms); v.n = kpc_ld_cg_u32(&e->kinc_n); v.closed = 1;
        pre = kpc_kc_combine(v, pre, k, SB);
        pre.closed = 1;
        break;
      }
      v.syms = kpc_ld_cg_u64(&e->kagg_syms); v.n = kpc_ld_cg_u32(&e->kagg_n); v.closed = kpc_ld_cg_u32(&e->kagg_closed);
      pre = kpc_kc_combine(v, pre, k, SB);
      --j;
    }
    // inclusive carry = (carry at tile start) then (this tile's summary)
    KpcKCarry inc = kpc_kc_combine(pre, sh.ksum, k, SB);
    d->kinc_syms = inc.syms; d->kinc_n = inc.n; d->kinc_closed = 1;
    kpc_flag_release(&d->flag2, (p.epoch << 2) | KPC_ST_INC);
    if (sh.tile == p.n_tiles - 1) {
      p.carry_out->kc.syms = inc.syms; p.carry_out->kc.n = inc.n; p.carry_out->kc.closed = 1;
      p.carry_out->last_byte = rawb(sh, (int)sh.len - 1);
    }
    if (pre.n == 0 || k == 1) return;
    // boundary windows: rebuild the rolling state from the carried symbols, then feed tile symbols until
    // k-1 of them have been consumed (later windows lie inside the tile) or a break ends the run
    const uint64_t mask_f = (k * SB >= 64) ? ~0ull : ((1ull << (k * SB)) - 1ull);
    const uint64_t abs0 = p.abs_base + tile_start(sh);
    Roll w; w.f = 0; w.r = 0; w.len = 0;
    for (int s = (int)pre.n - 1; s >= 0; --s) push(w, (uint32_t)((pre.syms >> (s * SB)) & ((1u << SB) - 1u)), k, mask_f);
    // (a line feed in FASTQ and a header in FASTA are breaks, so record and line are those of the tile start)
    const uint64_t rec = rec_of(sh.tile_in.count);
    const uint64_t line_start = sh.tile_in.last_nl;
    int used = 0;
    for (int i = 0; i < (int)sh.len && used < k - 1; ++i) {
      uint8_t x = sh.cls[Shared::px(i)];
      if (x == KPC_CLS_SKIP) continue;
      if (x == KPC_CLS_BREAK) break;
      push(w, x, k, mask_f);
      ++used;
      if (w.len >= (uint32_t)k) sink.emit(key_of(w), rank_of(p, abs0 + i, rec, line_start), rec);
    }
  }
};
