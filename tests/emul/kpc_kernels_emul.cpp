// tests/emul/kpc_kernels_emul.cpp -- TEST INFRASTRUCTURE ONLY (see kpc_rt_emul.cpp).
// Runs the phase functions of kpc_tile.cuh -- the very source the CUDA kernel is made of -- thread by thread
// on the CPU, tile after tile, and gives plain-loop versions of the auxiliary kernels of kpc_kernels.cu.
// The tile geometry is tiny and selectable (KPC_EMUL_TILE=NTxSEG) so that tile, segment and scan-group
// boundaries fall everywhere in the test inputs.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/kpopcount.h"
#include "../../kpop_b200/csrc/kpc_kernels.h"
#include "../../kpop_b200/csrc/kpc_fastq.h"
#include "../../kpop_b200/csrc/kpc_synth.h"

namespace {

int g_nt = 0, g_seg = 0;
void tile_geometry() {
  if (g_nt) return;
  g_nt = 8; g_seg = 4;
  const char *e = getenv("KPC_EMUL_TILE");
  if (e) sscanf(e, "%dx%d", &g_nt, &g_seg);
}

template <int NT, int SEG, int FMT, int CONTENT, class Sink>
void run_tiles(const KpcTileParams &p, const Sink &sink) {
  typedef KpcTileMachine<NT, SEG, FMT, CONTENT> M;
  typename M::Shared *shp = new typename M::Shared();
  typename M::Shared &sh = *shp;
  for (;;) {
    for (int t = 0; t < NT; ++t) M::claim(sh, p, t);
    if (sh.tile >= p.n_tiles) break;
    for (int t = 0; t < NT; ++t) M::load(sh, p, t);
    for (int t = 0; t < NT; ++t) M::census(sh, p, t);
    for (int t = 0; t < NT; ++t) M::scan_groups(sh, t);
    for (int t = 0; t < NT; ++t) M::lookback1(sh, p, t);
    for (int t = 0; t < NT; ++t) M::classify(sh, p, t);
    for (int t = 0; t < NT; ++t) M::ksummary(sh, p, t);
    // kmers and fixup run without a barrier in between on the device: any interleaving must work
    for (int t = NT - 1; t >= 0; --t) { M::kmers(sh, p, sink, t); M::fixup(sh, p, sink, t); }
  }
  delete shp;
}
template <int NT, int SEG, int FMT, int CONTENT>
void run_s(const KpcTileLaunch &L) {
  switch (L.sink) {
    case KPC_SINK_DENSE: run_tiles<NT, SEG, FMT, CONTENT>(L.p, L.dense); break;
    case KPC_SINK_HASH: run_tiles<NT, SEG, FMT, CONTENT>(L.p, L.hash); break;
    case KPC_SINK_TUPLE: run_tiles<NT, SEG, FMT, CONTENT>(L.p, L.tuple); break;
    case KPC_SINK_BCOUNT: run_tiles<NT, SEG, FMT, CONTENT>(L.p, L.bcount); break;
    case KPC_SINK_BSCATTER: run_tiles<NT, SEG, FMT, CONTENT>(L.p, L.bscatter); break;
    default: run_tiles<NT, SEG, FMT, CONTENT>(L.p, KpcNullSink()); break;
  }
}
template <int NT, int SEG, int FMT>
void run_c(const KpcTileLaunch &L) {
  switch (L.content) {
    case KPC_CONTENT_DNA_SS: run_s<NT, SEG, FMT, KPC_CONTENT_DNA_SS>(L); break;
    case KPC_CONTENT_DNA_DS: run_s<NT, SEG, FMT, KPC_CONTENT_DNA_DS>(L); break;
    default: run_s<NT, SEG, FMT, KPC_CONTENT_PROTEIN>(L); break;
  }
}
template <int NT, int SEG>
void run_f(const KpcTileLaunch &L) {
  if (L.fmt == KPC_FMT_FASTQ) run_c<NT, SEG, KPC_FMT_FASTQ>(L);
  else run_c<NT, SEG, KPC_FMT_FASTA>(L);
}

}  // namespace

uint32_t kpc_k_tile_bytes() { tile_geometry(); return (uint32_t)(g_nt * g_seg); }

void kpc_k_tiles(const KpcTileLaunch &L, rt_stream) {
  tile_geometry();
  if (g_nt == 1 && g_seg == 16) run_f<1, 16>(L);
  else if (g_nt == 4 && g_seg == 4) run_f<4, 4>(L);
  else if (g_nt == 8 && g_seg == 4) run_f<8, 4>(L);
  else if (g_nt == 64 && g_seg == 2) run_f<64, 2>(L);
  else if (g_nt == 64 && g_seg == 16) run_f<64, 16>(L);
  else if (g_nt == 256 && g_seg == 64) run_f<256, 64>(L);
  else throw KpcError(KPC_E_ARG, "KPC_EMUL_TILE: unsupported geometry");
}

// (the fast FASTQ pipeline is emulated by kpc_fastq_emul.cpp)

void kpc_k_count_newlines(const uint8_t *d, uint64_t n, unsigned long long *out, rt_stream) {
  unsigned long long c = 0;
  for (uint64_t i = 0; i < n; ++i) c += d[i] == '\n';
  *out += c;
}
void kpc_k_dense_fold(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream) {
  for (uint64_t i = 0; i < nbins; ++i)
    if (lo[i] >= 0x80000000u) { hi[i] += lo[i]; lo[i] = 0; }
}
void kpc_k_add_u64(unsigned long long *dst, const unsigned long long *src, uint64_t n, rt_stream) {
  for (uint64_t i = 0; i < n; ++i) dst[i] += src[i];
}
void kpc_k_dense_promote(uint32_t *lo, unsigned long long *hi, uint64_t nbins, rt_stream) {
  for (uint64_t i = 0; i < nbins; ++i) { hi[i] += lo[i]; lo[i] = 0; }
}
void kpc_k_dense_max(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *out, rt_stream) {
  unsigned long long m = *out;
  for (uint64_t i = 0; i < nbins; ++i) m = std::max<unsigned long long>(m, (unsigned long long)lo[i] + (hi ? hi[i] : 0));
  *out = m;
}
size_t kpc_k_scan_scratch_bytes(uint64_t) { return 64; }
void kpc_k_dense_extract(const uint32_t *lo, const unsigned long long *hi, uint64_t nbins, unsigned long long *keys,
                         unsigned long long *counts, unsigned long long *n_out, void *, rt_stream) {
  unsigned long long n = 0;
  for (uint64_t i = 0; i < nbins; ++i) {
    unsigned long long v = (unsigned long long)lo[i] + (hi ? hi[i] : 0);
    if (v) { keys[n] = i; counts[n] = v; ++n; }
  }
  *n_out = n;
}
void kpc_k_format(const unsigned long long *keys, const unsigned long long *counts, uint64_t n, int hex_width,
                  char *out, unsigned long long *out_len, void *, rt_stream) {
  size_t o = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (counts[i]) o += (size_t)sprintf(out + o, "%0*llx\t%llu\n", hex_width, keys[i], counts[i]);
  *out_len = o;
}
void kpc_k_bucket_scatter_staged(const KpcPair *stage, const unsigned long long *n, const KpcBucketScatterSink &sink, rt_stream) {
  for (unsigned long long i = 0; i < *n; ++i) sink.emit(stage[i].key, stage[i].rank, 0);
}
void kpc_k_bucket_offsets(const uint32_t *hist, uint32_t nb, uint32_t *offsets, void *, rt_stream) {
  uint32_t acc = 0;
  for (uint32_t i = 0; i < nb; ++i) { offsets[i] = acc; acc += hist[i]; }
  offsets[nb] = acc;
}
void kpc_k_hash_clear(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks, uint64_t cap, rt_stream) {
  for (uint64_t i = 0; i < cap; ++i) { keys[i] = ~0ull; counts[i] = 0; ranks[i] = ~0ull; }
}
void kpc_k_hash_rehash(const unsigned long long *okeys, const unsigned long long *ocounts,
                       const unsigned long long *oranks, uint64_t ocap, KpcHashSink nw, rt_stream) {
  for (uint64_t i = 0; i < ocap; ++i)
    if (okeys[i] != ~0ull) nw.insert(okeys[i], oranks[i], ocounts[i], true);
}
void kpc_k_hash_extract(const unsigned long long *keys, const unsigned long long *counts,
                        const unsigned long long *ranks, uint64_t cap, unsigned long long *okeys,
                        unsigned long long *ocounts, unsigned long long *oranks, unsigned long long *n_out, void *, rt_stream) {
  unsigned long long n = 0;
  for (uint64_t i = 0; i < cap; ++i)
    if (keys[i] != ~0ull && (long long)counts[i] > 0) { okeys[n] = keys[i]; ocounts[n] = counts[i]; oranks[n] = ranks[i]; ++n; }
  *n_out = n;
}
size_t kpc_k_order_scratch_bytes(uint64_t) { return 64; }
void kpc_k_order_entries(unsigned long long *keys, unsigned long long *counts, unsigned long long *ranks,
                         uint32_t *recs, uint64_t n, uint64_t bmask, const unsigned long long *bmask_per_rec,
                         uint32_t rec_off, void *, rt_stream) {
  std::vector<uint64_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  auto bm = [&](uint64_t i) { return (recs && bmask_per_rec) ? bmask_per_rec[recs[i] - rec_off] : bmask; };
  std::sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
    if (recs && recs[a] != recs[b]) return recs[a] < recs[b];
    uint64_t ba = keys[a] & bm(a), bb = keys[b] & bm(b);
    if (ba != bb) return ba < bb;
    return ranks[a] > ranks[b];
  });
  std::vector<unsigned long long> k2(n), c2(n), r2(n);
  std::vector<uint32_t> e2(n);
  for (uint64_t i = 0; i < n; ++i) { k2[i] = keys[idx[i]]; c2[i] = counts[idx[i]]; r2[i] = ranks[idx[i]]; if (recs) e2[i] = recs[idx[i]]; }
  for (uint64_t i = 0; i < n; ++i) { keys[i] = k2[i]; counts[i] = c2[i]; ranks[i] = r2[i]; if (recs) recs[i] = e2[i]; }
}
void kpc_k_tuple_bounds(const uint32_t *trecs, uint64_t n, const uint32_t *erec, uint64_t n_entries, uint32_t g_lo,
                        uint32_t g_hi, unsigned long long *bounds, unsigned long long *per_rec, rt_stream) {
  bounds[0] = (unsigned long long)(std::lower_bound(trecs, trecs + n, g_hi) - trecs);
  bounds[1] = (unsigned long long)(std::lower_bound(erec, erec + n_entries, g_hi) - erec);
  for (uint32_t r = g_lo; r < g_hi; ++r) per_rec[r - g_lo] = 0;
  for (uint64_t i = 0; i < n_entries; ++i)
    if (erec[i] >= g_lo && erec[i] < g_hi) per_rec[erec[i] - g_lo]++;
}
void kpc_k_tuple_reduce(unsigned long long *keys, unsigned long long *ranks, uint32_t *recs, uint64_t n,
                        unsigned long long *okeys, unsigned long long *ocounts, unsigned long long *oranks,
                        uint32_t *orecs, unsigned long long *n_out, void *, rt_stream) {
  std::vector<uint64_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
    if (recs[a] != recs[b]) return recs[a] < recs[b];
    return keys[a] < keys[b];
  });
  std::vector<unsigned long long> k2(n), r2(n);
  std::vector<uint32_t> e2(n);
  for (uint64_t i = 0; i < n; ++i) { k2[i] = keys[idx[i]]; r2[i] = ranks[idx[i]]; e2[i] = recs[idx[i]]; }
  unsigned long long m = 0;
  for (uint64_t i = 0; i < n; ++i) {
    keys[i] = k2[i]; ranks[i] = r2[i]; recs[i] = e2[i];
    if (i == 0 || k2[i] != k2[i - 1] || e2[i] != e2[i - 1]) { okeys[m] = k2[i]; orecs[m] = e2[i]; ocounts[m] = 0; oranks[m] = ~0ull; ++m; }
    ocounts[m - 1] += 1;
    oranks[m - 1] = std::min(oranks[m - 1], r2[i]);
  }
  *n_out = m;
}
void kpc_k_rec_counts(const uint32_t *recs, uint64_t n, unsigned long long *cnt, uint64_t n_recs, rt_stream) {
  for (uint64_t r = 0; r < n_recs; ++r) cnt[r] = 0;
  for (uint64_t i = 0; i < n; ++i) if (recs[i] < n_recs) cnt[recs[i]]++;
}
void kpc_k_synth_fastq(uint8_t *out, uint64_t first_record, uint64_t n_records, uint64_t seed, rt_stream) {
  const uint64_t base = kpc_synth_record_offset(first_record);
  for (uint64_t r = 0; r < n_records; ++r) {
    const uint64_t rec = first_record + r;
    const uint32_t nd = kpc_synth_digits(rec), len = 307u + nd;
    uint8_t *o = out + (kpc_synth_record_offset(rec) - base);
    for (uint32_t q = 0; q < len; ++q) o[q] = kpc_synth_byte(seed, rec, nd, q);
  }
}
