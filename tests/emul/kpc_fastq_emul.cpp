// tests/emul/kpc_fastq_emul.cpp -- TEST INFRASTRUCTURE ONLY (see kpc_rt_emul.cpp, simt_emul.h).
// Runs the partition / count kernels of kpop_b200/csrc/kpc_partition.cuh -- the very source nvcc compiles for sm_100a --
// under the SIMT emulator with small, selectable geometries (KPC_EMUL_FQ=NTxVPT[xGRID]), so that tile, batch, bucket and
// queue boundaries fall everywhere in the test inputs.
#define KPC_SIMT_EMUL 1
#include <cstdio>
#include <cstdlib>

#include "../../include/kpopcount.h"
#include "../../kpop_b200/csrc/kpc_partition.cuh"

namespace {
int g_nt = 0, g_vpt = 0, g_grid = 0;
void geometry() {
  if (g_nt) return;
  g_nt = 32; g_vpt = 1; g_grid = 3;
  const char *e = getenv("KPC_EMUL_FQ");
  if (e) {
    int n = sscanf(e, "%dx%dx%d", &g_nt, &g_vpt, &g_grid);
    if (n < 2) { g_nt = 32; g_vpt = 1; }
    if (n < 3) g_grid = 3;
  }
}
template <class G, bool DS, int KT>
void run_partition(const KpcFqLaunch &L) {
  unsigned grid = (unsigned)g_grid;
  if (grid > L.n_tiles) grid = L.n_tiles;
  if (grid < 1) return;
  simt::launch(grid, G::NT, sizeof(FqSmemT<G>), [&]() { fq_partition_body<G, DS, KT>(L, simt::smem); });
}
template <class G>
void run_geom(const KpcFqLaunch &L) {
  const bool ds = L.content == KPC_CONTENT_DNA_DS;
  if (L.k == 12) { if (ds) run_partition<G, true, 12>(L); else run_partition<G, false, 12>(L); }
  else { if (ds) run_partition<G, true, 0>(L); else run_partition<G, false, 0>(L); }
}
}  // namespace

uint32_t kpc_fq_tile_bytes() {
  geometry();
  return (uint32_t)(16 * g_vpt * g_nt);
}
void kpc_fq_partition(const KpcFqLaunch &L, rt_stream) {
  geometry();
  if (getenv("KPC_EMUL_TRACE")) fprintf(stderr, "emul fq_partition: n=%llu tiles=%u k=%d halo=%d max_lines=%llu\n", (unsigned long long)L.n, L.n_tiles, L.k, L.halo_ok, (unsigned long long)L.max_lines);
  if (!kpc_fq_supported(L.k, L.content) || L.n_slices > (uint32_t)FQ_MAXSLICES || L.n_slices < 1 ||
      (L.n_slices & (L.n_slices - 1)) || (FQ_BUCKET_ENTRIES / L.n_slices) % FQ_CHUNK || FQ_BUCKET_ENTRIES / L.n_slices < 2 * FQ_CHUNK ||
      (1u << L.slice_bits) != L.n_slices || L.lo_bits + L.slice_bits > 2 * L.k)
    throw KpcError(KPC_E_STATE, "internal: fast FASTQ path asked for an unsupported configuration");
  if (g_nt == 32 && g_vpt == 1) run_geom<FqGeom<32, 1>>(L);
  else if (g_nt == 32 && g_vpt == 2) run_geom<FqGeom<32, 2>>(L);
  else if (g_nt == 64 && g_vpt == 1) run_geom<FqGeom<64, 1>>(L);
  else if (g_nt == 64 && g_vpt == 4) run_geom<FqGeom<64, 4>>(L);
  else if (g_nt == 128 && g_vpt == 4) run_geom<FqGeom<128, 4>>(L);
  else throw KpcError(KPC_E_ARG, "KPC_EMUL_FQ: unsupported geometry");
}
void kpc_fq_count(const KpcFqLaunch &L, rt_stream) {
  unsigned grid = 3;
  if (grid > L.n_slices) grid = L.n_slices;
  simt::launch(grid, 64, ((size_t)4 << L.log_bins) + 16, [&]() { fq_count_body<64>(L, simt::smem); });
}
void kpc_fq_timing_enable(bool) {}
void kpc_fq_timing_read(double *a, double *b, unsigned long long *c, unsigned long long *d) { *a = 0; *b = 0; *c = 0; *d = 0; }
