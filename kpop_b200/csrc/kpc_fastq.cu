// kpc_fastq.cu -- fast FASTQ -> dense 4^k table pipeline for sm_100a (see kpc_fastq.h for the outline).
//
// Reference semantics restated here (paths relative to the KPop tree):
//   Files.FASTQ.iter_se        BiOCamLib/lib/Files.ml:201-221   4-line records, '@' / '+' checks, only line 2 is sequence
//   Sequences.Lint.dnaize      BiOCamLib/lib/Sequences.ml:41-67  [ACGTacgt] are bases, every other byte breaks k-mers
//   DNAHash*.iteri / iterc     BiOCamLib/lib/KMers.ml:319-349, 357-389   first base most significant, key = min f rc
//   IntHashFrequencies.add     BiOCamLib/lib/KMers.ml:107-111   count[key] += 1
//
// The kernels themselves live in kpc_partition.cuh (one source for nvcc and for the test-only SIMT emulator); this
// file instantiates them for the production geometry and launches them.  DESIGN.md section 5 has the measurements.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include <string>
#include <vector>

#include "../../include/kpopcount.h"
#include "kpc_fastq.h"

#define FQ_CUDA_CHECK(x)                                                                                      \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      throw KpcError(KPC_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +        \
                                     std::to_string(__LINE__) + ")");                                         \
  } while (0)

// dynamic shared memory of both kernels; at file scope so that its (constant) shared-space address can be named in PTX
extern __shared__ __align__(16) uint8_t fq_smem_raw[];

#include "kpc_partition.cuh"

namespace {

// production geometry (tools/gpu_define_variants.sh rebuilds with other values to measure them)
#ifndef FQ_PROD_NT
#define FQ_PROD_NT 512
#endif
#ifndef FQ_PROD_VPT
#define FQ_PROD_VPT 4
#endif
#ifndef FQ_PROD_CTAS
#define FQ_PROD_CTAS 2
#endif
typedef FqGeom<FQ_PROD_NT, FQ_PROD_VPT> FqProd;  // 512 threads x 64 bytes: 32 KiB tiles, two CTAs per SM

template <bool DS, int KT>
__global__ void __launch_bounds__(FqProd::NT, FQ_PROD_CTAS) fq_partition_kernel(const KpcFqLaunch p) {
  fq_partition_body<FqProd, DS, KT>(p, fq_smem_raw);
}

constexpr int FQ_CNT_NT = 1024;
__global__ void __launch_bounds__(FQ_CNT_NT, 1) fq_count_kernel(const KpcFqLaunch p) {
  fq_count_body<FQ_CNT_NT>(p, fq_smem_raw);
}

}  // namespace

uint32_t kpc_fq_tile_bytes() { return FqProd::TB; }

static cudaStream_t fq_cs(rt_stream s) { return (cudaStream_t)rt_stream_native(s); }

// ---- optional timing: one event triple per launch (before partition, between, after count) -------------------------
namespace {
struct FqTiming {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // 3 per launch
  size_t used = 0;              // launches recorded since the last read
  unsigned long long bytes = 0;
} g_fqt;
cudaEvent_t fq_timing_event(size_t i) {
  while (g_fqt.ev.size() <= i) {
    cudaEvent_t e;
    FQ_CUDA_CHECK(cudaEventCreate(&e));
    g_fqt.ev.push_back(e);
  }
  return g_fqt.ev[i];
}
}  // namespace
void kpc_fq_timing_enable(bool on) { g_fqt.on = on; g_fqt.used = 0; g_fqt.bytes = 0; }
void kpc_fq_timing_read(double *partition_ms, double *count_ms, unsigned long long *launches, unsigned long long *bytes) {
  double a = 0, b = 0;
  for (size_t i = 0; i < g_fqt.used; ++i) {
    float t = 0;
    FQ_CUDA_CHECK(cudaEventSynchronize(g_fqt.ev[3 * i + 2]));
    FQ_CUDA_CHECK(cudaEventElapsedTime(&t, g_fqt.ev[3 * i], g_fqt.ev[3 * i + 1])); a += t;
    FQ_CUDA_CHECK(cudaEventElapsedTime(&t, g_fqt.ev[3 * i + 1], g_fqt.ev[3 * i + 2])); b += t;
  }
  *partition_ms = a; *count_ms = b; *launches = g_fqt.used; *bytes = g_fqt.bytes;
  g_fqt.used = 0; g_fqt.bytes = 0;
}

template <bool DS, int KT>
static void launch_partition(const KpcFqLaunch &L, rt_stream s) {
  auto kern = fq_partition_kernel<DS, KT>;
  static int blocks_per_sm_dev[64] = {0};  // the shared-memory attribute is per device: one slot per device
  int &blocks_per_sm = blocks_per_sm_dev[rt_current_device() & 63];
  if (!blocks_per_sm) {
    FQ_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FqSmemT<FqProd>)));
    FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, FqProd::NT, sizeof(FqSmemT<FqProd>)));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    if (const char *e = getenv("KPC_FQ_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < blocks_per_sm) blocks_per_sm = v; }  // experiments
  }
  long long grid = (long long)rt_sm_count() * blocks_per_sm;
  if (grid > (long long)L.n_tiles) grid = L.n_tiles;
  if (grid < 1) return;
  kern<<<(unsigned)grid, FqProd::NT, sizeof(FqSmemT<FqProd>), fq_cs(s)>>>(L);
  FQ_CUDA_CHECK(cudaGetLastError());
}

void kpc_fq_partition(const KpcFqLaunch &L, rt_stream s) {
  if (!kpc_fq_supported(L.k, L.content) || L.n_slices > (uint32_t)FQ_MAXSLICES || L.n_slices < 1 ||
      (L.n_slices & (L.n_slices - 1)) || (FQ_BUCKET_ENTRIES / L.n_slices) % FQ_CHUNK || FQ_BUCKET_ENTRIES / L.n_slices < 2 * FQ_CHUNK ||
      (1u << L.slice_bits) != L.n_slices || L.lo_bits + L.slice_bits > 2 * L.k)
    throw KpcError(KPC_E_STATE, "internal: fast FASTQ path asked for an unsupported configuration");
  const bool ds = L.content == KPC_CONTENT_DNA_DS;
  if (g_fqt.on) FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used), fq_cs(s)));
  if (L.k == 12) { if (ds) launch_partition<true, 12>(L, s); else launch_partition<false, 12>(L, s); }
  else { if (ds) launch_partition<true, 0>(L, s); else launch_partition<false, 0>(L, s); }
  if (g_fqt.on) FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used + 1), fq_cs(s)));
#ifdef FQ_PHASE_CLOCKS
  {  // experiment: print and clear the per-phase cycle counts of this launch
    unsigned long long h[16], z[16] = {0};
    FQ_CUDA_CHECK(cudaStreamSynchronize(fq_cs(s)));
    FQ_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_fq_phase, sizeof h));
    FQ_CUDA_CHECK(cudaMemcpyToSymbol(g_fq_phase, z, sizeof z));
    unsigned long long tot = 0;
    for (int i = 0; i < 15; ++i) tot += h[i];
    fprintf(stderr, "fq phases (%% of %llu cycles, n = %llu B):", tot, (unsigned long long)L.n);
    for (int i = 0; i < 15; ++i) fprintf(stderr, " %d:%.1f", i, 100.0 * (double)h[i] / (double)(tot ? tot : 1));
    fprintf(stderr, "\n");
  }
#endif
}

void kpc_fq_count(const KpcFqLaunch &L, rt_stream s) {
  static bool attr_dev[64] = {false};
  bool &attr = attr_dev[rt_current_device() & 63];
  if (!attr) {
    FQ_CUDA_CHECK(cudaFuncSetAttribute(fq_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 << 15) + 16));
    attr = true;
  }
  long long grid = rt_sm_count();
  if (grid > (long long)L.n_slices) grid = L.n_slices;
  fq_count_kernel<<<(unsigned)grid, FQ_CNT_NT, ((size_t)4 << L.log_bins) + 16, fq_cs(s)>>>(L);
  FQ_CUDA_CHECK(cudaGetLastError());
  if (g_fqt.on) {
    FQ_CUDA_CHECK(cudaEventRecord(fq_timing_event(3 * g_fqt.used + 2), fq_cs(s)));
    g_fqt.used++;
    g_fqt.bytes += L.n;
  }
}
