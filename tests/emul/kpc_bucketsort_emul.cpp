// tests/emul/kpc_bucketsort_emul.cpp -- TEST INFRASTRUCTURE ONLY (see simt_emul.h).
// The per-bucket finalize kernels of kpop_b200/csrc/kpc_bucketsort.cuh under the SIMT emulator.
#define KPC_SIMT_EMUL 1
#define KPC_BS_SMALL_CFG 3   // tiny limits: the CTA-wide kernel and the fall-back to the hash table get exercised
#define KPC_BS_HEAVY_CFG 12
#include "../../kpop_b200/csrc/kpc_bucketsort.cuh"

void kpc_k_bucket_finalize(const KpcBucketFinalize &F, rt_stream) {
  unsigned grid = (F.nb + 63) / 64;
  if (grid > 3) grid = 3;
  if (grid < 1) grid = 1;
  simt::launch(grid, 64, 0, [&]() { kpc_bucket_finalize_small_body<64>(F); });
  simt::launch(2, 64, sizeof(KpcBsHeavySmem), [&]() { kpc_bucket_finalize_heavy_body<64>(F, simt::smem); });
}
