// tools/microbench.cu -- measures the scatter/increment primitives a k-mer counter can be built from on this GPU:
//   red      RED.ADD.u32 to uniformly random bins of an L2-resident table (what the dense k=12 path does per k-mer)
//   atom     the same with the old value returned (needed for sub-word counters with overflow detection)
//   smem     ATOMS.ADD to random bins of a per-CTA shared-memory table
//   smemrmw  plain LDS / STS read-modify-write to random shared-memory bins (no atomicity; issue-rate bound)
//   copy     128-bit streaming read of a large buffer (the input side of the roofline)
// Prints one JSON line per configuration.  Not part of the product; numbers are recorded in profiles/.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <int MODE>  // 0 red, 1 atom
__global__ void k_global(uint32_t *table, uint32_t mask, int iters, uint32_t *sinkv) {
  uint64_t s = mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
  uint32_t acc = 0;
  for (int i = 0; i < iters; i += 2) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t a = (uint32_t)(s >> 33) & mask, b = (uint32_t)(s >> 9) & mask;
    if (MODE == 0) { atomicAdd(table + a, 1u); atomicAdd(table + b, 1u); }
    else { acc += atomicAdd(table + a, 1u); acc += atomicAdd(table + b, 1u); }
  }
  if (MODE == 1 && acc == 0xdeadbeef) *sinkv = acc;
}
// keys sorted inside each warp instruction so that the 32 lanes touch `groups` distinct sectors
__global__ void k_global_clustered(uint32_t *table, uint32_t mask, int iters, int lanes_per_sector) {
  uint64_t s = mix((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / lanes_per_sector);
  const uint32_t sub = threadIdx.x % lanes_per_sector;
  for (int i = 0; i < iters; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t a = ((uint32_t)(s >> 33) & mask & ~7u) | (sub & 7u);
    atomicAdd(table + a, 1u);
  }
}
template <int MODE>  // 0 atomics, 1 plain read-modify-write
__global__ void k_smem(int bins_mask, int iters, uint32_t *out) {
  extern __shared__ uint32_t h[];
  for (int i = threadIdx.x; i <= bins_mask; i += blockDim.x) h[i] = 0;
  __syncthreads();
  uint64_t s = mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
  for (int i = 0; i < iters; i += 2) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t a = (uint32_t)(s >> 33) & bins_mask, b = (uint32_t)(s >> 9) & bins_mask;
    if (MODE == 0) { atomicAdd(h + a, 1u); atomicAdd(h + b, 1u); }
    else { h[a] += 1u; h[b] += 1u; }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}
__global__ void k_read(const uint4 *src, uint64_t nvec, uint32_t *out) {
  uint32_t acc = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 x;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src + i));
    acc += x.x ^ x.y ^ x.z ^ x.w;
  }
  if (acc == 0x12345678) *out = acc;
}

template <class F>
static float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
  uint32_t *table, *misc;
  CK(cudaMalloc(&table, (size_t)1 << 28));
  CK(cudaMalloc(&misc, 1 << 20));
  CK(cudaMemset(table, 0, (size_t)1 << 28));
  const int iters = 512;
  for (int cps : {2, 4, 8}) {
    for (int logbins : {24, 23, 20, 16, 10}) {
      const int grid = sms * cps;
      const double ops = (double)grid * 256 * iters;
      float ms = time_ms([&] { k_global<0><<<grid, 256>>>(table, (1u << logbins) - 1, iters, misc); });
      printf("{\"bench\": \"red\", \"ctas_per_sm\": %d, \"log2_bins\": %d, \"gops\": %.2f}\n", cps, logbins, ops / ms / 1e6);
      ms = time_ms([&] { k_global<1><<<grid, 256>>>(table, (1u << logbins) - 1, iters, misc); });
      printf("{\"bench\": \"atom\", \"ctas_per_sm\": %d, \"log2_bins\": %d, \"gops\": %.2f}\n", cps, logbins, ops / ms / 1e6);
      fflush(stdout);
    }
  }
  for (int lps : {1, 2, 4, 8, 32}) {
    const int grid = sms * 8;
    const double ops = (double)grid * 256 * iters;
    float ms = time_ms([&] { k_global_clustered<<<grid, 256>>>(table, (1u << 24) - 1, iters, lps); });
    printf("{\"bench\": \"red_clustered\", \"lanes_per_sector\": %d, \"gops\": %.2f}\n", lps, ops / ms / 1e6);
  }
  for (int logbins : {15, 13, 10}) {
    for (int threads : {256, 1024}) {
      const int smem = (1 << logbins) * 4;
      CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int cps = logbins >= 15 ? 1 : 4;
      const int grid = sms * cps;
      const double ops = (double)grid * threads * 2048;
      float ms = time_ms([&] { k_smem<0><<<grid, threads, smem>>>((1 << logbins) - 1, 2048, misc); });
      printf("{\"bench\": \"smem_atomic\", \"log2_bins\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"gops\": %.2f}\n", logbins, threads, cps, ops / ms / 1e6);
      ms = time_ms([&] { k_smem<1><<<grid, threads, smem>>>((1 << logbins) - 1, 2048, misc); });
      printf("{\"bench\": \"smem_rmw\", \"log2_bins\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"gops\": %.2f}\n", logbins, threads, cps, ops / ms / 1e6);
    }
  }
  {
    const size_t bytes = (size_t)4 << 30;
    uint4 *buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
    for (int cps : {4, 8, 16}) {
      float ms = time_ms([&] { k_read<<<sms * cps, 256>>>(buf, bytes / 16, misc); });
      printf("{\"bench\": \"read\", \"ctas_per_sm\": %d, \"gbs\": %.1f}\n", cps, bytes / ms / 1e6);
    }
    cudaFree(buf);
  }
  return 0;
}
