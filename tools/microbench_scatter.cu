// tools/microbench_scatter.cu -- what one k-mer append costs on a B200 SM, for the alternatives VERDICT r1 asked to measure
// before touching fq_partition_kernel again.  Every variant takes 16 pseudo-random keys per thread per round (the shape
// of a unit of kpc_partition.cuh: 2 x 512 threads per SM, rounds separated by barriers) and appends each key to the
// shared-memory bucket of its slice; buckets wrap instead of being copied out, so only the append itself is timed.
//
//   atoms      ATOMS.ADD (position) + STS.U16 into one of 512 slice buckets          -- the product kernel's append
//   atoms32    the same with 32 slices                                                -- (is it the slice count or the banks?)
//   ballot5    no atomic: 32 warp-private buckets by a 5-bit digit, ranked with 5 ballots + popc; the counters live in
//              registers (lane d holds the fill of bucket d).  512 slices would need TWO such passes (5 + 4 bits) and an
//              intermediate buffer, so its cost per k-mer is at least twice what is printed
//   match      __match_any_sync on the 9-bit slice, one ATOMS per group of equal slices + STS.U16
//   sts_only   the STS.U16 alone, position from a register counter (what the scatter would cost with free ranking)
//
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbs tools/microbench_scatter.cu && /tmp/mbs
// Prints one JSON line per variant: keys per second, SM cycles per warp-key (one key in each of 32 lanes).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int NT = 512, ROUNDS = 64, W = 16, SLICES = 512, STRIDE = 28;  // bucket stride in words, as in the product

extern __shared__ __align__(16) uint8_t dyn[];
constexpr int DYN_BYTES = 4 * SLICES + SLICES * STRIDE * 4;  // fill + buckets (stride 28 words, as in the product)

__device__ __forceinline__ uint32_t next_key(uint32_t &x) {
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  return x;
}

template <int NSL>
__global__ void __launch_bounds__(NT, 2) k_atoms(uint32_t *out, int rounds) {
  uint32_t *fill = reinterpret_cast<uint32_t *>(dyn);
  uint16_t *bucket = reinterpret_cast<uint16_t *>(dyn + 4 * SLICES);
  for (int i = threadIdx.x; i < SLICES; i += NT) fill[i] = 0;
  __syncthreads();
  uint32_t x = 0x9E3779B9u * (blockIdx.x * NT + threadIdx.x + 1);
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t key = next_key(x);
      const uint32_t sl = (key >> 8) & (NSL - 1);
      const uint32_t pos = atomicAdd(&fill[sl], 1u);
      bucket[sl * (STRIDE * 2) + (pos & 31u)] = (uint16_t)key;
    }
    __syncthreads();
  }
  if (x == 0x12345u) out[0] = bucket[threadIdx.x] + fill[threadIdx.x & (SLICES - 1)];
}

__global__ void __launch_bounds__(NT, 2) k_keys_only(uint32_t *out, int rounds) {  // the key generator alone (to subtract)
  uint32_t x = 0x9E3779B9u * (blockIdx.x * NT + threadIdx.x + 1), acc = 0;
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int j = 0; j < W; ++j) acc += (next_key(x) >> 8) & (SLICES - 1);
    __syncthreads();
  }
  if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(NT, 2) k_sts_only(uint32_t *out, int rounds) {
  uint16_t *bucket = reinterpret_cast<uint16_t *>(dyn + 4 * SLICES);
  uint32_t x = 0x9E3779B9u * (blockIdx.x * NT + threadIdx.x + 1), pos = threadIdx.x;
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t key = next_key(x);
      const uint32_t sl = (key >> 8) & (SLICES - 1);
      bucket[sl * (STRIDE * 2) + (pos++ & 31u)] = (uint16_t)key;
    }
    __syncthreads();
  }
  if (x == 0x12345u) out[0] = bucket[threadIdx.x];
}

__global__ void __launch_bounds__(NT, 2) k_match(uint32_t *out, int rounds) {
  uint32_t *fill = reinterpret_cast<uint32_t *>(dyn);
  uint16_t *bucket = reinterpret_cast<uint16_t *>(dyn + 4 * SLICES);
  for (int i = threadIdx.x; i < SLICES; i += NT) fill[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t x = 0x9E3779B9u * (blockIdx.x * NT + threadIdx.x + 1);
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t key = next_key(x);
      const uint32_t sl = (key >> 8) & (SLICES - 1);
      const unsigned m = __match_any_sync(0xffffffffu, sl);
      const int leader = __ffs(m) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&fill[sl], (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      const uint32_t pos = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
      bucket[sl * (STRIDE * 2) + (pos & 31u)] = (uint16_t)key;
    }
    __syncthreads();
  }
  if (x == 0x12345u) out[0] = bucket[threadIdx.x] + fill[threadIdx.x & (SLICES - 1)];
}

// 32 warp-private buckets of 16 entries (+ 2 entries of padding per bucket against bank conflicts between buckets)
__global__ void __launch_bounds__(NT, 2) k_ballot5(uint32_t *out, int rounds) {
  constexpr int CAPW = 16, ROW = CAPW + 2;
  __shared__ uint16_t wb[NT / 32][32 * ROW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = 0x9E3779B9u * (blockIdx.x * NT + threadIdx.x + 1);
  uint32_t counter = 0;  // fill of bucket `lane` of this warp
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t key = next_key(x);
      const uint32_t d = (key >> 8) & 31u;
      unsigned peers = 0xffffffffu, mine = 0xffffffffu;  // lanes with my digit / lanes whose digit is my lane number
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
        mine &= ((lane >> b) & 1) ? bal : ~bal;
      }
      const uint32_t base = __shfl_sync(0xffffffffu, counter, (int)d);
      const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      wb[w][d * ROW + (pos % CAPW)] = (uint16_t)key;
      counter += (uint32_t)__popc(mine);
    }
    __syncthreads();
  }
  if (x == 0x12345u) out[0] = wb[w][lane] + counter;
}

template <class K>
static void run(const char *name, K kern, int sms, double clock_hz) {
  uint32_t *out;
  cudaMalloc(&out, 64);
  const int grid = sms * 2;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_BYTES);
  kern<<<grid, NT, DYN_BYTES>>>(out, 4);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    kern<<<grid, NT, DYN_BYTES>>>(out, ROUNDS * 16);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double keys = (double)grid * NT * W * ROUNDS * 16;
  const double warp_keys_per_sm = keys / 32.0 / sms;
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"keys_per_s\": %.4g, \"sm_cycles_per_warp_key\": %.2f, \"error\": \"%s\"}\n", name, best,
         keys / (best * 1e-3), best * 1e-3 * clock_hz / warp_keys_per_sm, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double hz = p.clockRate * 1e3;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f, \"shape\": \"2 CTAs x 512 threads per SM, 16 keys per thread per round, barrier per round\"}\n",
         p.name, p.multiProcessorCount, hz / 1e6);
  run("keys_only (generator + barrier, to subtract)", k_keys_only, p.multiProcessorCount, hz);
  run("atoms (512 slices: ATOMS.ADD + STS.U16)", k_atoms<512>, p.multiProcessorCount, hz);
  run("atoms32 (32 slices)", k_atoms<32>, p.multiProcessorCount, hz);
  run("sts_only (512 slices, position from a register)", k_sts_only, p.multiProcessorCount, hz);
  run("match (match.any on the slice + one ATOMS per group + STS.U16)", k_match, p.multiProcessorCount, hz);
  run("ballot5 (ONE 5-bit pass into 32 warp-private buckets, no atomic; 512 slices need two passes)", k_ballot5, p.multiProcessorCount, hz);
  return 0;
}
