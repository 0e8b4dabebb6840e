// tests/emul/unit_checks.cpp -- TEST INFRASTRUCTURE ONLY.
// Exhaustive checks of the SIMD-in-register helpers of kpc_partition.cuh against their byte-at-a-time definitions.
#define KPC_SIMT_EMUL 1
#include <cstdio>
#include <cstdlib>

#include "../../kpop_b200/csrc/kpc_partition.cuh"

int main() {
  int bad = 0;
  // fq_classify4: every byte value in every lane, next to arbitrary neighbours
  uint64_t rng = 12345;
  for (int rep = 0; rep < 64; ++rep) {
    for (int lane = 0; lane < 4; ++lane) {
      for (int b = 0; b < 256; ++b) {
        rng = kpc_splitmix64(rng);
        uint32_t wd = (uint32_t)rng;
        if (rep == 0) wd = 0;
        if (rep == 1) wd = 0xFFFFFFFFu;
        wd = (wd & ~(0xFFu << (8 * lane))) | ((uint32_t)b << (8 * lane));
        uint32_t nz;
        const uint32_t x = fq_classify4(wd, nz);
        for (int l = 0; l < 4; ++l) {
          const uint8_t by = (uint8_t)(wd >> (8 * l));
          const uint8_t want = kpc_classify_dna(by);
          const bool inv = (nz >> (8 * l)) & 0x80u;
          if ((nz >> (8 * l)) & 0x7Fu) { ++bad; }
          if (inv != (want == KPC_CLS_BREAK)) { if (bad < 10) fprintf(stderr, "validity of byte %02x: got %d\n", by, (int)inv); ++bad; }
          if (!inv) {
            uint32_t c = (x >> (8 * l)) & 0xFFu;  // A C T G -> 0 1 2 3
            c ^= c >> 1;
            if (c != want) { if (bad < 10) fprintf(stderr, "code of byte %02x: got %u want %u\n", by, c, want); ++bad; }
          }
        }
      }
    }
  }
  // fq_nl_mask16 against a byte loop
  for (int rep = 0; rep < 200000; ++rep) {
    uint4 v;
    uint8_t bytes[16];
    for (int i = 0; i < 16; ++i) {
      rng = kpc_splitmix64(rng);
      const int r = (int)(rng & 7);
      bytes[i] = r == 0 ? '\n' : r == 1 ? 0x8A : r == 2 ? 0x0B : r == 3 ? 0x00 : (uint8_t)(rng >> 8);
    }
    memcpy(&v, bytes, 16);
    uint32_t want = 0;
    for (int i = 0; i < 16; ++i) want |= (uint32_t)(bytes[i] == '\n') << i;
    if (fq_nl_mask16(v) != want) { if (bad < 10) fprintf(stderr, "nl_mask16 mismatch\n"); ++bad; }
  }
  if (bad) { fprintf(stderr, "unit_checks: %d failures\n", bad); return 1; }
  printf("unit_checks ok\n");
  return 0;
}
