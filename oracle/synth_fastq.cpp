// oracle/synth_fastq.cpp -- TEST INFRASTRUCTURE ONLY.
// Host generator of the synthetic benchmark reads (same definition as the device generator: both include
// kpop_b200/csrc/kpc_synth.h).  Usage: synth_fastq <first_record> <n_records> <seed> > out.fq
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../kpop_b200/csrc/kpc_synth.h"

int main(int argc, char **argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s <first_record> <n_records> <seed>\n", argv[0]); return 1; }
  const uint64_t first = strtoull(argv[1], nullptr, 10), n = strtoull(argv[2], nullptr, 10), seed = strtoull(argv[3], nullptr, 10);
  std::vector<char> buf;
  buf.reserve(1 << 22);
  for (uint64_t r = first; r < first + n; ++r) {
    char rec[400];
    int len = snprintf(rec, sizeof rec, "@S%llu\n", (unsigned long long)r);
    uint64_t bw[5], nw[25];
    for (uint32_t w = 0; w < 5; ++w) bw[w] = kpc_synth_word(seed, r, w);
    for (uint32_t w = 0; w < 25; ++w) nw[w] = kpc_synth_word(seed, r, 8 + w);
    for (uint32_t j = 0; j < KPC_SYNTH_READ_LEN; ++j) {
      char c = "ACGT"[(bw[j / 32] >> (2 * (j % 32))) & 3u];
      if (((nw[j / 6] >> (10 * (j % 6))) & 1023u) == 0) c = 'N';
      rec[len++] = c;
    }
    rec[len++] = '\n'; rec[len++] = '+'; rec[len++] = '\n';
    for (uint32_t j = 0; j < KPC_SYNTH_READ_LEN; ++j) rec[len++] = 'I';
    rec[len++] = '\n';
    buf.insert(buf.end(), rec, rec + len);
    if (buf.size() > (1 << 22) - 512) { fwrite(buf.data(), 1, buf.size(), stdout); buf.clear(); }
  }
  fwrite(buf.data(), 1, buf.size(), stdout);
  return 0;
}
