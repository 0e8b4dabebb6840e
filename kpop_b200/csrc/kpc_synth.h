// kpc_synth.h -- deterministic synthetic single-end FASTQ of the C3/C5 shape:
//   record i = "@S<i>\n" + 150 bases + "\n+\n" + 150 x 'I' + "\n"        (307 + digits(i) bytes)
//   bases i.i.d. uniform over ACGT, each independently replaced by 'N' with probability 2^-10.
// One function gives any byte of any record, so the host generator (tests, oracle side) and the device
// generator (bench) produce identical streams.  Constants below are frozen: BASELINE-facing numbers depend on them.
#pragma once
#include "kpc_common.h"

#define KPC_SYNTH_READ_LEN 150

KPC_HD uint32_t kpc_synth_digits(uint64_t i) {
  uint32_t d = 1;
  while (i >= 10) { i /= 10; ++d; }
  return d;
}
KPC_HD uint64_t kpc_synth_record_len(uint64_t i) { return 307u + kpc_synth_digits(i); }
// stream offset of record i = 307*i + sum_{j<i} digits(j)
KPC_HD uint64_t kpc_synth_record_offset(uint64_t i) {
  uint64_t off = 307ull * i, start = 0, cnt = 10, d = 1;
  for (;;) {
    if (i < start + cnt) { off += d * (i - start); break; }
    off += d * cnt;
    start += cnt;
    cnt = (d == 1) ? 90 : cnt * 10;
    ++d;
  }
  return off;
}
KPC_HD uint64_t kpc_synth_word(uint64_t seed, uint64_t rec, uint32_t w) {
  return kpc_splitmix64(kpc_splitmix64(seed ^ (0x5851F42D4C957F2Dull * (rec + 1))) + w);
}
KPC_HD uint8_t kpc_synth_base(uint64_t seed, uint64_t rec, uint32_t j) {
  uint64_t nw = kpc_synth_word(seed, rec, 8 + j / 6);
  if (((nw >> (10 * (j % 6))) & 1023u) == 0) return 'N';
  uint64_t bw = kpc_synth_word(seed, rec, j / 32);
  return (uint8_t)("ACGT"[(bw >> (2 * (j % 32))) & 3u]);
}
// byte q (0-based) of record rec whose decimal index has nd digits
KPC_HD uint8_t kpc_synth_byte(uint64_t seed, uint64_t rec, uint32_t nd, uint32_t q) {
  if (q == 0) return '@';
  if (q == 1) return 'S';
  if (q < 2 + nd) {
    uint32_t pos = nd - 1 - (q - 2);
    uint64_t v = rec;
    for (uint32_t t = 0; t < pos; ++t) v /= 10;
    return (uint8_t)('0' + (v % 10));
  }
  q -= 2 + nd;
  if (q == 0) return '\n';
  q -= 1;
  if (q < KPC_SYNTH_READ_LEN) return kpc_synth_base(seed, rec, q);
  q -= KPC_SYNTH_READ_LEN;
  if (q == 0) return '\n';
  if (q == 1) return '+';
  if (q == 2) return '\n';
  q -= 3;
  if (q < KPC_SYNTH_READ_LEN) return 'I';
  return '\n';
}
