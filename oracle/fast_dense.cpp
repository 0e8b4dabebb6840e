// oracle/fast_dense.cpp -- TEST INFRASTRUCTURE ONLY.
// A fast multi-threaded CPU counter for the *dense DNA-ds FASTQ* case, used as the checker for parity runs that
// are too large for the faithful oracle (kpopcount_oracle.cpp), and a closed-form count of the valid windows of
// the synthetic stream.  It is itself validated against the faithful oracle on small inputs
// (tests/test_oracle_fastdense.py).  Semantics restated from the reference:
//   strict 4-line records, only line 2 is sequence, EOF inside a record drops it (Files.ml:201-221);
//   non-ACGT breaks a window (Sequences.ml:41-67, KMers.ml:369-383); key = min(fwd, rc) (KMers.ml:388).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../kpop_b200/csrc/kpc_synth.h"

namespace {
inline int code(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}
void count_line(const uint8_t *s, size_t l, int k, uint32_t *table) {
  const uint64_t mask_f = (1ull << (2 * k)) - 1;
  uint64_t f = 0, r = 0;
  int len = 0;
  for (size_t i = 0; i < l; ++i) {
    int c = code(s[i]);
    if (c < 0) { len = 0; continue; }
    f = ((f << 2) & mask_f) | (uint64_t)c;
    r = (r >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
    if (++len >= k) __atomic_fetch_add(table + (f < r ? f : r), 1u, __ATOMIC_RELAXED);
  }
}
}  // namespace

extern "C" {

// table must hold 4^k zero-initialised u32 (k <= 14).  Returns the number of complete records.
uint64_t fd_count_fastq_dense(const uint8_t *data, uint64_t n, int k, uint32_t *table, int nthreads) {
  // line starts (sequential memchr scan), then records are independent
  std::vector<uint64_t> starts;
  starts.push_back(0);
  for (const uint8_t *p = data, *e = data + n; p < e;) {
    const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
    if (!q) break;
    starts.push_back((uint64_t)(q - data) + 1);
    p = q + 1;
  }
  // number of lines input_line would return
  uint64_t nlines = starts.size() - 1;
  if (n > 0 && data[n - 1] != '\n') nlines += 1;
  else if (!starts.empty() && starts.back() == n) starts.pop_back();
  const uint64_t nrec = nlines / 4;
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      for (uint64_t r = (uint64_t)t; r < nrec; r += (uint64_t)nthreads) {
        const uint64_t a = starts[4 * r + 1];
        uint64_t b = (4 * r + 2 < starts.size()) ? starts[4 * r + 2] - 1 : n;
        if (b > n) b = n;
        count_line(data + a, (size_t)(b - a), k, table);
      }
    });
  for (auto &x : th) x.join();
  return nrec;
}

// number of fully valid k-mer windows in records [first, first+n) of the synthetic stream (kpc_synth.h)
uint64_t fd_synth_valid_windows(uint64_t first, uint64_t n, uint64_t seed, int k, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<uint64_t> part((size_t)nthreads, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      uint64_t tot = 0;
      for (uint64_t r = first + (uint64_t)t; r < first + n; r += (uint64_t)nthreads) {
        int run = 0;
        for (uint32_t j = 0; j < KPC_SYNTH_READ_LEN; ++j) {
          if (j % 6 == 0 || true) {
            const uint64_t nw = kpc_synth_word(seed, r, 8 + j / 6);
            if (((nw >> (10 * (j % 6))) & 1023u) == 0) { run = 0; continue; }
          }
          if (++run >= k) ++tot;
        }
      }
      part[(size_t)t] = tot;
    });
  for (auto &x : th) x.join();
  uint64_t tot = 0;
  for (uint64_t v : part) tot += v;
  return tot;
}

}  // extern "C"
