// tests/emul/device_feed_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// KPopCount-like driver that hands whole inputs to kpc_feed_device (the benchmark's entry point: consecutive launches
// of one buffer, each reading the 16 bytes before it) through the C ABI of the emulation build.
//   device_feed_driver <k> <DNA-ds|DNA-ss> <label> <fasta|single-end> file...
// KPC_DRIVER_BATCH=1: one spectrum per file (labels <label>0, <label>1, ...) from ONE context: kpc_reset_label between them.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kpopcount.h"

static int sink(void *, const char *b, size_t n) { return fwrite(b, 1, n, stdout) == n ? 0 : 1; }

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage\n"); return 1; }
  const int k = atoi(argv[1]);
  const int content = !strcmp(argv[2], "DNA-ss") ? KPC_DNA_SS : KPC_DNA_DS;
  const int fmt = !strcmp(argv[4], "fasta") ? KPC_FASTA : KPC_FASTQ_SE;
  kpc_ctx *ctx = nullptr;
  int dev = 0;
  if (kpc_create(&ctx, k, content, 16777216, argv[3], 1, &dev) != KPC_OK) { fprintf(stderr, "create: %s\n", kpc_error(ctx)); return 2; }
  kpc_set_sink(ctx, sink, nullptr);
  const bool batch = getenv("KPC_DRIVER_BATCH") != nullptr;
  for (int a = 5; a < argc; ++a) {
    if (batch) {
      if (a > 5 && kpc_finish(ctx) != KPC_OK) { fflush(stdout); fprintf(stderr, "finish: %s\n", kpc_error(ctx)); return 2; }
      const std::string lab = std::string(argv[3]) + std::to_string(a - 5);
      if (kpc_reset_label(ctx, lab.c_str()) != KPC_OK) { fprintf(stderr, "reset_label: %s\n", kpc_error(ctx)); return 2; }
    }
    FILE *f = fopen(argv[a], "rb");
    if (!f) { perror(argv[a]); return 2; }
    std::vector<char> raw;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + n);
    fclose(f);
    void *mem = nullptr;
    if (posix_memalign(&mem, 16, raw.size() + 64)) return 2;
    memset(mem, 0xAB, raw.size() + 64);
    memcpy(mem, raw.data(), raw.size());
    int rc = kpc_begin(ctx, fmt);
    if (rc == KPC_OK) rc = kpc_feed_device(ctx, 0, mem, raw.size(), 1);
    if (rc == KPC_OK) rc = kpc_end(ctx);
    free(mem);
    if (rc != KPC_OK) { fflush(stdout); fprintf(stderr, "error (code %d): %s\n", rc, kpc_error(ctx)); return 2; }
  }
  if (kpc_finish(ctx) != KPC_OK) { fflush(stdout); fprintf(stderr, "finish: %s\n", kpc_error(ctx)); return 2; }
  kpc_destroy(ctx);
  return 0;
}
