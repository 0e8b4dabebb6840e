// kpopcount_main.cpp -- the `KPopCount` executable: same command line, same output bytes, same exit codes as
// the reference program (bin/KPopCount.ml:105-250 @ a1fda68; option parser BiOCamLib/lib/Tools.ml:541-584,
// 751-766 with unique-prefix matching, Tools.ml:299-409), with the counting done by libkpopcount_gpu.so.
// This is the host side the north star describes, written in C++ because there is no OCaml toolchain in the
// build image; INTEGRATION.md shows the same calls as OCaml `external`s.
//
// Exit codes: 0 success; 1 command-line error or -h (Tools.ml:545, bin/KPopCount.ml:211); 2 anything the
// reference dies of with an uncaught exception (bad k, malformed FASTQ, quotes in a name, unreadable file).
// Text on stderr is not part of the contract and is written afresh.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "../../include/kpopcount.h"

namespace {

struct Input {
  int format;  // KPC_FASTA / KPC_FASTQ_SE / KPC_FASTQ_PE
  std::string file1, file2;
};
struct ArgvError {
  std::string msg;
};
struct ExitNow {
  int code;
};

const char *kNames[] = {"-k", "-K", "--k-mer-size", "--k-mer-length", "-M", "--max-results-size", "-C", "--content",
                        "-f", "--fasta", "-s", "--single-end", "-p", "--paired-end", "-l", "--label",
                        "-L", "--one-spectrum-per-sequence", "-o", "--output", "-v", "--verbose", "-V", "--version",
                        "--markdown", "-h", "--help"};

// Tools.Trie.find_string: an exact name wins; otherwise the argument must be a prefix of exactly one name
std::string resolve_option(const std::string &s) {
  std::string only;
  int n = 0;
  for (const char *name : kNames) {
    if (s == name) return s;
    if (strlen(name) > s.size() && strncmp(name, s.c_str(), s.size()) == 0) { only = name; ++n; }
  }
  return n == 1 ? only : std::string();
}

// OCaml int_of_string: [-+]? (0x|0o|0b|0u)? digits with '_' allowed after the first digit, 63-bit range
bool parse_ocaml_int(const std::string &s, long long &out) {
  size_t i = 0, n = s.size();
  if (!n) return false;
  bool neg = false;
  if (s[i] == '-') { neg = true; ++i; } else if (s[i] == '+') { ++i; }
  int base = 10;
  bool plain = true;
  if (i + 1 < n && s[i] == '0') {
    char c = s[i + 1];
    if (c == 'x' || c == 'X') { base = 16; i += 2; plain = false; }
    else if (c == 'o' || c == 'O') { base = 8; i += 2; plain = false; }
    else if (c == 'b' || c == 'B') { base = 2; i += 2; plain = false; }
    else if (c == 'u' || c == 'U') { i += 2; plain = false; }
  }
  auto digit = [&](char c) {
    int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : 99;
    return d < base ? d : -1;
  };
  if (i >= n || digit(s[i]) < 0) return false;
  unsigned __int128 v = 0, two62 = (unsigned __int128)1 << 62, two63 = (unsigned __int128)1 << 63;
  for (; i < n; ++i) {
    if (s[i] == '_') continue;
    int d = digit(s[i]);
    if (d < 0) return false;
    v = v * base + d;
    if (v >= two63 * 2) return false;
  }
  if (plain) {
    if (neg ? v > two62 : v >= two62) return false;
    out = neg ? -(long long)v : (long long)v;
  } else {
    if (v >= two63) return false;
    long long r = (long long)(v & (two63 - 1));
    if (v >= two62) r -= (long long)two63;
    out = neg ? -r : r;
  }
  return true;
}

// Matrix.Base.strip_external_quotes_and_check (BiOCamLib/lib/Matrix.ml:83-99)
bool strip_quotes(const std::string &s0, std::string &out) {
  size_t l = s0.size();
  if (!l) { out.clear(); return true; }
  if (l == 1 && s0[0] == '"') return false;
  out = s0;
  if (out[0] == '"' && out[l - 1] == '"') out = out.substr(1, l - 2);
  return out.find('"') == std::string::npos;
}

void usage(FILE *o) {
  fputs("This is the KPopCount program (version 18), B200 edition [libkpopcount_gpu]\n"
        " Usage:\n  KPopCount -l <output_vector_label>|-L [OPTIONS]\n"
        " Algorithmic parameters\n"
        "  -k|-K|--k-mer-size|--k-mer-length <k_mer_length>\n"
        "     k-mer length (must be positive, and <= 30 for DNA or <= 12 for protein)  (default='12')\n"
        "  -M|--max-results-size <positive_integer>\n"
        "     maximum number of k-mer hashes kept in memory at any given time; when it is reached at the end of\n"
        "     a sequence the table is printed out and emptied, so hashes can be repeated  (default='16777216')\n"
        " Input/Output\n"
        "  -C|--content 'DNA-ss'|'DNA-single-stranded'|'DNA-ds'|'DNA-double-stranded'|'protein'  (default='DNA-ds')\n"
        "  -f|--fasta <fasta_file_name>                           (can be repeated)\n"
        "  -s|--single-end <fastq_file_name>                      (can be repeated)\n"
        "  -p|--paired-end <fastq_file_name1> <fastq_file_name2>  (can be repeated)\n"
        "  -l|--label <output_vector_label>         one spectrum with this label\n"
        "  -L|--one-spectrum-per-sequence           one spectrum per sequence, labelled with its name\n"
        "  -o|--output <output_file_prefix>         writes <prefix>.KPopSpectra.txt (verbatim if /dev/*)  (default='<stdout>')\n"
        " Miscellaneous\n  -v|--verbose\n  -V|--version\n  -h|--help\n", o);
}

struct Output {
  FILE *f = nullptr;
  std::string *capture = nullptr;  // paired-end first pass: keep the text until the pair count is known
  static int sink(void *u, const char *b, size_t n) {
    Output *o = (Output *)u;
    if (o->capture) { o->capture->append(b, n); return 0; }
    return fwrite(b, 1, n, o->f) == n ? 0 : 1;
  }
};

struct Params {
  int content = KPC_DNA_DS;
  long long k = 12, max_results_size = 16777216;
  std::vector<Input> inputs;
  std::string label, output;
  bool verbose = false;
};

// stream one file (or file pair) through the context; returns a KPC_* code
int feed_files(kpc_ctx *ctx, const Input &in, std::string &err) {
  int rc = kpc_begin(ctx, in.format);  // the label header goes out here, before the file is opened (bin/KPopCount.ml:33-34)
  if (rc) return rc;
  const int mates = in.format == KPC_FASTQ_PE ? 2 : 1;
  int fd[2] = {-1, -1};
  bool done[2] = {false, mates == 1};
  for (int m = 0; m < mates; ++m) {
    const std::string &name = m == 0 ? in.file1 : in.file2;
    fd[m] = open(name.c_str(), O_RDONLY);
    if (fd[m] < 0) {
      err = name + ": " + strerror(errno);  // open_in raises Sys_error: uncaught in the reference
      if (m == 1) close(fd[0]);
      return KPC_E_IO;
    }
  }
  const int slots = kpc_staging_slots(ctx);
  int slot = 0;
  rc = KPC_OK;
  while (rc == KPC_OK && !(done[0] && done[1])) {
    for (int m = 0; m < mates && rc == KPC_OK; ++m) {
      if (done[m]) continue;
      size_t cap = 0;
      char *buf = (char *)kpc_staging(ctx, slot, &cap);
      if (!buf) { rc = KPC_E_NOMEM; break; }
      slot = (slot + 1) % slots;
      size_t got = 0;
      bool eof = false;
      while (got < cap) {
        ssize_t r = read(fd[m], buf + got, cap - got);
        if (r < 0) {
          if (errno == EINTR) continue;
          err = std::string("read: ") + strerror(errno);
          rc = KPC_E_IO;
          break;
        }
        if (r == 0) { eof = true; break; }
        got += (size_t)r;
      }
      if (rc == KPC_OK) rc = kpc_feed(ctx, m, buf, got, eof ? 1 : 0);
      if (eof) done[m] = true;
    }
  }
  for (int m = 0; m < mates; ++m) if (fd[m] >= 0) close(fd[m]);
  if (rc == KPC_OK) rc = kpc_end(ctx);
  return rc;
}

// one complete run; limits[j] < 0: no pair limit on input j.  Returns a KPC_* code; on KPC_E_PE_MISMATCH
// *bad_input / *pairs say which input stopped and how many complete pairs it holds
int run(const Params &P, Output &out, const std::vector<long long> &limits, bool single_pass, size_t *bad_input,
        long long *pairs, std::string &err, bool &ctx_failed_early) {
  kpc_ctx *ctx = nullptr;
  // which GPUs: KPC_DEVICES=0,1,... (several devices share the FASTQ inputs of a dense-table run), or KPC_DEVICE=n.
  // The argv surface stays the reference's (it keeps -t / --threads reserved but commented out, bin/KPopCount.ml:93,187-194).
  std::vector<int> devs;
  if (const char *e = getenv("KPC_DEVICES")) {
    for (const char *q = e; *q;) {
      char *endp = nullptr;
      const long v = strtol(q, &endp, 10);
      if (endp == q) break;
      devs.push_back((int)v);
      q = *endp == ',' ? endp + 1 : endp;
      if (*endp && *endp != ',') break;
    }
  }
  if (devs.empty()) devs.push_back(getenv("KPC_DEVICE") ? atoi(getenv("KPC_DEVICE")) : 0);
  // the functor application of bin/KPopCount.ml:239-249: the k range check fires before the output exists
  int rc = kpc_create(&ctx, (int)(P.k > 1000 ? 1000 : P.k), P.content, P.max_results_size, P.label.c_str(), (int)devs.size(),
                      devs.data());
  if (rc) {
    err = ctx ? kpc_error(ctx) : "out of memory";
    kpc_destroy(ctx);
    ctx_failed_early = true;
    return rc;
  }
  ctx_failed_early = false;
  if (!out.f && !out.capture) {
    if (P.output.empty()) out.f = stdout;
    else {
      out.f = fopen(P.output.c_str(), "wb");
      if (!out.f) { err = P.output + ": " + strerror(errno); kpc_destroy(ctx); return KPC_E_IO; }
    }
  }
  kpc_set_sink(ctx, Output::sink, &out);
  kpc_set_single_pass(ctx, single_pass ? 1 : 0);
  size_t j = 0;
  for (; j < P.inputs.size(); ++j) {
    kpc_set_pair_limit(ctx, limits[j]);
    rc = feed_files(ctx, P.inputs[j], err);
    if (rc) break;
  }
  if (rc == KPC_OK) rc = kpc_finish(ctx);
  if (rc && err.empty()) err = kpc_error(ctx);
  if (rc == KPC_E_PE_MISMATCH) { *pairs = kpc_complete_pairs(ctx); *bad_input = j; }
  if (P.verbose && rc == KPC_OK) fprintf(stderr, "(KPopCount): %llu device kernels launched\n", kpc_kernel_launches(ctx));
  kpc_destroy(ctx);
  return rc;
}

int real_main(int argc, char **argv) {
  Params P;
  bool option_l_or_L = false;
  int i = 1;
  auto error = [&](const std::string &m) { throw ArgvError{m}; };
  auto get_parameter = [&]() -> std::string {
    ++i;
    if (i >= argc) error(std::string("Option '") + argv[i - 1] + "' needs a parameter");
    return argv[i];
  };
  auto get_parameter_int_pos = [&]() -> long long {
    std::string p = get_parameter();
    long long v;
    if (!parse_ocaml_int(p, v)) error(std::string("Option '") + argv[i - 1] + "' needs an integer parameter");
    if (v <= 0) error(std::string("Option '") + argv[i - 1] + "' needs a positive integer parameter");
    return v;
  };
  while (i < argc) {
    const std::string arg = argv[i];
    const std::string opt = resolve_option(arg);
    if (opt.empty()) error("Unknown option '" + arg + "'");
    if (opt == "-k" || opt == "-K" || opt == "--k-mer-size" || opt == "--k-mer-length") P.k = get_parameter_int_pos();
    else if (opt == "-M" || opt == "--max-results-size") P.max_results_size = get_parameter_int_pos();
    else if (opt == "-C" || opt == "--content") {
      const std::string w = get_parameter();
      if (w == "DNA-ss" || w == "DNA-single-stranded") P.content = KPC_DNA_SS;
      else if (w == "DNA-ds" || w == "DNA-double-stranded") P.content = KPC_DNA_DS;
      else if (w == "protein" || w == "prot") P.content = KPC_PROTEIN;
      else {  // Content.Invalid_content is not caught in the reference
        fprintf(stderr, "Fatal error: exception KPopCount.Content.Invalid_content(\"%s\")\n", w.c_str());
        throw ExitNow{2};
      }
    } else if (opt == "-f" || opt == "--fasta") P.inputs.push_back(Input{KPC_FASTA, get_parameter(), ""});
    else if (opt == "-s" || opt == "--single-end") P.inputs.push_back(Input{KPC_FASTQ_SE, get_parameter(), ""});
    else if (opt == "-p" || opt == "--paired-end") {
      const std::string a = get_parameter();
      const std::string b = get_parameter();
      P.inputs.push_back(Input{KPC_FASTQ_PE, a, b});
    } else if (opt == "-l" || opt == "--label") {
      option_l_or_L = true;
      if (!strip_quotes(get_parameter(), P.label)) error("Spectrum labels must not contain quotes");
    } else if (opt == "-L" || opt == "--one-spectrum-per-sequence") option_l_or_L = true;
    else if (opt == "-o" || opt == "--output") {
      const std::string w = get_parameter();  // KMerDB.Spectra.make_filename (lib/KMerDB.ml:26-31)
      P.output = (w.size() >= 5 && w.compare(0, 5, "/dev/") == 0) ? w : w + ".KPopSpectra.txt";
    } else if (opt == "-v" || opt == "--verbose") P.verbose = true;
    else if (opt == "-V" || opt == "--version") { printf("18\n"); fflush(stdout); throw ExitNow{0}; }
    else if (opt == "--markdown") { usage(stderr); throw ExitNow{0}; }
    else if (opt == "-h" || opt == "--help") { usage(stderr); throw ExitNow{1}; }
    ++i;
  }
  if (!option_l_or_L) error("One of options '-l' and '-L' is mandatory");
  if (P.verbose) fprintf(stderr, "This is the KPopCount program (version 18), B200 edition [backend: %s]\n", kpc_backend());
  if (P.inputs.empty()) return 0;  // bin/KPopCount.ml:218
  for (size_t j = 1; j < P.inputs.size(); ++j)
    if ((P.inputs[j].format == KPC_FASTA) != (P.inputs[0].format == KPC_FASTA))
      error("You cannot process FASTA and FASTQ inputs together");

  bool has_pairs = false;
  for (const Input &in : P.inputs) has_pairs |= in.format == KPC_FASTQ_PE;
  Output out;
  std::string captured, err;
  // Regular files are counted mate by mate at full speed and, should a mate file turn out shorter, once more with a pair
  // limit (the text is held back until then).  Pipes and devices can only be read once: pairs are woven on the host.
  bool single_pass = false;
  if (has_pairs)
    for (const Input &in : P.inputs)
      for (const std::string *name : {&in.file1, &in.file2}) {
        struct stat sb;
        if (!name->empty() && stat(name->c_str(), &sb) == 0 && !S_ISREG(sb.st_mode)) single_pass = true;
      }
  if (single_pass) has_pairs = false;  // nothing to hold back or to repeat
  if (has_pairs) out.capture = &captured;
  std::vector<long long> limits(P.inputs.size(), -1);
  bool early = false;
  int rc;
  for (;;) {
    long long pairs = -1;
    size_t bad = 0;
    rc = run(P, out, limits, single_pass, &bad, &pairs, err, early);
    if (rc != KPC_E_PE_MISMATCH || limits[bad] >= 0) break;
    // FASTQ.iter_pe stops at the end of the shorter file (Files.ml:228-247): count again with that many pairs
    limits[bad] = pairs;
    captured.clear();
    err.clear();
  }
  if (has_pairs && !early) {
    FILE *f = stdout;
    if (!P.output.empty()) f = fopen(P.output.c_str(), "wb");
    if (!f) { fprintf(stderr, "Fatal error: exception Sys_error(\"%s: %s\")\n", P.output.c_str(), strerror(errno)); return 2; }
    fwrite(captured.data(), 1, captured.size(), f);
    if (f != stdout) fclose(f); else fflush(f);
  } else if (out.f) {
    fflush(out.f);
    if (out.f != stdout) fclose(out.f);
  }
  if (rc == KPC_OK) return 0;
  fprintf(stderr, "Fatal error: exception Failure(\"(KPopCount): %s\") [code %d]\n", err.c_str(), rc);
  return rc == KPC_E_ARG ? 1 : 2;
}

}  // namespace

int main(int argc, char **argv) {
  try {
    return real_main(argc, argv);
  } catch (const ArgvError &e) {
    usage(stderr);
    fprintf(stderr, "(Tools.Argv.parse): %s\n", e.msg.c_str());
    return 1;
  } catch (const ExitNow &e) {
    return e.code;
  }
}
