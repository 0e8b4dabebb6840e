// oracle/kpopcount_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may build or execute this file.  Nothing under kpop_b200/ links, loads or calls it.
//
// Single-threaded C++ restatement of the reference KPopCount hot path (PaoloRibeca/KPop
// @ a1fda68, BiOCamLib @ c6a17c9).  The reference is pure OCaml and there is no OCaml
// toolchain in this image, so the reference cannot be compiled here;
// this file follows it function by function:
//
//   bin/KPopCount.ml:20-64      KMerCounter.compute (header, per-record iterc, dump/clear rule)
//   bin/KPopCount.ml:66-95      Content, Parameters (defaults k=12, M=16777216, DNA-ds)
//   bin/KPopCount.ml:105-250    argv table, -l/-L rule, FASTA/FASTQ mixing rule, functor choice
//   BiOCamLib/lib/KMers.ml:99-113   IntHashFrequencies over IntHashtbl
//   BiOCamLib/lib/KMers.ml:142-258  ProteinHash
//   BiOCamLib/lib/KMers.ml:261-311  DNABaseHash (k<=30, to_hex width)
//   BiOCamLib/lib/KMers.ml:313-350  DNAHashSingleStranded.iteri/iterc
//   BiOCamLib/lib/KMers.ml:351-390  DNAHashDoubleStrandedLexicographic.iteri/iterc
//   BiOCamLib/lib/Files.ml:96-122   FASTA.iter
//   BiOCamLib/lib/Files.ml:201-250  FASTQ.iter_se / iter_pe
//   BiOCamLib/lib/Files.ml:350-368  ReadsIterate.iter
//   BiOCamLib/lib/Sequences.ml:41-67,87-151  Lint.dnaize / Lint.proteinize
//   BiOCamLib/lib/Better.ml:700-705,741      IntHash (identity hash), IntHashtbl
//   BiOCamLib/lib/Matrix.ml:83-99   strip_external_quotes_and_check
//   lib/KMerDB.ml:26-31             Spectra.make_filename
//   BiOCamLib/lib/Tools.ml:299-409  Trie (unique-prefix option matching)
//   BiOCamLib/lib/Tools.ml:541-584,751-766  Argv.error / get_parameter* / parse loop
//
// Arithmetic that lives OUTSIDE /root/reference: the emitted order is the iteration order of
// OCaml's Stdlib.Hashtbl (functorial interface, Hashtbl.Make; compiler version unpinned by
// the reference: dune-project asks for "ocaml", README.md:63 for >= 4.12).  Its published
// algorithm (stdlib/hashtbl.ml, 4.12 .. 5.x) is restated in class OcamlIntHashtbl below:
//   create n      -> bucket array of size power_2_above 16 n
//   key_index     -> (hash key) land (Array.length data - 1)        [hash = identity here]
//   add           -> cons at bucket head; size+1; if size > 2*buckets then resize
//   resize        -> double; walk old buckets 0..n-1 head-to-tail, APPEND to new bucket tail
//   iter          -> buckets 0..n-1, each head-to-tail
//   clear         -> if size > 0: size <- 0; fill with Empty (array length kept)
//
// PARITY UNPINNED: the reference ships no golden vector, known-answer test or fixture for
// this path and cannot be run here, so this oracle is pinned only against
// the survey's independently derived digests (tests/golden/) and hand-checkable KATs.
//
// Build: see oracle/Makefile.  Usage: identical to the reference KPopCount command line.

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

typedef long long ml_int;  // OCaml 63-bit int: every value on this path fits in 60 bits

// ---------------------------------------------------------------------------------------------
// Exceptions that are uncaught in the reference => OCaml runtime prints "Fatal error" and exits 2
// ---------------------------------------------------------------------------------------------
struct OcamlFailure : std::runtime_error {
  explicit OcamlFailure(const std::string &m) : std::runtime_error(m) {}
};
// Tools.Argv.error: usage + message on stderr, exit 1 (Tools.ml:541-545)
struct ArgvError : std::runtime_error {
  explicit ArgvError(const std::string &m) : std::runtime_error(m) {}
};
struct ExitNow {
  int code;
};

// ---------------------------------------------------------------------------------------------
// OCaml Stdlib.Hashtbl.Make(IntHash) restated (see header).  Values are the boxed counters of
// IntHashFrequencies (KMers.ml:99-113).
// ---------------------------------------------------------------------------------------------
class OcamlIntHashtbl {
 public:
  struct Cell {
    ml_int key;
    ml_int count;
    int64_t next;  // index into cells_, -1 = Empty
  };

  explicit OcamlIntHashtbl(ml_int n) {
    // power_2_above 16 n (Sys.max_array_length is 2^54-1 on 64 bit: never reached)
    uint64_t s = 16;
    while ((ml_int)s < n) s *= 2;
    data_.assign(s, -1);
    size_ = 0;
  }
  ml_int length() const { return size_; }
  uint64_t buckets() const { return data_.size(); }

  // IntHashFrequencies.add (KMers.ml:107-111): find_opt then either bump or H.add
  void add(ml_int key, ml_int occs) {
    uint64_t idx = (uint64_t)key & (data_.size() - 1);
    for (int64_t c = data_[idx]; c >= 0; c = cells_[c].next)
      if (cells_[c].key - key == 0) {  // IntHash.equal (Better.ml:703)
        cells_[c].count += occs;
        return;
      }
    // Hashtbl.add: new cell at the bucket head
    if (data_[idx] < 0) touched_.push_back(idx);
    cells_.push_back(Cell{key, occs, data_[idx]});
    data_[idx] = (int64_t)cells_.size() - 1;
    size_ += 1;
    if ((uint64_t)size_ > (data_.size() << 1)) resize();
  }

  // Hashtbl.iter: buckets in index order, each head to tail.  Only non-empty buckets are
  // visited here (the list of touched buckets is sorted first), which yields exactly the
  // sequence a scan over all buckets would.
  template <class F>
  void iter(F f) {
    std::sort(touched_.begin(), touched_.end());
    for (uint64_t idx : touched_)
      for (int64_t c = data_[idx]; c >= 0; c = cells_[c].next) f(cells_[c].key, cells_[c].count);
  }

  // Hashtbl.clear: size <- 0, fill with Empty, bucket array length unchanged
  void clear() {
    if (size_ > 0) {
      size_ = 0;
      for (uint64_t idx : touched_) data_[idx] = -1;
      touched_.clear();
      cells_.clear();
    }
  }

 private:
  // Hashtbl.resize + insert_all_buckets (inplace): order inside every new bucket keeps the
  // relative order the cells had in their old bucket.
  void resize() {
    uint64_t osize = data_.size(), nsize = osize * 2;
    std::vector<int64_t> ndata(nsize, -1), ntail(nsize, -1);
    std::sort(touched_.begin(), touched_.end());
    std::vector<uint64_t> ntouched;
    for (uint64_t i : touched_) {
      int64_t c = data_[i];
      while (c >= 0) {
        int64_t next = cells_[c].next;
        uint64_t nidx = (uint64_t)cells_[c].key & (nsize - 1);
        cells_[c].next = -1;
        if (ntail[nidx] < 0) {
          ndata[nidx] = c;
          ntouched.push_back(nidx);
        } else {
          cells_[ntail[nidx]].next = c;
        }
        ntail[nidx] = c;
        c = next;
      }
    }
    data_.swap(ndata);
    touched_.swap(ntouched);
  }

  std::vector<int64_t> data_;
  std::vector<Cell> cells_;
  std::vector<uint64_t> touched_;
  ml_int size_;
};

// ---------------------------------------------------------------------------------------------
// Sequences.Lint (Sequences.ml:41-67, 87-151), keep_lowercase=false keep_dashes=false
// ---------------------------------------------------------------------------------------------
std::string dnaize(const std::string &s) {
  std::string b(s);
  for (size_t i = 0; i < b.size(); ++i) {
    switch (b[i]) {
      case 'A': case 'a': b[i] = 'A'; break;
      case 'C': case 'c': b[i] = 'C'; break;
      case 'G': case 'g': b[i] = 'G'; break;
      case 'T': case 't': b[i] = 'T'; break;
      default: b[i] = 'N';
    }
  }
  return b;
}
const char kProteinAlphabet[] = "ACDEFGHIKLMNOPQRSTUVWY";  // KMers.ml:150
std::string proteinize(const std::string &s) {
  std::string b(s);
  for (size_t i = 0; i < b.size(); ++i) {
    char c = b[i];
    if (c == '*') continue;
    char u = (c >= 'a' && c <= 'z') ? (char)(c - 32) : c;
    if (u >= 'A' && u <= 'Z' && strchr(kProteinAlphabet, u))
      b[i] = u;
    else
      b[i] = 'X';
  }
  return b;
}

// ---------------------------------------------------------------------------------------------
// KMers: hash families.  Each exposes k, to_hex and iterc (KMers.ml:127-140 IntHash_t).
// ---------------------------------------------------------------------------------------------
struct HashFamily {
  int k = 0;
  int hex_width = 0;
  virtual ~HashFamily() {}
  virtual void iterc(OcamlIntHashtbl &hf, const std::string &s) const = 0;
  std::string to_hex(ml_int h) const {
    char buf[32];
    snprintf(buf, sizeof buf, "%0*llx", hex_width, (unsigned long long)h);
    return buf;
  }
};

inline int dna_code(char c) {  // KMers.ml:272-277
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

// KMers.ml:313-350
struct DNAHashSingleStranded : HashFamily {
  explicit DNAHashSingleStranded(int n) {
    if (n > 30)  // KMers.ml:264-267
      throw OcamlFailure("(BiOCamLib.KMers.DNABaseHash.k): Invalid argument (k must be <= 30, found " +
                         std::to_string(n) + ")");
    k = n;
    hex_width = (k * 2 + 3) / 4;
  }
  void iterc(OcamlIntHashtbl &hf, const std::string &s) const override {
    const ml_int l = (ml_int)s.size();
    const ml_int mask = ((ml_int)1 << (2 * k)) - 4;
    ml_int start = 0, hash = 0, pos = 0;
    for (;;) {  // let rec shift start hash pos
      if (pos - start >= k) hf.add(hash, 1);
      ml_int incr_pos = pos + 1;
      if (incr_pos > l) break;
      int c = dna_code(s[pos]);
      if (c >= 0) {
        hash = ((hash << 2) & mask) | c;
        pos = incr_pos;
      } else {
        if (incr_pos + k <= l) {
          start = incr_pos; hash = 0; pos = incr_pos;
        } else {
          break;
        }
      }
    }
  }
};

// KMers.ml:351-390
struct DNAHashDoubleStrandedLexicographic : HashFamily {
  explicit DNAHashDoubleStrandedLexicographic(int n) {
    if (n > 30)
      throw OcamlFailure("(BiOCamLib.KMers.DNABaseHash.k): Invalid argument (k must be <= 30, found " +
                         std::to_string(n) + ")");
    k = n;
    hex_width = (k * 2 + 3) / 4;
  }
  void iterc(OcamlIntHashtbl &hf, const std::string &s) const override {
    const ml_int l = (ml_int)s.size();
    const ml_int mask_f = ((ml_int)1 << (2 * k)) - 4;
    const ml_int mask_r = ((ml_int)1 << (2 * k - 2)) - 1;
    ml_int start = 0, hf_ = 0, hr_ = 0, pos = 0;
    for (;;) {
      if (pos - start >= k) hf.add(std::min(hf_, hr_), 1);  // KMers.ml:388
      ml_int incr_pos = pos + 1;
      if (incr_pos > l) break;
      int c = dna_code(s[pos]);
      if (c >= 0) {
        hf_ = ((hf_ << 2) & mask_f) | c;
        hr_ = ((hr_ >> 2) & mask_r) | ((ml_int)(3 - c) << (2 * (k - 1)));
        pos = incr_pos;
      } else {
        if (incr_pos + k <= l) {
          start = incr_pos; hf_ = 0; hr_ = 0; pos = incr_pos;
        } else {
          break;
        }
      }
    }
  }
};

// KMers.ml:142-258
struct ProteinHash : HashFamily {
  explicit ProteinHash(int n) {
    if (n > 12)  // KMers.ml:145-148
      throw OcamlFailure("(BiOCamLib.KMers.ProteinHash.k): Invalid argument (k must be <= 12, found " +
                         std::to_string(n) + ")");
    k = n;
    hex_width = (k * 5 + 3) / 4;
  }
  static int encode_char(char c) {  // KMers.ml:153-176
    char u = (c >= 'a' && c <= 'z') ? (char)(c - 32) : c;
    if (!(u >= 'A' && u <= 'Z')) return -1;
    const char *p = strchr(kProteinAlphabet, u);
    return p ? (int)(p - kProteinAlphabet) : -1;
  }
  void iterc(OcamlIntHashtbl &hf, const std::string &s) const override {
    const ml_int l = (ml_int)s.size();
    const ml_int mask = ((ml_int)1 << (5 * k)) - 32;
    ml_int start = 0, hash = 0, pos = 0;
    for (;;) {
      if (pos - start >= k) hf.add(hash, 1);
      ml_int incr_pos = pos + 1;
      if (incr_pos > l) break;
      int e = encode_char(s[pos]);
      if (e >= 0) {
        hash = ((hash << 5) & mask) | e;
        pos = incr_pos;
      } else {
        if (incr_pos + k <= l) {
          start = incr_pos; hash = 0; pos = incr_pos;
        } else {
          break;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// in_channel + input_line (OCaml stdlib): split at '\n', terminator removed, '\r' kept,
// final unterminated line returned, End_of_file when nothing is pending.
// ---------------------------------------------------------------------------------------------
class InChannel {
 public:
  explicit InChannel(const std::string &name) {
    f_ = fopen(name.c_str(), "rb");
    if (!f_)  // open_in raises Sys_error: uncaught => exit 2
      throw OcamlFailure("Sys_error(\"" + name + ": " + strerror(errno) + "\")");
    buf_.resize(1 << 20);
  }
  ~InChannel() {
    if (f_) fclose(f_);
  }
  bool input_line(std::string &line) {  // false == End_of_file
    line.clear();
    bool any = false;
    for (;;) {
      if (pos_ == len_) {
        len_ = fread(&buf_[0], 1, buf_.size(), f_);
        pos_ = 0;
        if (len_ == 0) return any;
      }
      any = true;
      const char *p = &buf_[pos_];
      const char *nl = (const char *)memchr(p, '\n', len_ - pos_);
      if (nl) {
        line.append(p, nl - p);
        pos_ += (size_t)(nl - p) + 1;
        return true;
      }
      line.append(p, len_ - pos_);
      pos_ = len_;
    }
  }

 private:
  FILE *f_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
};

// Files.Type (Files.ml:315-323): only the three kinds reachable from KPopCount's argv
struct Input {
  enum Kind { FASTA, SingleEndFASTQ, PairedEndFASTQ } kind;
  std::string file1, file2;
};

// Matrix.Base.strip_external_quotes_and_check (BiOCamLib/lib/Matrix.ml:83-99)
struct QuotesInName : std::runtime_error {
  explicit QuotesInName(const std::string &s) : std::runtime_error(s) {}
};
std::string strip_external_quotes_and_check(const std::string &s0) {
  size_t l = s0.size();
  if (l == 0) return "";
  if (l == 1 && s0 == "\"") throw QuotesInName(s0);
  std::string s = s0;
  if (s[0] == '"' && s[l - 1] == '"') s = s.substr(1, l - 2);
  if (s.find('"') != std::string::npos) throw QuotesInName(s);
  return s;
}

// KMerDB.Spectra.make_filename (lib/KMerDB.ml:26-31)
std::string make_filename(const std::string &w) {
  if (w.size() >= 5 && w.compare(0, 5, "/dev/") == 0) return w;
  return w + ".KPopSpectra.txt";
}

// ---------------------------------------------------------------------------------------------
// KMerCounter.compute (bin/KPopCount.ml:26-63) with Files.ReadsIterate.iter inlined
// ---------------------------------------------------------------------------------------------
struct Counter {
  const HashFamily &kih;
  bool protein;
  ml_int max_results_size;
  std::string label;
  FILE *output;
  OcamlIntHashtbl res;
  unsigned long long kmers_added = 0;

  Counter(const HashFamily &h, bool prot, ml_int M, const std::string &lab, FILE *out)
      : kih(h), protein(prot), max_results_size(M), label(lab), output(out), res(M) {}

  std::string lint(const std::string &s) const { return protein ? proteinize(s) : dnaize(s); }

  void dump() {
    res.iter([&](ml_int k, ml_int f) {
      fputs(kih.to_hex(k).c_str(), output);
      fprintf(output, "\t%lld\n", f);
      kmers_added += (unsigned long long)f;
    });
  }

  // the closure passed to ReadsIterate.iter (bin/KPopCount.ml:37-54)
  void on_read(const std::string &tag, const std::string &seq) {
    kih.iterc(res, seq);
    if (label.empty() || res.length() >= max_results_size) {
      if (label.empty()) {
        std::string t = strip_external_quotes_and_check(tag);  // may raise: uncaught => exit 2
        fprintf(output, "\t%s\n", t.c_str());
      }
      dump();
      res.clear();
    }
  }

  // Files.FASTA.iter (Files.ml:96-122)
  void fasta(const std::string &filename) {
    InChannel file(filename);
    std::string current, seq, line;
    auto process_current = [&]() {
      if (!current.empty()) on_read(current, seq);
      seq.clear();
    };
    while (file.input_line(line)) {
      if (!line.empty()) {
        if (line[0] == '>') {
          process_current();
          current = line.substr(1);
        } else {
          seq += lint(line);
        }
      }
    }
    process_current();
  }

  static void check_fastq(const std::string &tag, const std::string &tmp, const char *fn, ml_int read,
                          const std::string &files) {
    // tag.[0] / tmp.[0] on an empty string raise Invalid_argument "index out of bounds"
    if (tag.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
    if (tag[0] == '@' && tmp.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
    if (tag[0] != '@' || tmp[0] != '+')
      throw OcamlFailure(std::string("(") + fn + "): On line " + std::to_string(read) + ": Malformed FASTQ file" +
                         files);
  }

  // Files.FASTQ.iter_se (Files.ml:201-221)
  void fastq_se(const std::string &file) {
    InChannel input(file);
    ml_int read = 0;
    std::string tag, seq, tmp, qua;
    for (;;) {
      if (!input.input_line(tag)) break;
      if (!input.input_line(seq)) break;
      if (!input.input_line(tmp)) break;
      if (!input.input_line(qua)) break;
      read += 4;
      check_fastq(tag, tmp, "BiOCamLib.Files.FASTQ.iter_se", read, " '" + file + "'");
      on_read(tag.substr(1), lint(seq));
    }
  }

  // Files.FASTQ.iter_pe (Files.ml:222-250) + ReadsIterate.iter (Files.ml:363-368)
  void fastq_pe(const std::string &file1, const std::string &file2) {
    InChannel input1(file1);
    InChannel input2(file2);
    ml_int read = 0;
    std::string tag1, seq1, tmp1, qua1, tag2, seq2, tmp2, qua2;
    for (;;) {
      if (!input1.input_line(tag1)) break;
      if (!input1.input_line(seq1)) break;
      if (!input1.input_line(tmp1)) break;
      if (!input1.input_line(qua1)) break;
      if (!input2.input_line(tag2)) break;
      if (!input2.input_line(seq2)) break;
      if (!input2.input_line(tmp2)) break;
      if (!input2.input_line(qua2)) break;
      read += 8;
      // tag1.[0] <> '@' || tmp1.[0] <> '+' || tag2.[0] <> '@' || tmp2.[0] <> '+'  (left to right)
      const std::string files = "(s) '" + file1 + "' and/or '" + file2 + "'";
      if (tag1.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
      bool bad = tag1[0] != '@';
      if (!bad) {
        if (tmp1.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
        bad = tmp1[0] != '+';
      }
      if (!bad) {
        if (tag2.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
        bad = tag2[0] != '@';
      }
      if (!bad) {
        if (tmp2.empty()) throw OcamlFailure("Invalid_argument(\"index out of bounds\")");
        bad = tmp2[0] != '+';
      }
      if (bad)
        throw OcamlFailure("(BiOCamLib.Files.FASTQ.iter_pe): On line " + std::to_string(read) +
                           ": Malformed FASTQ file" + files);
      std::string l1 = lint(seq1), l2 = lint(seq2);  // both linted before the callbacks run
      on_read(tag1.substr(1), l1);
      on_read(tag2.substr(1), l2);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Tools.Trie.find_string (Tools.ml:299-409) over the option names: exact match wins (Unique or
// Contained), otherwise a unique proper prefix resolves to its completion, otherwise "".
// ---------------------------------------------------------------------------------------------
std::string trie_find_string(const std::vector<std::string> &names, const std::string &s) {
  std::vector<std::string> ext;
  bool exact = false;
  for (const std::string &n : names) {
    if (n == s) exact = true;
    else if (n.size() > s.size() && n.compare(0, s.size(), s) == 0) ext.push_back(n);
  }
  if (exact) return s;
  if (ext.size() == 1) return ext[0];
  return "";
}

// OCaml int_of_string: [-+]? (0[xX] hex | 0[oO] oct | 0[bB] bin | 0u dec | dec), '_' allowed
// after the first digit; failure on anything else or on overflow of the 63-bit int.
bool ocaml_int_of_string(const std::string &s, ml_int &out) {
  size_t i = 0, n = s.size();
  if (n == 0) return false;
  bool neg = false;
  if (s[i] == '-') { neg = true; ++i; }
  else if (s[i] == '+') { ++i; }
  int base = 10;
  bool unsigned_dec = false;
  if (i + 1 < n && s[i] == '0') {
    char c = s[i + 1];
    if (c == 'x' || c == 'X') { base = 16; i += 2; }
    else if (c == 'o' || c == 'O') { base = 8; i += 2; }
    else if (c == 'b' || c == 'B') { base = 2; i += 2; }
    else if (c == 'u' || c == 'U') { unsigned_dec = true; i += 2; }
  }
  if (i >= n) return false;
  auto digit = [&](char c) -> int {
    int d;
    if (c >= '0' && c <= '9') d = c - '0';
    else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
    else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
    else return -1;
    return d < base ? d : -1;
  };
  if (digit(s[i]) < 0) return false;
  unsigned __int128 v = 0;
  const unsigned __int128 lim_signed = ((unsigned __int128)1 << 62);  // max_int + 1
  const unsigned __int128 lim_unsigned = ((unsigned __int128)1 << 63);
  for (; i < n; ++i) {
    if (s[i] == '_') continue;
    int d = digit(s[i]);
    if (d < 0) return false;
    v = v * base + d;
    if (v >= lim_unsigned * 2) return false;
  }
  if (base == 10 && !unsigned_dec) {
    if (neg ? v > lim_signed : v >= lim_signed) return false;
    out = neg ? -(ml_int)v : (ml_int)v;
  } else {
    if (v >= lim_unsigned) return false;
    ml_int r = (ml_int)(v & (lim_unsigned - 1));
    if (v >= lim_signed) r -= (ml_int)lim_unsigned;  // wraps into the negative range like OCaml
    out = neg ? -r : r;
  }
  return true;
}

void usage(FILE *o) {
  // stderr text is not part of the parity contract: written afresh, abbreviated
  fputs("This is the KPopCount oracle (C++ restatement of KPopCount version 18)\n"
        " Usage:\n  KPopCount -l <output_vector_label>|-L [OPTIONS]\n"
        "  -k|-K|--k-mer-size|--k-mer-length <k_mer_length>   (default='12')\n"
        "  -M|--max-results-size <positive_integer>           (default='16777216')\n"
        "  -C|--content 'DNA-ss'|'DNA-single-stranded'|'DNA-ds'|'DNA-double-stranded'|'protein' (default='DNA-ds')\n"
        "  -f|--fasta <fasta_file_name>\n  -s|--single-end <fastq_file_name>\n"
        "  -p|--paired-end <fastq_file_name1> <fastq_file_name2>\n"
        "  -l|--label <output_vector_label>\n  -L|--one-spectrum-per-sequence\n"
        "  -o|--output <output_file_prefix>\n  -v|--verbose\n  -V|--version\n  -h|--help\n", o);
}

int real_main(int argc, char **argv) {
  // Parameters (bin/KPopCount.ml:84-95)
  bool option_l_or_L = false;
  enum { DNA_ss, DNA_ds, Protein } content = DNA_ds;
  ml_int k = 12, max_results_size = 16777216;
  std::vector<Input> inputs;
  std::string label, output;
  bool verbose = false;

  const std::vector<std::string> names = {
      "-k", "-K", "--k-mer-size", "--k-mer-length", "-M", "--max-results-size", "-C", "--content",
      "-f", "--fasta", "-s", "--single-end", "-p", "--paired-end", "-l", "--label",
      "-L", "--one-spectrum-per-sequence", "-o", "--output", "-v", "--verbose", "-V", "--version",
      "--markdown", "-h", "--help"};

  int i = 1;
  auto error = [&](const std::string &msg) -> void { throw ArgvError(msg); };
  auto get_parameter = [&]() -> std::string {
    ++i;
    if (i >= argc) error(std::string("Option '") + argv[i - 1] + "' needs a parameter");
    return argv[i];
  };
  auto get_parameter_int_pos = [&]() -> ml_int {
    std::string p = get_parameter();
    ml_int v;
    if (!ocaml_int_of_string(p, v)) error(std::string("Option '") + argv[i - 1] + "' needs an integer parameter");
    if (!(v > 0)) error(std::string("Option '") + argv[i - 1] + "' needs a positive integer parameter");
    return v;
  };

  while (i < argc) {
    std::string arg = argv[i];
    std::string opt = trie_find_string(names, arg);
    if (opt.empty()) error("Unknown option '" + arg + "'");
    if (opt == "-k" || opt == "-K" || opt == "--k-mer-size" || opt == "--k-mer-length") {
      k = get_parameter_int_pos();
    } else if (opt == "-M" || opt == "--max-results-size") {
      max_results_size = get_parameter_int_pos();
    } else if (opt == "-C" || opt == "--content") {
      std::string w = get_parameter();
      if (w == "DNA-ss" || w == "DNA-single-stranded") content = DNA_ss;
      else if (w == "DNA-ds" || w == "DNA-double-stranded") content = DNA_ds;
      else if (w == "protein" || w == "prot") content = Protein;
      else throw OcamlFailure("KPopCount.Content.Invalid_content(\"" + w + "\")");  // uncaught => exit 2
    } else if (opt == "-f" || opt == "--fasta") {
      inputs.push_back(Input{Input::FASTA, get_parameter(), ""});
    } else if (opt == "-s" || opt == "--single-end") {
      inputs.push_back(Input{Input::SingleEndFASTQ, get_parameter(), ""});
    } else if (opt == "-p" || opt == "--paired-end") {
      std::string n1 = get_parameter();
      std::string n2 = get_parameter();
      inputs.push_back(Input{Input::PairedEndFASTQ, n1, n2});
    } else if (opt == "-l" || opt == "--label") {
      option_l_or_L = true;
      std::string res = get_parameter();
      try {
        label = strip_external_quotes_and_check(res);
      } catch (const QuotesInName &) {
        error("Spectrum labels must not contain quotes");
      }
    } else if (opt == "-L" || opt == "--one-spectrum-per-sequence") {
      option_l_or_L = true;
    } else if (opt == "-o" || opt == "--output") {
      output = make_filename(get_parameter());
    } else if (opt == "-v" || opt == "--verbose") {
      verbose = true;
    } else if (opt == "-V" || opt == "--version") {
      printf("18\n");
      fflush(stdout);
      throw ExitNow{0};
    } else if (opt == "--markdown") {
      usage(stderr);
      throw ExitNow{0};
    } else if (opt == "-h" || opt == "--help") {
      usage(stderr);
      throw ExitNow{1};
    }
    ++i;
  }
  if (!option_l_or_L) error("One of options '-l' and '-L' is mandatory");
  if (verbose) fputs("This is the KPopCount oracle (C++ restatement of KPopCount version 18)\n", stderr);
  if (inputs.empty()) return 0;  // bin/KPopCount.ml:218
  bool is_format_fasta = inputs[0].kind == Input::FASTA;
  for (size_t j = 1; j < inputs.size(); ++j)
    if ((inputs[j].kind == Input::FASTA) != is_format_fasta)
      error("You cannot process FASTA and FASTQ inputs together");

  // functor application: the k range check fires here, before the output is opened
  HashFamily *fam = nullptr;
  DNAHashSingleStranded *ss = nullptr;
  DNAHashDoubleStrandedLexicographic *ds = nullptr;
  ProteinHash *pr = nullptr;
  if (k > 1000000) k = 1000000;  // any value > 30 behaves the same
  switch (content) {
    case DNA_ss: fam = ss = new DNAHashSingleStranded((int)k); break;
    case DNA_ds: fam = ds = new DNAHashDoubleStrandedLexicographic((int)k); break;
    case Protein: fam = pr = new ProteinHash((int)k); break;
  }
  (void)ss; (void)ds; (void)pr;

  FILE *out = stdout;
  if (!output.empty()) {
    out = fopen(output.c_str(), "wb");
    if (!out) throw OcamlFailure("Sys_error(\"" + output + ": " + strerror(errno) + "\")");
  }
  static std::vector<char> obuf(1 << 22);
  setvbuf(out, &obuf[0], _IOFBF, obuf.size());
  int rc = 0;
  try {
    if (!label.empty()) fprintf(out, "\t%s\n", label.c_str());
    Counter c(*fam, content == Protein, max_results_size, label, out);
    for (const Input &in : inputs) {
      switch (in.kind) {
        case Input::FASTA: c.fasta(in.file1); break;
        case Input::SingleEndFASTQ: c.fastq_se(in.file1); break;
        case Input::PairedEndFASTQ: c.fastq_pe(in.file1, in.file2); break;
      }
    }
    c.dump();
    if (verbose) fprintf(stderr, "(oracle): %llu k-mers counted\n", c.kmers_added);
  } catch (const OcamlFailure &e) {
    // at_exit flushes every open channel before the runtime reports the uncaught exception
    fflush(out);
    fprintf(stderr, "Fatal error: exception %s\n", e.what());
    rc = 2;
  } catch (const QuotesInName &e) {
    fflush(out);
    fprintf(stderr, "Fatal error: exception BiOCamLib.Matrix.Base.Quotes_in_name(\"%s\")\n", e.what());
    rc = 2;
  }
  fflush(out);
  if (out != stdout) fclose(out);
  delete fam;
  return rc;
}

}  // namespace

int main(int argc, char **argv) {
  try {
    return real_main(argc, argv);
  } catch (const ArgvError &e) {
    usage(stderr);
    fprintf(stderr, "(BiOCamLib.Tools.Argv.parse): %s\n", e.what());
    return 1;
  } catch (const ExitNow &e) {
    return e.code;
  } catch (const OcamlFailure &e) {
    fflush(stdout);
    fprintf(stderr, "Fatal error: exception %s\n", e.what());
    return 2;
  }
}
