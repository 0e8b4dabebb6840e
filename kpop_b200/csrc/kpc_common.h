// kpc_common.h -- definitions shared by host and device code of libkpopcount_gpu.
//
// Encoding rules restated from the reference (BiOCamLib/lib/KMers.ml):
//   DNA      A=0 C=1 G=2 T=3, case-insensitive, first base in the most significant bits (:272-277, :324-330)
//   protein  ACDEFGHIKLMNOPQRSTUVWY -> 0..21, 5 bits per residue                         (:150-176, :235-238)
//   hex      width (bits*k+3)/4, lower case                                                 (:151, :270)
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define KPC_HD __host__ __device__ __forceinline__
#define KPC_D __device__ __forceinline__
#else
#define KPC_HD inline
#define KPC_D inline
#endif

enum { KPC_CONTENT_DNA_SS = 0, KPC_CONTENT_DNA_DS = 1, KPC_CONTENT_PROTEIN = 2 };
enum { KPC_FMT_FASTA = 0, KPC_FMT_FASTQ = 1 };

// byte classes written by the framing pass (one per input byte)
enum : uint8_t {
  KPC_CLS_SKIP = 0x40,   // transparent (FASTA line feed)
  KPC_CLS_BREAK = 0x80   // ends every open k-mer window
};

KPC_HD int kpc_symbol_bits(int content) { return content == KPC_CONTENT_PROTEIN ? 5 : 2; }
KPC_HD int kpc_max_k(int content) { return content == KPC_CONTENT_PROTEIN ? 12 : 30; }
KPC_HD int kpc_hex_width(int content, int k) { return (kpc_symbol_bits(content) * k + 3) / 4; }

// Sequences.Lint.dnaize + DNABaseHash.encode_char folded into one map: symbol code or BREAK
KPC_HD uint8_t kpc_classify_dna(uint8_t b) {
  uint8_t u = b & 0xDF;  // only 'A'/'a' map to 0x41 etc.: bit 5 is the sole difference
  uint8_t c = (b >> 1) & 3;  // A->0 C->1 T->2 G->3
  c ^= c >> 1;               // A->0 C->1 G->2 T->3
  bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
  return ok ? c : (uint8_t)KPC_CLS_BREAK;
}
// Sequences.Lint.proteinize + ProteinHash.encode_char: 22 residues, everything else ('*', 'X', ...) breaks
KPC_HD uint8_t kpc_classify_protein(uint8_t b) {
  uint32_t idx = (uint32_t)(b & 0xDF) - 'A';  // b&0xDF in 'A'..'Z'  <=>  b is an ASCII letter
  const uint32_t kMask = 0x17FFDFDu;           // letters of the alphabet (B, J, X, Z excluded)
  if (idx > 25u || !((kMask >> idx) & 1u)) return KPC_CLS_BREAK;
  uint32_t below = kMask & ((1u << idx) - 1u);  // the alphabet is sorted: code = #valid letters below
#if defined(__CUDA_ARCH__)
  return (uint8_t)__popc(below);
#else
  return (uint8_t)__builtin_popcount(below);
#endif
}

KPC_HD uint64_t kpc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
