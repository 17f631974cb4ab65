// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Builds oracle/_ref/libdynaref.so: the reference's own C++ (src/minHash.cpp and
// src/pairwiseSeqAlign.cpp), compiled UNMODIFIED from /root/reference (passed as
// -I$(REF)/src; no reference source is copied into this repository), behind a
// small C ABI that the tests and bench.py's cpu_baseline / --impl reference leg
// drive through ctypes.
//
// The only intervention is seed injection for similarityMH: the reference seeds
// HashFamily from std::random_device (src/minHash.cpp:73,137), so bit-exact
// comparison is only defined for a given seed.  <random> is included first, then
// the token `random_device` is macro-renamed to a functor returning the injected
// seed while minHash.cpp is parsed.  The reference code itself is untouched.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <Rcpp.h>

static unsigned int g_injected_seed = 42u;
struct dyna_fixed_rd {
  unsigned int operator()() const { return g_injected_seed; }
};

#define random_device dyna_fixed_rd
#include "minHash.cpp"            // reference, from -I$(REF)/src
#undef random_device
#include "pairwiseSeqAlign.cpp"   // reference, from -I$(REF)/src

namespace {
thread_local std::string g_err;

Rcpp::CharacterVector to_cv(const char* residues, const int64_t* offsets, int64_t n) {
  std::vector<std::string> v(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i)
    v[static_cast<size_t>(i)].assign(residues + offsets[i], static_cast<size_t>(offsets[i + 1] - offsets[i]));
  return Rcpp::CharacterVector(std::move(v));
}

int check_dimnames(const Rcpp::NumericMatrix& m, int64_t n) {
  const Rcpp::List& dn = m.dimnames();
  if (dn.items.size() != 2) return 0;
  for (int a = 0; a < 2; ++a) {
    if (dn.items[a].length() != n) return 0;
    for (int64_t i = 0; i < n; ++i)
      if (dn.items[a][static_cast<size_t>(i)] != std::to_string(i + 1)) return 0;
  }
  return 1;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}
int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// reference similarityMH (src/minHash.cpp:119) with random_device -> seed
int ref_similarityMH(const char* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                     unsigned int seed, double* out_colmajor, int* dimnames_ok) {
  try {
    g_injected_seed = seed;
    Rcpp::NumericMatrix m = similarityMH(to_cv(residues, offsets, n), k, n_hash);
    if (out_colmajor) std::memcpy(out_colmajor, m.begin(), sizeof(double) * m.nrow() * m.ncol());
    if (dimnames_ok) *dimnames_ok = check_dimnames(m, n);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// reference similarityNW (src/pairwiseSeqAlign.cpp:331)
int ref_similarityNW(const char* residues, const int64_t* offsets, int64_t n, const char* matrix_name,
                     int gap_open, int gap_ext, double* out_colmajor, int* dimnames_ok) {
  try {
    Rcpp::NumericMatrix m = similarityNW(to_cv(residues, offsets, n), std::string(matrix_name), gap_open, gap_ext);
    if (out_colmajor) std::memcpy(out_colmajor, m.begin(), sizeof(double) * m.nrow() * m.ncol());
    if (dimnames_ok) *dimnames_ok = check_dimnames(m, n);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// reference calculate_similarity (src/pairwiseSeqAlign.cpp:209) for one ordered pair
int ref_calculate_similarity(const char* a, int64_t la, const char* b, int64_t lb, const char* matrix_name,
                             int gap_open, int gap_ext, double* out) {
  try {
    const int (*S)[24] = getSubstitutionMatrix(std::string(matrix_name));
    *out = calculate_similarity(std::string(a, static_cast<size_t>(la)), std::string(b, static_cast<size_t>(lb)),
                                S, gap_open, gap_ext);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// reference substitution table (src/pairwiseSeqAlign.cpp:190) -> 576 ints row-major
int ref_substitution_matrix(const char* matrix_name, int* out576) {
  try {
    const int (*S)[24] = getSubstitutionMatrix(std::string(matrix_name));
    for (int i = 0; i < 24; ++i) for (int j = 0; j < 24; ++j) out576[i * 24 + j] = S[i][j];
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// reference aa_to_index (src/pairwiseSeqAlign.cpp:15): 256-entry table, -1 = not in alphabet
void ref_aa_index_table(int* out256) {
  for (int c = 0; c < 256; ++c) {
    auto it = aa_to_index.find(static_cast<char>(c));
    out256[c] = (it == aa_to_index.end()) ? -1 : it->second;
  }
}

// reference murmur3_32 (src/minHash.cpp:21)
uint32_t ref_murmur3_32(const char* key, uint64_t len, uint32_t seed) {
  // the reference reads 32-bit blocks through a cast pointer; hand it an aligned copy
  std::vector<uint32_t> buf((len + 7) / 4 + 1, 0u);
  std::memcpy(buf.data(), key, len);
  return murmur3_32(reinterpret_cast<const char*>(buf.data()), static_cast<size_t>(len), seed);
}

// reference HashFamily seeds (src/minHash.cpp:67-89): recovered by hashing through the
// class is not possible (seeds are private), so expose hash(kmer, index) and the
// documented construction (mt19937(seed) stream) separately for cross-checking.
int ref_hashfamily_hash(unsigned int seed, int n_hash, const char* kmer, int64_t klen, uint32_t* out_nhash) {
  try {
    HashFamily hf(n_hash, seed);
    std::string s(kmer, static_cast<size_t>(klen));
    for (int h = 0; h < n_hash; ++h) out_nhash[h] = hf.hash(s, h);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// signatures exactly as the reference's loop builds them (src/minHash.cpp:140-157)
int ref_mh_signatures(const char* residues, const int64_t* offsets, int64_t n, int k, int n_hash,
                      unsigned int seed, uint32_t* sig_rowmajor) {
  try {
    HashFamily hf(n_hash, seed);
    for (int64_t i = 0; i < n; ++i) {
      std::string seq(residues + offsets[i], static_cast<size_t>(offsets[i + 1] - offsets[i]));
      std::vector<std::string> kmers = generate_kmers(seq, k);
      uint32_t* row = sig_rowmajor + i * n_hash;
      for (int h = 0; h < n_hash; ++h) row[h] = UINT32_MAX;
      for (const std::string& km : kmers)
        for (int h = 0; h < n_hash; ++h) row[h] = std::min(row[h], hf.hash(km, h));
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"
