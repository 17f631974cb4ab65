// TEST INFRASTRUCTURE: compiles the R package's Rcpp shims (rpkg/src/dyna_shims.cpp) against the stub Rcpp.h and
// exposes them through a C interface, so the marshalling layer a maintainer would ship (string flattening, error
// propagation, dimnames, matrix layout) is exercised end to end even though R itself is not installed here.
#include <Rcpp.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../rpkg/src/dyna_shims.cpp"

namespace {
thread_local std::string g_err;
Rcpp::CharacterVector to_cv(const char* residues, const int64_t* offsets, int64_t n) {
  std::vector<std::string> v(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) v[static_cast<size_t>(i)].assign(residues + offsets[i], static_cast<size_t>(offsets[i + 1] - offsets[i]));
  return Rcpp::CharacterVector(std::move(v));
}
int dimnames_ok(const Rcpp::NumericMatrix& m, int64_t n) {
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
const char* shim_last_error() { return g_err.c_str(); }

int shim_similarityMH(const char* residues, const int64_t* offsets, int64_t n, int k, int n_hash, double* out, int* dn_ok) {
  try {
    Rcpp::NumericMatrix m = similarityMH(to_cv(residues, offsets, n), k, n_hash);
    if (out && n) std::memcpy(out, m.begin(), sizeof(double) * m.nrow() * m.ncol());
    if (dn_ok) *dn_ok = dimnames_ok(m, n);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int shim_similarityNW(const char* residues, const int64_t* offsets, int64_t n, const char* name, int go, int ge, double* out,
                      int* dn_ok) {
  try {
    Rcpp::NumericMatrix m = similarityNW(to_cv(residues, offsets, n), std::string(name), go, ge);
    if (out && n) std::memcpy(out, m.begin(), sizeof(double) * m.nrow() * m.ncol());
    if (dn_ok) *dn_ok = dimnames_ok(m, n);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int shim_mh_signatures_linear(const int* ranks, int64_t n_ranks, const double* offsets, int64_t n_off, const double* a,
                              const double* b, double m, int n_hash, double* out) {
  try {
    Rcpp::NumericMatrix r = mh_signatures_linear(Rcpp::IntegerVector(std::vector<int>(ranks, ranks + n_ranks)),
                                                 Rcpp::NumericVector(std::vector<double>(offsets, offsets + n_off)),
                                                 Rcpp::NumericVector(std::vector<double>(a, a + n_hash)),
                                                 Rcpp::NumericVector(std::vector<double>(b, b + n_hash)), m, n_hash);
    std::memcpy(out, r.begin(), sizeof(double) * r.nrow() * r.ncol());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int shim_mh_distance_matrix(const int* codes, int n_hash, int n_docs, double* out) {
  try {
    Rcpp::NumericMatrix r = mh_distance_matrix(
        Rcpp::IntegerMatrix(n_hash, n_docs, std::vector<int>(codes, codes + static_cast<size_t>(n_hash) * n_docs)));
    std::memcpy(out, r.begin(), sizeof(double) * r.nrow() * r.ncol());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// edges come back through a caller buffer of `cap` rows; *n_edges is the real count either way
int shim_similarityMH_edges(const char* residues, const int64_t* offsets, int64_t n, int k, int n_hash, double thresh_p,
                            int64_t cap, double* edges_colmajor, int64_t* n_edges, double* threshold) {
  try {
    Rcpp::NumericMatrix m = similarityMH_edges(to_cv(residues, offsets, n), k, n_hash, thresh_p);
    *n_edges = static_cast<int64_t>(m.nrow());
    *threshold = m.scalar_attr();
    if (static_cast<int64_t>(m.nrow()) <= cap)
      for (size_t c = 0; c < 3; ++c)
        for (size_t r = 0; r < m.nrow(); ++r) edges_colmajor[c * static_cast<size_t>(cap) + r] = m(r, c);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int shim_similarityNW_edges(const char* residues, const int64_t* offsets, int64_t n, const char* name, int go, int ge,
                            double thresh_p, int64_t cap, double* edges_colmajor, int64_t* n_edges, double* threshold,
                            double* self_out /* n */) {
  try {
    Rcpp::NumericMatrix m = similarityNW_edges(to_cv(residues, offsets, n), std::string(name), go, ge, thresh_p);
    *n_edges = static_cast<int64_t>(m.nrow());
    *threshold = m.scalar_attr();
    if (static_cast<int64_t>(m.vector_attr().size()) != n) throw std::runtime_error("attr self has the wrong length");
    for (int64_t i = 0; i < n; ++i) self_out[i] = m.vector_attr()[static_cast<size_t>(i)];
    if (static_cast<int64_t>(m.nrow()) <= cap)
      for (size_t c = 0; c < 3; ++c)
        for (size_t r = 0; r < m.nrow(); ++r) edges_colmajor[c * static_cast<size_t>(cap) + r] = m(r, c);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}
