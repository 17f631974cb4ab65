// Rcpp marshalling shims over the C ABI (include/dynaalign_b200.h).  These replace the bodies of the reference's
// similarityMH (src/minHash.cpp:119) and similarityNW (src/pairwiseSeqAlign.cpp:331): same C++ prototypes, so the
// generated wrappers in RcppExports.cpp are unchanged.  All R API use (string access, allocation, dimnames)
// happens here on R's main thread before/after the GPU call; the CUDA side never sees a SEXP.
#include <Rcpp.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include "dynaalign_b200.h"

using namespace Rcpp;

namespace {

struct Flat {
  std::vector<uint8_t> residues;
  std::vector<int64_t> offsets;
};

Flat flatten(const CharacterVector& sequences) {
  Flat f;
  const size_t n = sequences.length();
  f.offsets.assign(n + 1, 0);
  for (size_t i = 0; i < n; ++i) {
    const std::string s = as<std::string>(sequences[i]);
    f.residues.insert(f.residues.end(), s.begin(), s.end());
    f.offsets[i + 1] = static_cast<int64_t>(f.residues.size());
  }
  if (f.residues.empty()) f.residues.push_back(0);
  return f;
}

int env_gpus() {
  const char* e = std::getenv("DYNAALIGN_GPUS");
  return e ? std::atoi(e) : 1;
}

void set_dimnames(NumericMatrix& m, size_t n) {
  CharacterVector labels(n);
  for (size_t i = 0; i < n; ++i) labels[i] = std::to_string(i + 1);
  m.attr("dimnames") = List::create(labels, labels);
}

void raise_on_error(int rc) {
  if (rc != DYNA_OK) Rcpp::stop("%s", std::string(dyna_last_error()));
}

}  // namespace

// [[Rcpp::export]]
NumericMatrix similarityMH(CharacterVector sequences, int k = 4, int n_hash = 50) {
  const size_t n = sequences.length();
  const Flat f = flatten(sequences);
  NumericMatrix similarityMatrix(n, n);
  // HashFamily seed: std::random_device as in the reference, unless DYNAALIGN_SEED pins it (tests)
  std::vector<uint32_t> seeds;
  const uint32_t* seed_ptr = nullptr;
  if (const char* e = std::getenv("DYNAALIGN_SEED")) {
    if (n_hash > 0) {
      seeds.resize(static_cast<size_t>(n_hash));
      raise_on_error(dyna_hashfamily_seeds(static_cast<uint32_t>(std::strtoul(e, nullptr, 10)), n_hash, seeds.data()));
      seed_ptr = seeds.data();
    }
  }
  raise_on_error(dyna_similarityMH(f.residues.data(), f.offsets.data(), static_cast<int64_t>(n), k, n_hash, seed_ptr,
                                   n ? &similarityMatrix(0, 0) : nullptr, env_gpus()));
  set_dimnames(similarityMatrix, n);
  return similarityMatrix;
}

// [[Rcpp::export]]
NumericMatrix similarityNW(CharacterVector sequences, std::string matrixName = "BLOSUM62", int gapOpen = 10,
                           int gapExt = 4) {
  const size_t n = sequences.length();
  const Flat f = flatten(sequences);
  NumericMatrix similarityMatrix(n, n);
  raise_on_error(dyna_similarityNW(f.residues.data(), f.offsets.data(), static_cast<int64_t>(n), matrixName.c_str(),
                                   gapOpen, gapExt, n ? &similarityMatrix(0, 0) : nullptr, env_gpus()));
  set_dimnames(similarityMatrix, n);
  return similarityMatrix;
}

// similarityMH followed by clusterbreak's threshold step (R/clusterbreak.R:217-221) without the dense n x n matrix:
// returns the edges with similarity >= quantile(sim[upper.tri(sim)], thresh_p) as an (edges x 3) matrix of
// (from, to, weight), 1-based, row-major pair order, with the threshold itself in attr(, "threshold").
// [[Rcpp::export]]
NumericMatrix similarityMH_edges(CharacterVector sequences, int k = 4, int n_hash = 50, double thresh_p = 0.8) {
  const size_t n = sequences.length();
  if (n == 0) Rcpp::stop("Input sequences vector cannot be empty");
  if (k <= 0) Rcpp::stop("'k' must be a positive integer");
  if (n_hash <= 0) Rcpp::stop("Number of hash functions must be positive");
  const Flat f = flatten(sequences);
  std::vector<uint32_t> seeds(static_cast<size_t>(n_hash));
  const char* e = std::getenv("DYNAALIGN_SEED");
  const uint32_t seed = e ? static_cast<uint32_t>(std::strtoul(e, nullptr, 10)) : dyna_random_seed();
  raise_on_error(dyna_hashfamily_seeds(seed, n_hash, seeds.data()));
  dyna_mh_plan* plan = dyna_mh_plan_create(static_cast<int64_t>(n), n_hash, 0, static_cast<int64_t>(n), 0);
  if (!plan) Rcpp::stop("%s", std::string(dyna_last_error()));
  std::vector<uint64_t> hist(static_cast<size_t>(n_hash) + 1, 0);
  std::vector<int32_t> ei, ej;
  std::vector<uint16_t> ec;
  double threshold = 0.0;
  int64_t n_edges = 0;
  int rc = dyna_mh_plan_upload_sequences(plan, f.residues.data(), f.offsets.data(), k, seeds.data(), nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_run_signatures(plan, nullptr);
  // match counts: the join on equal signature values when the data is sparse enough for it (exact either way; the cap is
  // where the all-pairs kernel is certainly faster), otherwise the all-pairs kernel
  int joined = 0;
  if (rc == DYNA_OK) {
    const int64_t pairs = dyna_mh_plan_pairs(plan);
    int64_t incidences = 0;
    rc = dyna_mh_plan_run_match_sparse(plan, std::max<int64_t>(int64_t{1} << 20, pairs * n_hash / 4000), &incidences, &joined, nullptr);
  }
  if (rc == DYNA_OK && !joined) rc = dyna_mh_plan_run_match(plan, nullptr);
  if (rc == DYNA_OK) rc = dyna_mh_plan_count_histogram(plan, hist.data(), nullptr);
  int min_count = 0;
  if (rc == DYNA_OK) rc = dyna_quantile_type7_counts(hist.data(), n_hash, thresh_p, &threshold, &min_count);
  if (rc == DYNA_OK) {
    uint64_t cap = 0;
    for (int c = min_count > 1 ? min_count : 1; c <= n_hash; ++c) cap += hist[static_cast<size_t>(c)];
    ei.resize(cap ? cap : 1);
    ej.resize(cap ? cap : 1);
    ec.resize(cap ? cap : 1);
    rc = dyna_mh_plan_threshold_edges(plan, min_count, static_cast<int64_t>(cap), ei.data(), ej.data(), ec.data(), &n_edges,
                                      nullptr);
  }
  dyna_mh_plan_destroy(plan);  // released before any R error is raised
  raise_on_error(rc);
  NumericMatrix edges(static_cast<size_t>(n_edges), 3);
  for (int64_t q = 0; q < n_edges; ++q) {
    edges(q, 0) = ei[static_cast<size_t>(q)] + 1;
    edges(q, 1) = ej[static_cast<size_t>(q)] + 1;
    edges(q, 2) = static_cast<double>(ec[static_cast<size_t>(q)]) / n_hash;
  }
  edges.attr("threshold") = threshold;
  return edges;
}

// similarityNW followed by the same threshold step: (edges x 3) matrix of (from, to, weight = matches / alignment length),
// 1-based, row-major pair order; attr(, "threshold") is the quantile and attr(, "self") the n self-alignment identities
// (the diagonal of the reference's matrix: the self-loop weights graph_from_adjacency_matrix(mode = "upper") would see).
// [[Rcpp::export]]
NumericMatrix similarityNW_edges(CharacterVector sequences, std::string matrixName = "BLOSUM62", int gapOpen = 10,
                                 int gapExt = 4, double thresh_p = 0.8) {
  const size_t n = sequences.length();
  const Flat f = flatten(sequences);
  dyna_nw_plan* plan = dyna_nw_plan_create(f.residues.data(), f.offsets.data(), static_cast<int64_t>(n), matrixName.c_str(),
                                           gapOpen, gapExt, 0, static_cast<int64_t>(n), 0);
  if (!plan) Rcpp::stop("%s", std::string(dyna_last_error()));
  const size_t ml = static_cast<size_t>(dyna_nw_plan_max_len(plan));
  const size_t mdim = ml + 1, ldim = 2 * ml + 1;
  std::vector<uint64_t> hist(mdim * ldim, 0);
  std::vector<int32_t> ei(1), ej(1);
  std::vector<uint32_t> em(1), el(1), dm(n ? n : 1), dl(n ? n : 1);
  double threshold = 0.0;
  int64_t n_edges = 0;
  int rc = dyna_nw_plan_run(plan, nullptr);
  if (rc == DYNA_OK) rc = dyna_nw_plan_fetch_diagonal(plan, nullptr, 0, dm.data(), dl.data(), nullptr);
  if (rc == DYNA_OK && n >= 2) {  // upper.tri of a 0 x 0 or 1 x 1 matrix is empty: no threshold, no edges
    rc = dyna_nw_plan_stat_histogram(plan, nullptr, 0, hist.data(), nullptr);
    if (rc == DYNA_OK) rc = dyna_quantile_type7_identities(hist.data(), static_cast<int64_t>(mdim), static_cast<int64_t>(ldim), thresh_p, &threshold);
    if (rc == DYNA_OK) {
      uint64_t cap = 0;
      for (size_t m = 1; m < mdim; ++m)
        for (size_t l = 1; l < ldim; ++l)
          if (static_cast<double>(m) / static_cast<double>(l) >= threshold) cap += hist[m * ldim + l];
      ei.resize(cap ? cap : 1);
      ej.resize(cap ? cap : 1);
      em.resize(cap ? cap : 1);
      el.resize(cap ? cap : 1);
      rc = dyna_nw_plan_threshold_edges(plan, nullptr, 0, threshold, static_cast<int64_t>(cap), ei.data(), ej.data(), em.data(),
                                        el.data(), &n_edges, nullptr);
    }
  }
  dyna_nw_plan_destroy(plan);  // released before any R error is raised
  raise_on_error(rc);
  NumericMatrix edges(static_cast<size_t>(n_edges), 3);
  for (int64_t q = 0; q < n_edges; ++q) {
    edges(q, 0) = ei[static_cast<size_t>(q)] + 1;
    edges(q, 1) = ej[static_cast<size_t>(q)] + 1;
    edges(q, 2) = static_cast<double>(em[static_cast<size_t>(q)]) / static_cast<double>(el[static_cast<size_t>(q)]);
  }
  std::vector<double> self(n);
  for (size_t i = 0; i < n; ++i) self[i] = static_cast<double>(dm[i]) / static_cast<double>(dl[i]);  // NaN for an empty string
  edges.attr("threshold") = threshold;
  edges.attr("self") = self;
  return edges;
}

// GPU half of compute_signature_matrix (R/minHash.R): returns the n_hash x n_docs double matrix, Inf where a
// document has no shingle.
// [[Rcpp::export]]
NumericMatrix mh_signatures_linear(IntegerVector ranks, NumericVector offsets, NumericVector a, NumericVector b,
                                   double m, int n_hash) {
  const int64_t n_docs = static_cast<int64_t>(offsets.length()) - 1;
  std::vector<int32_t> rk(ranks.begin(), ranks.end());
  if (rk.empty()) rk.push_back(1);
  std::vector<int64_t> off(offsets.begin(), offsets.end()), av(a.begin(), a.end()), bv(b.begin(), b.end());
  std::vector<uint32_t> sig(static_cast<size_t>(n_docs) * static_cast<size_t>(n_hash));
  raise_on_error(dyna_mh_signatures_linear(rk.data(), off.data(), n_docs, av.data(), bv.data(), static_cast<int64_t>(m),
                                           n_hash, sig.data()));
  NumericMatrix out(n_hash, n_docs);  // column-major: element (h, doc) at h + doc * n_hash == sig[doc * n_hash + h]
  for (int64_t d = 0; d < n_docs; ++d)
    for (int h = 0; h < n_hash; ++h) {
      const uint32_t v = sig[static_cast<size_t>(d) * n_hash + h];
      out(h, d) = (v == 0xFFFFFFFFu) ? std::numeric_limits<double>::infinity() : static_cast<double>(v);
    }
  return out;
}

// GPU half of compute_distance_matrix: codes is the n_hash x n_docs integer matrix of per-row relabelled signatures
// [[Rcpp::export]]
NumericMatrix mh_distance_matrix(IntegerMatrix codes) {
  const int n_hash = codes.nrow();
  const int64_t n_docs = codes.ncol();
  std::vector<uint32_t> sig(static_cast<size_t>(n_docs) * static_cast<size_t>(n_hash));
  for (int64_t d = 0; d < n_docs; ++d)
    for (int h = 0; h < n_hash; ++h) sig[static_cast<size_t>(d) * n_hash + h] = static_cast<uint32_t>(codes(h, d));
  NumericMatrix out(n_docs, n_docs);
  raise_on_error(dyna_mh_match_matrix(sig.data(), n_docs, n_hash, DYNA_MH_DISTANCE, n_docs ? &out(0, 0) : nullptr, 1));
  return out;
}
