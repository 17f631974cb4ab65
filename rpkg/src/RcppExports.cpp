// Routine registration for the DynaAlign shared object, hand-written (not compileAttributes output).
// The first two entry points keep the reference's symbols and arities -- _DynaAlign_similarityMH/3 and
// _DynaAlign_similarityNW/4 (its src/RcppExports.cpp:15,28,41-45) -- so `useDynLib(DynaAlign, .registration = TRUE)`
// and the R stubs keep working unchanged.  Four more serve this package's own R code: the GPU halves of the pure-R
// minhash() pipeline and the sparse threshold step of clusterbreak.
#include <Rcpp.h>

#include <string>

using Rcpp::CharacterVector;
using Rcpp::IntegerMatrix;
using Rcpp::IntegerVector;
using Rcpp::NumericMatrix;
using Rcpp::NumericVector;

// implemented in dyna_shims.cpp
NumericMatrix similarityMH(CharacterVector sequences, int k, int n_hash);
NumericMatrix similarityNW(CharacterVector sequences, std::string matrixName, int gapOpen, int gapExt);
NumericMatrix similarityMH_edges(CharacterVector sequences, int k, int n_hash, double thresh_p);
NumericMatrix similarityNW_edges(CharacterVector sequences, std::string matrixName, int gapOpen, int gapExt, double thresh_p);
NumericMatrix mh_signatures_linear(IntegerVector ranks, NumericVector offsets, NumericVector a, NumericVector b, double m,
                                   int n_hash);
NumericMatrix mh_distance_matrix(IntegerMatrix codes);

// BEGIN_RCPP / END_RCPP turn C++ exceptions (Rcpp::stop in the shims) into R errors, as in the reference's wrappers;
// an RNGScope is held for the two reference entry points because the reference's wrappers hold one as well.
RcppExport SEXP _DynaAlign_similarityMH(SEXP seqs, SEXP k, SEXP nHash) {
  BEGIN_RCPP
  Rcpp::RNGScope rng;
  return Rcpp::wrap(similarityMH(Rcpp::as<CharacterVector>(seqs), Rcpp::as<int>(k), Rcpp::as<int>(nHash)));
  END_RCPP
}

RcppExport SEXP _DynaAlign_similarityNW(SEXP seqs, SEXP table, SEXP open, SEXP ext) {
  BEGIN_RCPP
  Rcpp::RNGScope rng;
  return Rcpp::wrap(similarityNW(Rcpp::as<CharacterVector>(seqs), Rcpp::as<std::string>(table), Rcpp::as<int>(open),
                                 Rcpp::as<int>(ext)));
  END_RCPP
}

RcppExport SEXP _DynaAlign_similarityMH_edges(SEXP seqs, SEXP k, SEXP nHash, SEXP p) {
  BEGIN_RCPP
  return Rcpp::wrap(similarityMH_edges(Rcpp::as<CharacterVector>(seqs), Rcpp::as<int>(k), Rcpp::as<int>(nHash),
                                       Rcpp::as<double>(p)));
  END_RCPP
}

RcppExport SEXP _DynaAlign_similarityNW_edges(SEXP seqs, SEXP table, SEXP open, SEXP ext, SEXP p) {
  BEGIN_RCPP
  return Rcpp::wrap(similarityNW_edges(Rcpp::as<CharacterVector>(seqs), Rcpp::as<std::string>(table), Rcpp::as<int>(open),
                                       Rcpp::as<int>(ext), Rcpp::as<double>(p)));
  END_RCPP
}

RcppExport SEXP _DynaAlign_mh_signatures_linear(SEXP ranks, SEXP offsets, SEXP a, SEXP b, SEXP m, SEXP nHash) {
  BEGIN_RCPP
  return Rcpp::wrap(mh_signatures_linear(Rcpp::as<IntegerVector>(ranks), Rcpp::as<NumericVector>(offsets),
                                         Rcpp::as<NumericVector>(a), Rcpp::as<NumericVector>(b), Rcpp::as<double>(m),
                                         Rcpp::as<int>(nHash)));
  END_RCPP
}

RcppExport SEXP _DynaAlign_mh_distance_matrix(SEXP codes) {
  BEGIN_RCPP
  return Rcpp::wrap(mh_distance_matrix(Rcpp::as<IntegerMatrix>(codes)));
  END_RCPP
}

static const R_CallMethodDef kCallEntries[] = {
    {"_DynaAlign_similarityMH", reinterpret_cast<DL_FUNC>(&_DynaAlign_similarityMH), 3},
    {"_DynaAlign_similarityNW", reinterpret_cast<DL_FUNC>(&_DynaAlign_similarityNW), 4},
    {"_DynaAlign_similarityMH_edges", reinterpret_cast<DL_FUNC>(&_DynaAlign_similarityMH_edges), 4},
    {"_DynaAlign_similarityNW_edges", reinterpret_cast<DL_FUNC>(&_DynaAlign_similarityNW_edges), 5},
    {"_DynaAlign_mh_signatures_linear", reinterpret_cast<DL_FUNC>(&_DynaAlign_mh_signatures_linear), 6},
    {"_DynaAlign_mh_distance_matrix", reinterpret_cast<DL_FUNC>(&_DynaAlign_mh_distance_matrix), 1},
    {nullptr, nullptr, 0}};

RcppExport void R_init_DynaAlign(DllInfo* dll) {
  R_registerRoutines(dll, nullptr, kCallEntries, nullptr, nullptr);
  R_useDynamicSymbols(dll, FALSE);
}
