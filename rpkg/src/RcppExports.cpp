// Routine registration for the DynaAlign shared object.  The first two entries are byte-for-byte the reference's
// symbols and arities (_DynaAlign_similarityMH/3, _DynaAlign_similarityNW/4; its src/RcppExports.cpp:15,28,41-45),
// so `useDynLib(DynaAlign, .registration = TRUE)` and the R stubs are unchanged.  Two internal entries serve the
// GPU halves of the pure-R minhash() pipeline and one the sparse threshold step of clusterbreak.
#include <Rcpp.h>

using namespace Rcpp;

NumericMatrix similarityMH(CharacterVector sequences, int k, int n_hash);
RcppExport SEXP _DynaAlign_similarityMH(SEXP sequencesSEXP, SEXP kSEXP, SEXP n_hashSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::RNGScope rcpp_rngScope_gen;
    Rcpp::traits::input_parameter< CharacterVector >::type sequences(sequencesSEXP);
    Rcpp::traits::input_parameter< int >::type k(kSEXP);
    Rcpp::traits::input_parameter< int >::type n_hash(n_hashSEXP);
    rcpp_result_gen = Rcpp::wrap(similarityMH(sequences, k, n_hash));
    return rcpp_result_gen;
END_RCPP
}

NumericMatrix similarityNW(CharacterVector sequences, std::string matrixName, int gapOpen, int gapExt);
RcppExport SEXP _DynaAlign_similarityNW(SEXP sequencesSEXP, SEXP matrixNameSEXP, SEXP gapOpenSEXP, SEXP gapExtSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::RNGScope rcpp_rngScope_gen;
    Rcpp::traits::input_parameter< CharacterVector >::type sequences(sequencesSEXP);
    Rcpp::traits::input_parameter< std::string >::type matrixName(matrixNameSEXP);
    Rcpp::traits::input_parameter< int >::type gapOpen(gapOpenSEXP);
    Rcpp::traits::input_parameter< int >::type gapExt(gapExtSEXP);
    rcpp_result_gen = Rcpp::wrap(similarityNW(sequences, matrixName, gapOpen, gapExt));
    return rcpp_result_gen;
END_RCPP
}

NumericMatrix mh_signatures_linear(IntegerVector ranks, NumericVector offsets, NumericVector a, NumericVector b, double m, int n_hash);
RcppExport SEXP _DynaAlign_mh_signatures_linear(SEXP ranksSEXP, SEXP offsetsSEXP, SEXP aSEXP, SEXP bSEXP, SEXP mSEXP, SEXP n_hashSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< IntegerVector >::type ranks(ranksSEXP);
    Rcpp::traits::input_parameter< NumericVector >::type offsets(offsetsSEXP);
    Rcpp::traits::input_parameter< NumericVector >::type a(aSEXP);
    Rcpp::traits::input_parameter< NumericVector >::type b(bSEXP);
    Rcpp::traits::input_parameter< double >::type m(mSEXP);
    Rcpp::traits::input_parameter< int >::type n_hash(n_hashSEXP);
    rcpp_result_gen = Rcpp::wrap(mh_signatures_linear(ranks, offsets, a, b, m, n_hash));
    return rcpp_result_gen;
END_RCPP
}

NumericMatrix mh_distance_matrix(IntegerMatrix codes);
RcppExport SEXP _DynaAlign_mh_distance_matrix(SEXP codesSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< IntegerMatrix >::type codes(codesSEXP);
    rcpp_result_gen = Rcpp::wrap(mh_distance_matrix(codes));
    return rcpp_result_gen;
END_RCPP
}

NumericMatrix similarityMH_edges(CharacterVector sequences, int k, int n_hash, double thresh_p);
RcppExport SEXP _DynaAlign_similarityMH_edges(SEXP sequencesSEXP, SEXP kSEXP, SEXP n_hashSEXP, SEXP thresh_pSEXP) {
BEGIN_RCPP
    Rcpp::RObject rcpp_result_gen;
    Rcpp::traits::input_parameter< CharacterVector >::type sequences(sequencesSEXP);
    Rcpp::traits::input_parameter< int >::type k(kSEXP);
    Rcpp::traits::input_parameter< int >::type n_hash(n_hashSEXP);
    Rcpp::traits::input_parameter< double >::type thresh_p(thresh_pSEXP);
    rcpp_result_gen = Rcpp::wrap(similarityMH_edges(sequences, k, n_hash, thresh_p));
    return rcpp_result_gen;
END_RCPP
}

static const R_CallMethodDef CallEntries[] = {
    {"_DynaAlign_similarityMH", (DL_FUNC) &_DynaAlign_similarityMH, 3},
    {"_DynaAlign_similarityNW", (DL_FUNC) &_DynaAlign_similarityNW, 4},
    {"_DynaAlign_mh_signatures_linear", (DL_FUNC) &_DynaAlign_mh_signatures_linear, 6},
    {"_DynaAlign_mh_distance_matrix", (DL_FUNC) &_DynaAlign_mh_distance_matrix, 1},
    {"_DynaAlign_similarityMH_edges", (DL_FUNC) &_DynaAlign_similarityMH_edges, 4},
    {NULL, NULL, 0}
};

RcppExport void R_init_DynaAlign(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
