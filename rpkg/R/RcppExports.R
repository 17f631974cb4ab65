# R stubs of the native entry points.  Names, argument order, defaults and the .Call symbols of the first two are those
# of the reference package (its R/RcppExports.R:15,34), so callers such as clusterbreak(sim_fn = ...) need no change.

#' @export
similarityMH <- function(sequences, k = 4L, n_hash = 50L)
    .Call(`_DynaAlign_similarityMH`, sequences, k, n_hash)

#' @export
similarityNW <- function(sequences, matrixName = "BLOSUM62", gapOpen = 10L, gapExt = 4L)
    .Call(`_DynaAlign_similarityNW`, sequences, matrixName, gapOpen, gapExt)

#' similarityMH + clusterbreak's quantile threshold, as an edge list (from, to, weight) instead of a dense matrix
#' @export
similarityMH_edges <- function(sequences, k = 4L, n_hash = 50L, thresh_p = 0.8)
    .Call(`_DynaAlign_similarityMH_edges`, sequences, k, n_hash, thresh_p)

#' similarityNW + clusterbreak's quantile threshold, as an edge list (from, to, weight); attr "threshold", attr "self"
#' (the self-alignment identities, i.e. the diagonal of the dense matrix)
#' @export
similarityNW_edges <- function(sequences, matrixName = "BLOSUM62", gapOpen = 10L, gapExt = 4L, thresh_p = 0.8)
    .Call(`_DynaAlign_similarityNW_edges`, sequences, matrixName, gapOpen, gapExt, thresh_p)

# internal: GPU halves of the pure-R MinHash pipeline (minhashGpu.R)
.mh_signatures_linear <- function(ranks, offsets, a, b, m, n_hash)
    .Call(`_DynaAlign_mh_signatures_linear`, ranks, offsets, a, b, m, n_hash)

.mh_distance_matrix <- function(codes)
    .Call(`_DynaAlign_mh_distance_matrix`, codes)
