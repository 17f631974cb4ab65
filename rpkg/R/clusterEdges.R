# Sparse counterpart of netcluster() (reference R/clusterbreak.R:112-136) for the edge list returned by
# similarityMH_edges(): the graph igraph sees is the one graph_from_adjacency_matrix(pepmat, mode = "upper",
# weighted = TRUE) would build from the thresholded dense matrix -- every pair with similarity >= threshold plus the
# self-loop that the matrix diagonal contributes to every vertex (1.0 for similarityMH; the self-alignment identities
# that similarityNW_edges() returns in attr(, "self")) -- without an n x n double matrix.
# Louvain itself stays igraph's.  Unexecuted here (no R in the build image); see INTEGRATION.md section 5.

#' @export
netcluster_edges <- function(edges, n, self_weight = if (is.null(attr(edges, "self"))) rep(1, n) else attr(edges, "self"),
                             cluster_func = function(x, ...) igraph::cluster_louvain(x, resolution = 1.05, ...)$membership,
                             cluster_weight = TRUE) {
  if (ncol(edges) != 3) {
    stop("Input must be an edge matrix with columns from, to, weight")
  }
  loops <- seq_len(n)
  el <- rbind(cbind(loops, loops), edges[, 1:2, drop = FALSE])
  network <- igraph::make_empty_graph(n = n, directed = FALSE)
  network <- igraph::add_edges(network, as.vector(t(el)), weight = c(self_weight, edges[, 3]))
  if (cluster_weight) {
    out <- cluster_func(network, weights = igraph::E(network)$weight)
  } else {
    out <- cluster_func(network)
  }
  if (is.numeric(out) && is.vector(out)) {
    return(out)
  } else {
    stop("Wrong clustering output format. Output should be a numeric vector of cluster assignment.")
  }
}
