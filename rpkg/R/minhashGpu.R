# MinHash pipeline with the reference's exported names, argument lists, return values and error messages
# (reference: R/minHash.R).  Everything that handles strings -- shingles, vocabulary, characteristic matrix, parameter
# sampling -- stays in R; the two quadratic loops, the signature minimum and the all-pairs distance, run on the GPU
# through the internal .Call stubs of RcppExports.R.  Unexecuted in the build image (no R); see INTEGRATION.md.

.stop_plain <- function(msg) stop(msg, call. = FALSE)

#' @export
shingle <- function(x, k) {
  single_string <- is.character(x) && length(x) == 1L
  if (!single_string) .stop_plain("Input 'x' must be a single character string")
  len <- nchar(x)
  k_ok <- is.numeric(k) && length(k) == 1L && k >= 1 && k <= len
  if (!k_ok) .stop_plain(sprintf("'k' must be a positive integer between 1 and %d", len))
  first <- seq_len(len - k + 1L)
  substring(x, first, first + k - 1L)
}

#' @export
create_vocab <- function(sequences, k) {
  all_shingles <- unlist(lapply(sequences, shingle, k = k))
  sort(unique(all_shingles))
}

#' @export
create_char_matrix <- function(sequences, vocab, k) {
  per_doc <- lapply(sequences, shingle, k = k)
  vapply(per_doc, function(s) as.integer(vocab %in% s), integer(length(vocab)))
}

#' @export
create_hash_parameters <- function(n_hash, max_val) {
  if (n_hash < 1) {
    stop("Number of hash functions must be positive")
  }
  if (max_val < 2) {
    stop("Maximum value must be at least 2")
  }
  draw <- function(lowest) sample(lowest:max_val, n_hash, replace = TRUE)
  list(a = draw(1L), b = draw(0L))
}

#' @export
apply_hash <- function(x, a, b, m) (a * x + b) %% m

#' @export
compute_signature_matrix <- function(char_matrix, hash_params, max_val) {
  docs <- ncol(char_matrix)
  # (vocabulary rank, document) of every set bit, grouped by document: the rank lists the device kernel walks
  hit <- which(char_matrix == 1, arr.ind = TRUE)
  hit <- hit[order(hit[, 2], hit[, 1]), , drop = FALSE]
  starts <- c(0, cumsum(tabulate(hit[, 2], nbins = docs)))
  .mh_signatures_linear(as.integer(hit[, 1]), as.numeric(starts), as.numeric(hash_params$a),
                        as.numeric(hash_params$b), as.numeric(max_val), length(hash_params$a))
}

#' @export
compute_distance_matrix <- function(sig_matrix) {
  # only equality matters: relabel every hash row to dense integer codes, then count matches on the GPU
  # (filled row by row: t(apply(...)) would come back transposed when there is a single document)
  codes <- matrix(0L, nrow = nrow(sig_matrix), ncol = ncol(sig_matrix))
  for (h in seq_len(nrow(sig_matrix))) {
    r <- sig_matrix[h, ]
    codes[h, ] <- match(r, unique(r)) - 1L
  }
  .mh_distance_matrix(codes)
}

#' @export
minhash <- function(sequences, k, n_hash) {
  vocabulary <- create_vocab(sequences, k)
  membership <- create_char_matrix(sequences, vocabulary, k)
  params <- create_hash_parameters(n_hash, length(vocabulary))
  signatures <- compute_signature_matrix(membership, params, length(vocabulary))
  list(vocabulary = vocabulary, char_matrix = membership, sig_matrix = signatures,
       dist_matrix = compute_distance_matrix(signatures))
}
