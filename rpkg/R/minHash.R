# MinHash pipeline with the reference's exported names and semantics.  String handling (shingles, vocabulary,
# characteristic matrix, parameter sampling) stays in R, exactly as in the reference; the two quadratic loops --
# the signature minimum and the all-pairs distance -- run on the GPU.

#' @export
shingle <- function(x, k) {
  if (!is.character(x) || length(x) != 1)
    stop("Input 'x' must be a single character string", call. = FALSE)
  if (!is.numeric(k) || length(k) != 1 || k < 1 || k > nchar(x))
    stop(sprintf("'k' must be a positive integer between 1 and %d", nchar(x)), call. = FALSE)
  starts <- seq_len(nchar(x) - k + 1)
  substring(x, starts, starts + k - 1)
}

#' @export
create_vocab <- function(sequences, k) {
  sort(unique(unlist(lapply(sequences, shingle, k = k))))
}

#' @export
create_char_matrix <- function(sequences, vocab, k) {
  sapply(lapply(sequences, shingle, k = k), function(s) as.integer(vocab %in% s))
}

#' @export
create_hash_parameters <- function(n_hash, max_val) {
  if (n_hash < 1) stop("Number of hash functions must be positive")
  if (max_val < 2) stop("Maximum value must be at least 2")
  list(a = sample(1:max_val, n_hash, replace = TRUE),
       b = sample(0:max_val, n_hash, replace = TRUE))
}

#' @export
apply_hash <- function(x, a, b, m) {
  (a * x + b) %% m
}

#' @export
compute_signature_matrix <- function(char_matrix, hash_params, max_val) {
  n_hash <- length(hash_params$a)
  n_docs <- ncol(char_matrix)
  hit <- which(char_matrix == 1, arr.ind = TRUE)            # row = 1-based vocabulary rank, col = document
  hit <- hit[order(hit[, 2], hit[, 1]), , drop = FALSE]
  offsets <- c(0, cumsum(tabulate(hit[, 2], nbins = n_docs)))
  .mh_signatures_linear(as.integer(hit[, 1]), as.numeric(offsets), as.numeric(hash_params$a),
                        as.numeric(hash_params$b), as.numeric(max_val), as.integer(n_hash))
}

#' @export
compute_distance_matrix <- function(sig_matrix) {
  # equality-preserving relabelling of each hash row to dense integer codes, then the GPU match counter
  codes <- t(apply(sig_matrix, 1, function(r) match(r, unique(r)) - 1L))
  storage.mode(codes) <- "integer"
  .mh_distance_matrix(codes)
}

#' @export
minhash <- function(sequences, k, n_hash) {
  vocab <- create_vocab(sequences, k)
  char_matrix <- create_char_matrix(sequences, vocab, k)
  max_val <- length(vocab)
  hash_params <- create_hash_parameters(n_hash, max_val)
  sig_matrix <- compute_signature_matrix(char_matrix, hash_params, max_val)
  dist_matrix <- compute_distance_matrix(sig_matrix)
  list(vocabulary = vocab, char_matrix = char_matrix, sig_matrix = sig_matrix, dist_matrix = dist_matrix)
}
