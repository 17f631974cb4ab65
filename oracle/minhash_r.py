"""TEST INFRASTRUCTURE ONLY: numpy restatement of the pure-R MinHash pipeline (R/minHash.R).

R is not installed in this image, so this file follows the R source line by line instead of running
it.  It is pinned by the reference's own testthat vectors (tests/testthat/test-minHash.R, restated in
tests/test_r_pipeline_oracle.py): exact values for shingle(), the three error strings, the sorted /
unique / width properties of create_vocab, the {0,1} characteristic matrix, parameter ranges, the
value apply_hash(5,2,3,100)=13, and the 3x3 mock distance matrix (1/3, 1, 1).

Function <-> reference lines:
  shingle                  R/minHash.R:12-23
  create_vocab             R/minHash.R:38-41
  create_char_matrix       R/minHash.R:60-66
  create_hash_parameters   R/minHash.R:81-88   (R's sample() stream cannot be reproduced without R:
                                                ranges and lengths follow the source, the generator is numpy)
  apply_hash               R/minHash.R:104-106
  compute_signature_matrix R/minHash.R:126-143
  compute_distance_matrix  R/minHash.R:166-182
  minhash                  R/minHash.R:206-221
"""
import numpy as np


class RError(Exception):
    """Stands in for an R condition raised by stop()."""


def shingle(x, k):
    # :13-14  !is.character(x) || length(x) != 1
    if not isinstance(x, str):
        raise RError("Input 'x' must be a single character string")
    # :15-16  !is.numeric(k) || length(k) != 1 || k < 1 || k > nchar(x)
    if isinstance(k, bool) or not isinstance(k, (int, float, np.integer, np.floating)) or k < 1 or k > len(x):
        raise RError("'k' must be a positive integer between 1 and %d" % len(x))
    k = int(k)
    n = len(x)
    # :18-21  substr(x, i, i+k-1) for i in 1:(n-k+1)
    return [x[i:i + k] for i in range(0, n - k + 1)]


def create_vocab(sequences, k):
    # :39 unique(unlist(lapply(sequences, shingle, k)))  :40 sort()
    seen = {}
    for s in sequences:
        for sh in shingle(s, k):
            seen.setdefault(sh, None)
    # R's sort() collates by locale; for equal-width upper-case ASCII that is byte order
    return sorted(seen.keys())


def create_char_matrix(sequences, vocab, k):
    # :61-65  sapply(seq_shingles, function(s) as.integer(vocab %in% s))  -> V x N integer
    pos = {v: i for i, v in enumerate(vocab)}
    m = np.zeros((len(vocab), len(sequences)), dtype=np.int32)
    for j, s in enumerate(sequences):
        for sh in shingle(s, k):
            if sh in pos:
                m[pos[sh], j] = 1
    return m


def create_hash_parameters(n_hash, max_val, rng=None):
    # :82-83
    if n_hash < 1:
        raise RError("Number of hash functions must be positive")
    if max_val < 2:
        raise RError("Maximum value must be at least 2")
    rng = np.random.default_rng() if rng is None else rng
    # :85 sample(1:max_val, n_hash, replace=TRUE)   :86 sample(0:max_val, n_hash, replace=TRUE)
    a = rng.integers(1, max_val + 1, size=n_hash, dtype=np.int64)
    b = rng.integers(0, max_val + 1, size=n_hash, dtype=np.int64)
    return {"a": a, "b": b}


def apply_hash(x, a, b, m):
    # :105  (a * x + b) %% m
    return (np.asarray(a) * np.asarray(x) + np.asarray(b)) % m


def compute_signature_matrix(char_matrix, hash_params, max_val):
    # :127-129
    a = np.asarray(hash_params["a"], dtype=np.int64)
    b = np.asarray(hash_params["b"], dtype=np.int64)
    n_hash = len(a)
    n_docs = char_matrix.shape[1]
    sig = np.full((n_hash, n_docs), np.inf, dtype=np.float64)
    # :131-141  row index i is the 1-based vocabulary rank
    for i in range(1, char_matrix.shape[0] + 1):
        hv = apply_hash(i, a, b, max_val).astype(np.float64)
        for j in range(n_docs):
            if char_matrix[i - 1, j] == 1:
                sig[:, j] = np.minimum(sig[:, j], hv)
    return sig


def compute_distance_matrix(sig_matrix):
    # :168-179
    sig_matrix = np.asarray(sig_matrix)
    n_docs = sig_matrix.shape[1]
    n_hash = sig_matrix.shape[0]
    d = np.zeros((n_docs, n_docs), dtype=np.float64)
    for i in range(n_docs):
        for j in range(i, n_docs):
            if i != j:
                cnt = int(np.count_nonzero(sig_matrix[:, i] == sig_matrix[:, j]))
                # mean() of a logical: long double sum / n, rounded to double
                sim = float(np.longdouble(cnt) / np.longdouble(n_hash))
                d[i, j] = 1 - sim
                d[j, i] = d[i, j]
    return d


def minhash(sequences, k, n_hash, rng=None, hash_params=None):
    # :208-220
    vocab = create_vocab(sequences, k)
    char_matrix = create_char_matrix(sequences, vocab, k)
    max_val = len(vocab)
    if hash_params is None:
        hash_params = create_hash_parameters(n_hash, max_val, rng)
    sig_matrix = compute_signature_matrix(char_matrix, hash_params, max_val)
    dist_matrix = compute_distance_matrix(sig_matrix)
    return {"vocabulary": vocab, "char_matrix": char_matrix, "sig_matrix": sig_matrix, "dist_matrix": dist_matrix,
            "hash_params": hash_params}


def shingle_ranks(sequences, vocab, k):
    """Helper for the GPU path's input: per document, the 1-based vocabulary rank of every shingle
    (duplicates kept; min is idempotent).  Returns (int32 ranks, int64 offsets[n+1])."""
    pos = {v: i + 1 for i, v in enumerate(vocab)}
    ranks, offsets = [], [0]
    for s in sequences:
        ranks.extend(pos[sh] for sh in shingle(s, k))
        offsets.append(len(ranks))
    return np.asarray(ranks, dtype=np.int32), np.asarray(offsets, dtype=np.int64)
