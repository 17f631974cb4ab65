"""Host-side mirror of the reference's R-facing interface for the all-pairs similarity hot path.

Same names, argument meaning, defaults and error text as the reference:

  similarityMH(sequences, k=4, n_hash=50)                        R/RcppExports.R:15  -> src/minHash.cpp:119
  similarityNW(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4)
                                                                 R/RcppExports.R:34  -> src/pairwiseSeqAlign.cpp:331
  shingle, create_vocab, create_char_matrix, create_hash_parameters, apply_hash,
  compute_signature_matrix, compute_distance_matrix, minhash     R/minHash.R

R is not available in the build image, so this Python layer plays the role of the R stubs: it marshals the
character vector into the flat (residues, offsets) form and calls the C ABI through ctypes -- exactly what the
Rcpp shims in rpkg/src do.  All numeric work happens in the CUDA library; nothing here computes a similarity.
Results are column-major (Fortran-order) float64 n x n arrays, symmetric, like the reference's NumericMatrix;
``dimnames`` ("1".."n") are returned by :func:`dimnames`.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import DynaAlignError, check, flatten, lib, ptr

__all__ = ["similarityMH", "similarityNW", "shingle", "create_vocab", "create_char_matrix", "create_hash_parameters",
           "apply_hash", "compute_signature_matrix", "compute_distance_matrix", "minhash", "dimnames",
           "hashfamily_seeds", "mh_signatures", "mh_match_counts", "nw_pair_stats", "partition_rows",
           "substitution_matrix", "quantile_type7_counts", "similarityMH_edges", "similarityNW_edges", "quantile_type7_identities", "MinHashPlan", "NWPlan", "vocab_ranks", "minhash_gpu", "DynaAlignError",
           "nw_pair_stats8", "checksum", "checksum_weights"]


def dimnames(n):
    """list(as.character(1:n), as.character(1:n)) -- what both reference functions attach (src/minHash.cpp:181-185)."""
    labels = [str(i + 1) for i in range(n)]
    return [labels, list(labels)]


def hashfamily_seeds(seed, n_hash):
    """HashFamily(n_hash, seed) seed vector (src/minHash.cpp:73-80)."""
    out = np.zeros(max(int(n_hash), 1), dtype=np.uint32)
    check(lib().dyna_hashfamily_seeds(C.c_uint32(seed), int(n_hash), ptr(out, C.c_uint32)))
    return out[:n_hash]


def similarityMH(sequences, k=4, n_hash=50, *, seed=None, seeds=None, n_gpus=1):
    """Drop-in for the reference's similarityMH.  ``seed``/``seeds`` inject the HashFamily seed (tests);
    by default the seed comes from std::random_device, as in the reference."""
    sequences = list(sequences)
    n = len(sequences)
    res, off = flatten(sequences)
    if seeds is None and seed is not None and n_hash > 0:
        seeds = hashfamily_seeds(seed, n_hash)
    sp = None
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        sp = ptr(seeds, C.c_uint32)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    check(lib().dyna_similarityMH(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, int(k), int(n_hash), sp,
                                  ptr(out, C.c_double), int(n_gpus)))
    return out


def similarityNW(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, *, n_gpus=1):
    """Drop-in for the reference's similarityNW (identity = matches / alignment length of the NW traceback)."""
    sequences = list(sequences)
    n = len(sequences)
    res, off = flatten(sequences)
    out = np.zeros((n, n), dtype=np.float64, order="F")
    check(lib().dyna_similarityNW(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, matrixName.encode(), int(gapOpen),
                                  int(gapExt), ptr(out, C.c_double), int(n_gpus)))
    return out


# ----------------------------------------------------------------------------- pieces of the C++ path
def mh_signatures(sequences, k, seeds):
    sequences = list(sequences)
    res, off = flatten(sequences)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    out = np.zeros((len(sequences), len(seeds)), dtype=np.uint32)
    check(lib().dyna_mh_signatures_murmur3(ptr(res, C.c_uint8), ptr(off, C.c_int64), len(sequences), int(k),
                                           ptr(seeds, C.c_uint32), len(seeds), ptr(out, C.c_uint32)))
    return out


def tri_strict_size(n, row_begin=0, row_end=None):
    row_end = n if row_end is None else row_end
    f = lambda r: r * n - r * (r + 1) // 2
    return f(row_end) - f(row_begin)


def tri_diag_size(n, row_begin=0, row_end=None):
    row_end = n if row_end is None else row_end
    f = lambda r: r * n - r * (r - 1) // 2
    return f(row_end) - f(row_begin)


def mh_match_counts(sig, row_begin=0, row_end=None):
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    n, n_hash = sig.shape
    row_end = n if row_end is None else row_end
    sz = tri_strict_size(n, row_begin, row_end)
    out = np.zeros(max(sz, 1), dtype=np.uint16)
    check(lib().dyna_mh_match_counts(ptr(sig, C.c_uint32), n, n_hash, row_begin, row_end, ptr(out, C.c_uint16)))
    return out[:sz]


def nw_pair_stats(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, row_begin=0, row_end=None):
    sequences = list(sequences)
    n = len(sequences)
    row_end = n if row_end is None else row_end
    res, off = flatten(sequences)
    sz = tri_diag_size(n, row_begin, row_end)
    mt = np.zeros(max(sz, 1), dtype=np.uint32)
    ln = np.zeros(max(sz, 1), dtype=np.uint32)
    check(lib().dyna_nw_pair_stats(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, matrixName.encode(), int(gapOpen),
                                   int(gapExt), row_begin, row_end, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32)))
    return mt[:sz], ln[:sz]


def nw_pair_stats8(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, row_begin=0, row_end=None):
    """nw_pair_stats in the 2-bytes-per-pair host form (every sequence <= 127 residues): (matches u8, length u8)."""
    sequences = list(sequences)
    n = len(sequences)
    row_end = n if row_end is None else row_end
    res, off = flatten(sequences)
    sz = tri_diag_size(n, row_begin, row_end)
    mt = np.zeros(max(sz, 1), dtype=np.uint8)
    ln = np.zeros(max(sz, 1), dtype=np.uint8)
    check(lib().dyna_nw_pair_stats8(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, matrixName.encode(), int(gapOpen),
                                    int(gapExt), row_begin, row_end, ptr(mt, C.c_uint8), ptr(ln, C.c_uint8)))
    return mt[:sz], ln[:sz]


def checksum_weights(first_index, count):
    """w(k) of dyna_*_plan_checksum for global packed pair indices first_index .. first_index+count-1 (numpy uint64,
    wrap-around arithmetic): x = (k+1) * 0x9E3779B97F4A7C15; w = x ^ (x >> 31)."""
    with np.errstate(over="ignore"):
        k = np.arange(first_index, first_index + count, dtype=np.uint64) + np.uint64(1)
        x = k * np.uint64(0x9E3779B97F4A7C15)
        return x ^ (x >> np.uint64(31))


def checksum(values, first_index=0):
    """Host restatement of the device checksum: sum(values[k] * w(first_index + k)) mod 2^64."""
    v = np.asarray(values).astype(np.uint64)
    with np.errstate(over="ignore"):
        return int((v * checksum_weights(first_index, v.size)).sum(dtype=np.uint64))


def partition_rows(n, nshards, weights=None, include_diagonal=False):
    out = np.zeros(nshards + 1, dtype=np.int64)
    wp = None
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.int64)
        wp = ptr(weights, C.c_int64)
    check(lib().dyna_partition_rows(n, wp, 1 if include_diagonal else 0, nshards, ptr(out, C.c_int64)))
    return out


def substitution_matrix(name):
    out = np.zeros((24, 24), dtype=np.int8)
    check(lib().dyna_substitution_matrix(name.encode(), ptr(out, C.c_int8)))
    return out


# ----------------------------------------------------------------------------- threshold + sparsify (clusterbreak's next step)
def quantile_type7_counts(hist, n_hash, prob):
    """quantile(count/n_hash, prob, type=7) from the histogram of match counts -> (threshold, min_count)."""
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    thr, mc = C.c_double(0), C.c_int(0)
    check(lib().dyna_quantile_type7_counts(ptr(hist, C.c_uint64), int(n_hash), float(prob), C.byref(thr), C.byref(mc)))
    return thr.value, mc.value


def quantile_type7_identities(hist, prob):
    """quantile(matches / length, prob, type = 7) from a (matches, length) histogram [max_len + 1, 2 max_len + 1]."""
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    thr = C.c_double(0)
    check(lib().dyna_quantile_type7_identities(ptr(hist, C.c_uint64), hist.shape[0], hist.shape[1], float(prob), C.byref(thr)))
    return thr.value


def identities_at_least(hist, threshold):
    """Number of histogram pairs with 0 < matches / length >= threshold (the exact size of the edge list)."""
    hist = np.asarray(hist)
    m = np.arange(hist.shape[0], dtype=np.float64)[:, None]
    l = np.arange(hist.shape[1], dtype=np.float64)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        keep = (m > 0) & (l > 0) & (m / l >= threshold)
    return int(hist[keep].sum())


def similarityNW_edges(sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, thresh_p=0.8, *, device=0):
    """similarityNW followed by clusterbreak's thresholding (R/clusterbreak.R:217-221) without the dense n x n matrix.
    Returns (threshold, i, j, weight) with 0-based i < j in row-major order and weight = matches / alignment_length."""
    plan = NWPlan(sequences, matrixName, gapOpen, gapExt, device=device)
    try:
        plan.run()
        return plan.threshold_edges(thresh_p)
    finally:
        plan.close()


def similarityMH_edges(sequences, k=4, n_hash=50, thresh_p=0.8, *, seed=None, seeds=None, device=0, join="auto"):
    """similarityMH followed by clusterbreak's thresholding (R/clusterbreak.R:217-221) without ever materialising the
    dense n x n matrix:  threshold <- quantile(sim[upper.tri(sim)], thresh_p);  sim[sim < threshold] <- 0.
    Returns (threshold, i, j, weight) with 0-based i < j in row-major order and weight = count / n_hash.
    join: "auto" (join on equal signature values when the data is sparse enough, else all-pairs), "never", "always"."""
    plan = MinHashPlan(sequences, k, n_hash, seed=seed, seeds=seeds, device=device)
    plan.join = join
    try:
        return plan.threshold_edges(thresh_p)
    finally:
        plan.close()


class MinHashPlan:
    """Device-resident MinHash state for one set of sequences: hash once, then match, threshold and recurse into
    sub-clusters (clusterbreak's loop, R/clusterbreak.R:203-259) without leaving the GPU or re-hashing."""

    _matched = False

    def __init__(self, sequences=None, k=4, n_hash=50, *, seed=None, seeds=None, device=0, _handle=None, _n=None):
        self.n_hash = int(n_hash)
        if _handle is not None:
            self._h, self.n = _handle, _n
            return
        sequences = list(sequences)
        self.n = len(sequences)
        if self.n == 0:
            raise DynaAlignError(L.ERR_INVALID, "Input sequences vector cannot be empty")
        if k <= 0:
            raise DynaAlignError(L.ERR_INVALID, "'k' must be a positive integer")
        if n_hash <= 0:
            raise DynaAlignError(L.ERR_INVALID, "Number of hash functions must be positive")
        if seeds is None:
            seeds = hashfamily_seeds(lib().dyna_random_seed() if seed is None else seed, n_hash)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        res, off = flatten(sequences)
        self._h = lib().dyna_mh_plan_create(self.n, self.n_hash, 0, self.n, int(device))
        if not self._h:
            raise DynaAlignError(L.ERR_CUDA, L.last_error())
        check(lib().dyna_mh_plan_upload_sequences(self._h, ptr(res, C.c_uint8), ptr(off, C.c_int64), int(k), ptr(seeds, C.c_uint32), None))
        check(lib().dyna_mh_plan_run_signatures(self._h, None))

    def close(self):
        if getattr(self, "_h", None):
            lib().dyna_mh_plan_destroy(self._h)
            self._h = None

    __del__ = close

    # "auto": try the join on equal signature values first (exact, O(n * n_hash + matches)); fall back to the all-pairs
    # kernel when the data has too many matches for it to pay or the plan has no sorted rows.  "never" / "always" force one.
    join = "auto"
    joined = False       # the last match ran as a join
    incidences = None    # (pair, hash function) matches the join counted, when it was consulted

    def match_sparse(self, max_incidences=None):
        """Run the join (dyna_mh_plan_run_match_sparse); returns True if it ran.  Default cap: the point where the
        all-pairs kernel (~3e13 compares/s) is certainly faster than sorting the incidences."""
        if max_incidences is None:
            pairs = lib().dyna_mh_plan_pairs(self._h)
            max_incidences = max(1 << 20, pairs * self.n_hash // 4000)
        inc, done = C.c_int64(0), C.c_int(0)
        check(lib().dyna_mh_plan_run_match_sparse(self._h, int(max_incidences), C.byref(inc), C.byref(done), None))
        self.incidences = inc.value if inc.value >= 0 else None
        if done.value:
            self._matched, self.joined = True, True
        return bool(done.value)

    def _match(self):
        if self._matched:
            return
        if self.join != "never" and self.match_sparse(None if self.join == "auto" else 0):
            return
        check(lib().dyna_mh_plan_run_match(self._h, None))
        self._matched, self.joined = True, False

    def signatures(self):
        out = np.zeros((self.n, self.n_hash), dtype=np.uint32)
        check(lib().dyna_mh_plan_fetch_signatures(self._h, ptr(out, C.c_uint32), None))
        return out

    def match_counts(self):
        """The dense u16 triangle on the host (a join result is scattered into it on the device first)."""
        out = np.zeros(max(tri_strict_size(self.n), 1), dtype=np.uint16)
        if self._matched:
            check(lib().dyna_mh_plan_fetch_counts(self._h, ptr(out, C.c_uint16), None))
        else:  # match and copy back chunk by chunk, overlapped
            check(lib().dyna_mh_plan_run_match_fetch(self._h, ptr(out, C.c_uint16), None))
            self._matched = True
        return out[:tri_strict_size(self.n)]

    def match_counts8(self, esc_capacity=1 << 20):
        """Match and fetch in the narrow form: (counts saturated at 255 as uint8, escape pair indices, escape counts);
        the exact u16 triangle is counts8 with counts8[escape index] replaced by the escape count."""
        sz = tri_strict_size(self.n)
        out = np.zeros(max(sz, 1), dtype=np.uint8)
        ei = np.zeros(max(esc_capacity, 1), dtype=np.int64)
        ec = np.zeros(max(esc_capacity, 1), dtype=np.uint16)
        ne = C.c_int64(0)
        check(lib().dyna_mh_plan_run_match_fetch8(self._h, ptr(out, C.c_uint8), int(esc_capacity), ptr(ei, C.c_int64),
                                                  ptr(ec, C.c_uint16), C.byref(ne), None))
        self._matched = True
        order = np.argsort(ei[:ne.value], kind="stable")
        return out[:sz], ei[:ne.value][order], ec[:ne.value][order]

    def checksum(self):
        self._match()
        h = C.c_uint64(0)
        check(lib().dyna_mh_plan_checksum(self._h, C.byref(h), None))
        return int(h.value)

    def histogram(self):
        self._match()
        hist = np.zeros(self.n_hash + 1, dtype=np.uint64)
        check(lib().dyna_mh_plan_count_histogram(self._h, ptr(hist, C.c_uint64), None))
        return hist

    def threshold_edges(self, thresh_p):
        """(threshold, i, j, weight): quantile(sim[upper.tri], thresh_p) and the pairs with sim >= threshold (sim > 0)."""
        hist = self.histogram()
        thr, mc = quantile_type7_counts(hist, self.n_hash, thresh_p)
        cap = int(hist[max(mc, 1):].sum())
        ei = np.zeros(max(cap, 1), dtype=np.int32)
        ej = np.zeros(max(cap, 1), dtype=np.int32)
        ec = np.zeros(max(cap, 1), dtype=np.uint16)
        ne = C.c_int64(0)
        check(lib().dyna_mh_plan_threshold_edges(self._h, mc, cap, ptr(ei, C.c_int32), ptr(ej, C.c_int32), ptr(ec, C.c_uint16),
                                                 C.byref(ne), None))
        m = ne.value
        return thr, ei[:m], ej[:m], ec[:m].astype(np.float64) / self.n_hash

    def subset(self, indices):
        """Plan over sequences[indices] (0-based), reusing this plan's signatures."""
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        h = lib().dyna_mh_plan_create_subset(self._h, ptr(idx, C.c_int64), len(idx), 0, len(idx))
        if not h:
            raise DynaAlignError(L.ERR_INVALID, L.last_error())
        child = MinHashPlan(n_hash=self.n_hash, _handle=h, _n=len(idx))
        child.join = self.join
        return child


class NWPlan:
    """Device-resident Needleman-Wunsch state for one set of sequences and one row range: validate, encode and build
    the work units once, then run (and re-run) the kernels and fetch the (matches, length) slab when wanted."""

    def __init__(self, sequences, matrixName="BLOSUM62", gapOpen=10, gapExt=4, row_begin=0, row_end=None, device=0):
        sequences = list(sequences)
        self.n = len(sequences)
        self.row_begin, self.row_end = int(row_begin), self.n if row_end is None else int(row_end)
        res, off = flatten(sequences)
        self._h = lib().dyna_nw_plan_create(ptr(res, C.c_uint8), ptr(off, C.c_int64), self.n, matrixName.encode(), int(gapOpen),
                                            int(gapExt), self.row_begin, self.row_end, int(device))
        if not self._h:
            msg = L.last_error()
            raise DynaAlignError(L.ERR_CUDA if msg.startswith("DynaAlign CUDA") else L.ERR_INVALID, msg)

    pairs = property(lambda self: lib().dyna_nw_plan_pairs(self._h))
    cells = property(lambda self: lib().dyna_nw_plan_cells(self._h))

    def run(self, stream=None):
        check(lib().dyna_nw_plan_run(self._h, stream))
        return self

    def fetch(self):
        """(matches, length) of the plan's pairs, packed upper triangle (diagonal included), row-major."""
        sz = self.pairs
        mt = np.zeros(max(sz, 1), dtype=np.uint32)
        ln = np.zeros(max(sz, 1), dtype=np.uint32)
        check(lib().dyna_nw_plan_fetch(self._h, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
        return mt[:sz], ln[:sz]

    def fetch_packed8(self):
        """(matches, length) as uint8 each (2 bytes per pair); only when every alignment length fits a byte."""
        sz = self.pairs
        mt = np.zeros(max(sz, 1), dtype=np.uint8)
        ln = np.zeros(max(sz, 1), dtype=np.uint8)
        check(lib().dyna_nw_plan_fetch_packed8(self._h, ptr(mt, C.c_uint8), ptr(ln, C.c_uint8), None))
        return mt[:sz], ln[:sz]

    def checksum(self):
        """(matches, length) position-weighted checksums of the plan's slab (see checksum())."""
        h = (C.c_uint64 * 2)()
        check(lib().dyna_nw_plan_checksum(self._h, h, None))
        return int(h[0]), int(h[1])

    # ---- threshold + sparsify (clusterbreak's next step, R/clusterbreak.R:217-221) on the computed triangle
    max_len = property(lambda self: lib().dyna_nw_plan_max_len(self._h))

    @staticmethod
    def _members(members):
        if members is None:
            return None, None, 0
        m = np.ascontiguousarray(members, dtype=np.int32)
        return m, ptr(m, C.c_int32), len(m)

    def stat_histogram(self, members=None):
        """uint64[max_len + 1, 2 max_len + 1] counts of (matches, length) over the strict upper triangle of the node
        (`members`: strictly increasing 0-based indices, None = all sequences) within the plan's row range."""
        ml = self.max_len
        hist = np.zeros((ml + 1, 2 * ml + 1), dtype=np.uint64)
        _keep, mp, nm = self._members(members)
        check(lib().dyna_nw_plan_stat_histogram(self._h, mp, nm, ptr(hist, C.c_uint64), None))
        return hist

    def diagonal(self, members=None):
        """(matches, length) of the node's self-alignments: sim[i, i], the self-loop weights netcluster's graph carries."""
        _keep, mp, nm = self._members(members)
        cnt = self.n if members is None else nm
        mt = np.zeros(max(cnt, 1), dtype=np.uint32)
        ln = np.zeros(max(cnt, 1), dtype=np.uint32)
        check(lib().dyna_nw_plan_fetch_diagonal(self._h, mp, nm, ptr(mt, C.c_uint32), ptr(ln, C.c_uint32), None))
        return mt[:cnt], ln[:cnt]

    def edges_at(self, threshold, capacity, members=None):
        """Pairs a < b of the node with 0 < (double)matches/length >= threshold: (i, j, matches, length), row-major,
        node-local 0-based indices."""
        _keep, mp, nm = self._members(members)
        cap = max(int(capacity), 1)
        ei = np.zeros(cap, dtype=np.int32)
        ej = np.zeros(cap, dtype=np.int32)
        em = np.zeros(cap, dtype=np.uint32)
        el = np.zeros(cap, dtype=np.uint32)
        ne = C.c_int64(0)
        check(lib().dyna_nw_plan_threshold_edges(self._h, mp, nm, float(threshold), int(capacity), ptr(ei, C.c_int32),
                                                 ptr(ej, C.c_int32), ptr(em, C.c_uint32), ptr(el, C.c_uint32), C.byref(ne), None))
        k = ne.value
        return ei[:k], ej[:k], em[:k], el[:k]

    def threshold_edges(self, thresh_p, members=None):
        """(threshold, i, j, weight): quantile(sim[upper.tri(sim)], thresh_p) (type 7, exact) and the pairs that survive
        `sim[sim < threshold] <- 0` with a non-zero similarity; weight = matches / length as the reference computes it."""
        hist = self.stat_histogram(members)
        thr = quantile_type7_identities(hist, thresh_p)
        cap = identities_at_least(hist, thr)
        ei, ej, em, el = self.edges_at(thr, cap, members)
        return thr, ei, ej, em.astype(np.float64) / el.astype(np.float64)

    def close(self):
        if getattr(self, "_h", None):
            lib().dyna_nw_plan_destroy(self._h)
            self._h = None

    __del__ = close


# ----------------------------------------------------------------------------- R pipeline (R/minHash.R)
class RError(DynaAlignError):
    def __init__(self, message):
        super().__init__(L.ERR_INVALID, message)


def shingle(x, k):
    """R/minHash.R:12-23 (host string handling, as in R)."""
    if not isinstance(x, str):
        raise RError("Input 'x' must be a single character string")
    if isinstance(k, bool) or not isinstance(k, (int, float, np.integer, np.floating)) or k < 1 or k > len(x):
        raise RError("'k' must be a positive integer between 1 and %d" % len(x))
    k = int(k)
    return [x[i:i + k] for i in range(len(x) - k + 1)]


def create_vocab(sequences, k):
    """R/minHash.R:38-41: sort(unique(all shingles)); byte order (== R's collation for upper-case ASCII)."""
    seen = set()
    for s in sequences:
        seen.update(shingle(s, k))
    return sorted(seen)


def create_char_matrix(sequences, vocab, k):
    """R/minHash.R:60-66: V x N integer 0/1 matrix."""
    pos = {v: i for i, v in enumerate(vocab)}
    m = np.zeros((len(vocab), len(sequences)), dtype=np.int32)
    for j, s in enumerate(sequences):
        for sh in shingle(s, k):
            i = pos.get(sh)
            if i is not None:
                m[i, j] = 1
    return m


def create_hash_parameters(n_hash, max_val, rng=None):
    """R/minHash.R:81-88.  R's sample() stream cannot be reproduced outside R; ranges and lengths follow the source."""
    if n_hash < 1:
        raise RError("Number of hash functions must be positive")
    if max_val < 2:
        raise RError("Maximum value must be at least 2")
    rng = np.random.default_rng() if rng is None else rng
    return {"a": rng.integers(1, max_val + 1, size=n_hash, dtype=np.int64),
            "b": rng.integers(0, max_val + 1, size=n_hash, dtype=np.int64)}


def apply_hash(x, a, b, m):
    """R/minHash.R:104-106."""
    return (np.asarray(a) * np.asarray(x) + np.asarray(b)) % m


def compute_signature_matrix(char_matrix, hash_params, max_val):
    """R/minHash.R:126-143 on the GPU: n_hash x n_docs double matrix (Inf where a document has no shingle)."""
    char_matrix = np.asarray(char_matrix)
    a = np.ascontiguousarray(hash_params["a"], dtype=np.int64)
    b = np.ascontiguousarray(hash_params["b"], dtype=np.int64)
    n_docs = char_matrix.shape[1]
    rows, docs = np.nonzero(char_matrix.T == 1)[::-1]  # per document, ascending rank
    order = np.lexsort((rows, docs))
    ranks = (rows[order] + 1).astype(np.int32)
    counts = np.bincount(docs, minlength=n_docs)
    roff = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(counts, out=roff[1:])
    if ranks.size == 0:
        ranks = np.ones(1, dtype=np.int32)
    sig = np.zeros((n_docs, len(a)), dtype=np.uint32)
    check(lib().dyna_mh_signatures_linear(ptr(ranks, C.c_int32), ptr(roff, C.c_int64), n_docs, ptr(a, C.c_int64),
                                          ptr(b, C.c_int64), int(max_val), len(a), ptr(sig, C.c_uint32)))
    out = sig.T.astype(np.float64)
    out[sig.T == np.uint32(0xFFFFFFFF)] = np.inf
    return out


def compute_distance_matrix(sig_matrix):
    """R/minHash.R:166-182 on the GPU.  Arbitrary doubles are relabelled per hash row to dense integer codes
    (equality-preserving), then the match-count kernel runs with the R distance table."""
    sig_matrix = np.asarray(sig_matrix, dtype=np.float64)
    if np.isnan(sig_matrix).any():
        raise RError("NA/NaN in signature matrix")
    n_hash, n_docs = sig_matrix.shape
    codes = np.empty((n_docs, n_hash), dtype=np.uint32)
    for h in range(n_hash):
        codes[:, h] = np.unique(sig_matrix[h], return_inverse=True)[1].astype(np.uint32)
    out = np.zeros((n_docs, n_docs), dtype=np.float64, order="F")
    check(lib().dyna_mh_match_matrix(ptr(codes, C.c_uint32), n_docs, n_hash, L.MH_DISTANCE, ptr(out, C.c_double), 1))
    return out


def vocab_ranks(sequences, k):
    """create_vocab + the vocabulary rank of every shingle, on the device (no dense V x N matrix).
    Returns (vocabulary list[str], ranks int32[], offsets int64[n+1])."""
    sequences = list(sequences)
    res, off = flatten(sequences)
    n = len(sequences)
    total = int(sum(max(len(s) - k + 1, 0) for s in sequences))
    ranks = np.zeros(max(total, 1), dtype=np.int32)
    roff = np.zeros(n + 1, dtype=np.int64)
    keys = np.zeros(max(total, 1), dtype=np.uint64)
    V = C.c_int64(0)
    check(lib().dyna_minhash_vocab_ranks(ptr(res, C.c_uint8), ptr(off, C.c_int64), n, int(k), ptr(keys, C.c_uint64), len(keys),
                                         C.byref(V), ptr(ranks, C.c_int32), ptr(roff, C.c_int64)))
    vocab = [int(x).to_bytes(8, "big")[8 - k:].decode("latin-1") for x in keys[:V.value]]
    return vocab, ranks[:total], roff


def minhash_gpu(sequences, k, n_hash, rng=None, hash_params=None):
    """minhash() with the whole pipeline on the device: vocabulary by sort/unique, signatures from rank lists, distance by
    the match kernel.  Same list as minhash() except that char_matrix is not materialised (ranks/offsets instead)."""
    sequences = list(sequences)
    vocab, ranks, roff = vocab_ranks(sequences, k)
    max_val = len(vocab)
    if hash_params is None:
        hash_params = create_hash_parameters(n_hash, max_val, rng)
    a = np.ascontiguousarray(hash_params["a"], dtype=np.int64)
    b = np.ascontiguousarray(hash_params["b"], dtype=np.int64)
    n = len(sequences)
    sig = np.zeros((n, len(a)), dtype=np.uint32)
    rk = ranks if ranks.size else np.ones(1, dtype=np.int32)
    check(lib().dyna_mh_signatures_linear(ptr(rk, C.c_int32), ptr(roff, C.c_int64), n, ptr(a, C.c_int64), ptr(b, C.c_int64),
                                          int(max_val), len(a), ptr(sig, C.c_uint32)))
    dist = np.zeros((n, n), dtype=np.float64, order="F")
    check(lib().dyna_mh_match_matrix(ptr(sig, C.c_uint32), n, len(a), L.MH_DISTANCE, ptr(dist, C.c_double), 1))
    return {"vocabulary": vocab, "ranks": ranks, "rank_offsets": roff, "sig_matrix": sig.T.astype(np.float64), "dist_matrix": dist}


def minhash(sequences, k, n_hash, rng=None, hash_params=None):
    """R/minHash.R:206-221."""
    sequences = list(sequences)
    vocab = create_vocab(sequences, k)
    char_matrix = create_char_matrix(sequences, vocab, k)
    max_val = len(vocab)
    if hash_params is None:
        hash_params = create_hash_parameters(n_hash, max_val, rng)
    sig_matrix = compute_signature_matrix(char_matrix, hash_params, max_val)
    dist_matrix = compute_distance_matrix(sig_matrix)
    return {"vocabulary": vocab, "char_matrix": char_matrix, "sig_matrix": sig_matrix, "dist_matrix": dist_matrix}
